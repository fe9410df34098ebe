"""ICM and Disagreement on the GPU vs the reference's golden vectors and the oracle.
Tolerance (north_star): bonus / loss within 1e-2 relative (bf16 operands, fp32 accumulation)."""
import pytest
import torch

import recipes as R
from helpers import golden, check_grads
from oracle import curiosity_oracle as O
from curiosity_baselines_b200.curiosity.icm import ICM, Disagreement

pytestmark = pytest.mark.gpu
DEV = "cuda"


def relmax(a, b):
    a, b = a.detach().cpu().double(), torch.as_tensor(b).double()
    return float(((a - b).abs() / b.abs().clamp_min(1e-12)).max())


def test_state_dict_keys_match_reference():
    m = ICM(image_shape=(4, 84, 84), action_size=4, feature_encoding="idf_burda")
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == dict(R.icm_shapes(4, "idf_burda"))
    m = ICM(image_shape=(4, 84, 84), action_size=7, feature_encoding="idf")
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == dict(R.icm_shapes(7, "idf"))
    d = Disagreement(image_shape=(4, 84, 84), action_size=7, feature_encoding="idf_burda")
    assert {k: tuple(v.shape) for k, v in d.state_dict().items()} == dict(R.disagreement_shapes(7, "idf_burda"))
    assert m.feature_size == 288 and d.feature_size == 512


@pytest.mark.parametrize("enc,A", [("idf_burda", 4), ("idf", 7)])
def test_icm_golden(enc, A):
    g = golden(f"icm_{enc}")
    T, B, seed = int(g["T"]), int(g["B"]), int(g["seed"])
    m = ICM(image_shape=(4, 84, 84), action_size=A, feature_encoding=enc, prediction_beta=0.7,
            forward_loss_wt=0.2).to(DEV)
    m.load_state_dict(R.make_params(R.icm_shapes(A, enc), seed))
    obs, nxt = R.make_frames(seed + 1, T, B).to(DEV), R.make_frames(seed + 2, T, B).to(DEV)
    act = R.onehot(R.make_actions(seed + 3, T, B, A), A).to(DEV)
    valid = (1 - R.make_suffix_done(seed + 4, T, B, frac=0.5)).to(DEV)
    bonus = m.compute_bonus(obs, nxt, act)
    assert relmax(bonus, g["bonus"]) < 1e-2
    inv, fwd = m.compute_loss(obs, nxt, act, valid)
    (inv + fwd).backward()
    assert abs(float(inv) - float(g["inv_loss"])) / float(g["inv_loss"]) < 1e-2
    assert abs(float(fwd) - float(g["fwd_loss"])) / float(g["fwd_loss"]) < 1e-2
    grads = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    assert len(grads) == len(list(m.parameters()))
    # 12-15 samples: one bf16-flipped LeakyReLU mask is visible in the deepest conv gradients (measured worst
    # cosine 0.9939, norm 3.0 %); the SURVEY 8(d) gate 0.999 / 1e-2 is enforced on >= 1024 samples in
    # tests/test_gpu_parity_scale.py
    bad = check_grads(grads, g, cos_min=0.99, norm_tol=4e-2)
    assert not bad, bad


def test_disagreement_golden():
    g = golden("disagreement")
    T, B, A, seed = int(g["T"]), int(g["B"]), int(g["A"]), int(g["seed"])
    m = Disagreement(image_shape=(4, 84, 84), action_size=A, feature_encoding="idf_burda", prediction_beta=1.0,
                     forward_loss_wt=0.2).to(DEV)
    m.load_state_dict(R.make_params(R.disagreement_shapes(A, "idf_burda"), seed))
    obs, nxt = R.make_frames(seed + 1, T, B).to(DEV), R.make_frames(seed + 2, T, B).to(DEV)
    act = R.onehot(R.make_actions(seed + 3, T, B, A), A).to(DEV)
    valid = (1 - R.make_suffix_done(seed + 4, T, B, frac=0.5)).to(DEV)
    bonus = m.compute_bonus(obs, nxt, act)
    assert relmax(bonus, g["bonus"]) < 1e-2          # measured 1.6e-3 (two-pass variance of the fp32 accumulators)
    masks = [R.make_mask(seed + 10 + e, (T, B, 512), 0.2).to(DEV) for e in range(4)]
    inv, fwd = m.compute_loss(obs, nxt, act, valid, dropout_masks=masks)
    (inv + fwd).backward()
    assert abs(float(inv) - float(g["inv_loss"])) / float(g["inv_loss"]) < 1e-2
    assert abs(float(fwd) - float(g["fwd_loss"])) / float(g["fwd_loss"]) < 1e-2
    grads = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    bad = check_grads(grads, g, cos_min=0.99, norm_tol=4e-2)      # few-sample case, see test_icm_golden
    assert not bad, bad
    inv2, fwd2 = m.compute_loss(obs, nxt, act, valid, dropout_masks=None)
    assert abs(float(fwd2) - float(g["fwd_loss_nodrop"])) / float(g["fwd_loss_nodrop"]) < 1e-2
    _, fwd3 = m.compute_loss(obs, nxt, act, valid)          # default: random dropout, same expectation
    assert abs(float(fwd3) - float(fwd2)) / float(fwd2) < 0.1


@pytest.mark.parametrize("enc,A,T,B", [("idf_burda", 4, 16, 8), ("idf_maze", 5, 12, 6)])
def test_icm_oracle_parity(enc, A, T, B):
    seed = 300 + T
    shape = (4, 84, 84) if enc != "idf_maze" else (3, 5, 5)
    params = R.make_params(R.icm_shapes(A, enc, c_in=shape[0]) if enc != "idf_maze" else
                           R.maze_shapes("encoder.model.", 3) + R.icm_shapes(A, "idf_burda")[8:12] +
                           R.res_forward_shapes("forward_model.", 256, A), seed)
    if enc == "idf_maze":      # rebuild the inverse-model shapes for F = 256
        shapes = R.maze_shapes("encoder.model.", 3) + [
            ("inverse_model.0.weight", (256, 512)), ("inverse_model.0.bias", (256,)),
            ("inverse_model.2.weight", (A, 256)), ("inverse_model.2.bias", (A,))] + R.res_forward_shapes("forward_model.", 256, A)
        params = R.make_params(shapes, seed)
        obs, nxt = R.make_binary_frames(seed + 1, T, B), R.make_binary_frames(seed + 2, T, B)
    else:
        obs, nxt = R.make_frames(seed + 1, T, B), R.make_frames(seed + 2, T, B)
    m = ICM(image_shape=shape, action_size=A, feature_encoding=enc, prediction_beta=1.0, forward_loss_wt=0.2).to(DEV)
    m.load_state_dict(params)
    act = R.onehot(R.make_actions(seed + 3, T, B, A), A)
    valid = 1 - R.make_suffix_done(seed + 4, T, B, frac=0.3)
    b_ref = O.icm_bonus(params, obs, nxt, act, enc, 1.0)
    assert relmax(m.compute_bonus(obs.to(DEV), nxt.to(DEV), act.to(DEV)), b_ref) < 1e-2
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    inv_r, fwd_r = O.icm_loss(p, obs, nxt, act, valid, enc, 0.2)
    (inv_r + fwd_r).backward()
    inv, fwd = m.compute_loss(obs.to(DEV), nxt.to(DEV), act.to(DEV), valid.to(DEV))
    (inv + fwd).backward()
    assert abs(float(inv) - float(inv_r)) / float(inv_r) < 1e-2
    assert abs(float(fwd) - float(fwd_r)) / float(fwd_r) < 1e-2
    for k, prm in m.named_parameters():
        gref, gg = p[k].grad.flatten(), prm.grad.cpu().flatten()
        cos = float(gg @ gref / (gg.norm() * gref.norm()).clamp_min(1e-30))
        # 72-128 samples: bias gradients are signed sums of bf16-rounded dy rows (cancellation), 8 % on their norm;
        # the strict 0.999 / 1e-2 gate runs at n >= 1024 in test_gpu_parity_scale.py
        tol = 8e-2 if gg.numel() <= 64 else 3e-2
        assert cos > 0.99 and abs(float(gg.norm() / gref.norm()) - 1) < tol, (k, cos, float(gg.norm() / gref.norm()))


@pytest.mark.parametrize("cls_name", ["icm", "disagreement"])
def test_forward_reuse_between_bonus_and_loss(cls_name):
    """reuse_features: compute_loss on the batch compute_bonus just saw (what the reference passes) reuses the whole
    forward pass after an equality check of observations, next observations and actions -- identical losses -- and
    recomputes for any other batch or after a weight update."""
    T, B, A, seed = 5, 4, 7, 23
    cls = ICM if cls_name == "icm" else Disagreement
    shapes = R.icm_shapes(A, "idf_burda") if cls_name == "icm" else R.disagreement_shapes(A, "idf_burda")
    params = R.make_params(shapes, seed)
    obs, nxt = R.make_frames(seed + 1, T, B).to(DEV), R.make_frames(seed + 2, T, B).to(DEV)
    act = R.onehot(R.make_actions(seed + 3, T, B, A), A).to(DEV)
    valid = torch.ones(T, B, device=DEV)

    def run(reuse):
        m = cls(image_shape=(4, 84, 84), action_size=A, feature_encoding="idf_burda").to(DEV)
        m.load_state_dict(params)
        m.reuse_features = reuse
        bonus = m.compute_bonus(obs, nxt, act)
        inv, fwd = m.compute_loss(obs.clone(), nxt.clone(), act.clone(), valid, dropout_masks=None)
        (inv + fwd).backward()
        return m, bonus, inv.detach(), fwd.detach(), {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}

    _, b0, i0, f0, g0 = run(False)
    m1, b1, i1, f1, g1 = run(True)
    assert m1.reuse_hits == {"full": 1, "miss": 0}
    assert torch.equal(b0, b1) and torch.equal(i0, i1) and torch.equal(f0, f1)
    for k in g0:      # weight gradients are summed with fp32 atomics: equal up to summation order
        assert float((g0[k] - g1[k]).abs().max()) <= 1e-4 * float(g0[k].abs().max()) + 1e-12, k
    other = nxt.clone(); other[0, 0, 0, 0, 0] ^= 1
    m1.compute_loss(obs, other, act, valid, dropout_masks=None)
    assert m1.reuse_hits["miss"] == 1
    with torch.no_grad():
        next(m1.parameters()).add_(0.0)          # bumps the parameter's version: weights may have changed
    m1.compute_loss(obs, nxt, act, valid, dropout_masks=None)
    assert m1.reuse_hits["miss"] == 2
