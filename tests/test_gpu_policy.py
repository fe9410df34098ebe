"""The policy network (AtariLstmModel) and the PPO loss / update on the GPU vs the reference's golden vectors
(tests/golden/policy_ppo_icm.npz, produced by the reference's own agent + PPO.loss) and the CPU oracle.
Tolerances (north_star): outputs and losses within 1e-2 relative (bf16 operands, fp32 accumulation); the
elementwise fp32 kernels (LSTM cell, PPO loss) within 1e-5 of the same formula in torch fp32."""
import numpy as np
import pytest
import torch

import recipes as R
from helpers import golden, check_grads, rel_err, cosine
from oracle import curiosity_oracle as O
from test_oracle_golden import policy_case
from curiosity_baselines_b200 import _lib as L
from curiosity_baselines_b200 import layers as LY
from curiosity_baselines_b200.algos.ppo import PPO, PpoBatch
from curiosity_baselines_b200.pg.atari_lstm_model import AtariLstmModel

pytestmark = pytest.mark.gpu
DEV = "cuda"
ICM_KW = dict(curiosity_alg="icm", feature_encoding="idf_burda", batch_norm=False, prediction_beta=1.0,
              forward_loss_wt=0.2)


def make_model(A, params, icm_params):
    m = AtariLstmModel((4, 84, 84), A, curiosity_kwargs=ICM_KW).to(DEV)
    assert m._fused_heads() == (A <= 7)
    sd = dict(params)
    sd.update({"curiosity_model." + k: v for k, v in icm_params.items()})
    assert set(sd) == set(m.state_dict())
    m.load_state_dict(sd)
    return m


def test_state_dict_keys_match_reference():
    m = AtariLstmModel((4, 84, 84), 4, curiosity_kwargs=ICM_KW)
    got = {k: tuple(v.shape) for k, v in m.state_dict().items() if not k.startswith("curiosity_model.")}
    assert got == dict(R.policy_shapes(4, "idf_burda"))


def test_lstm_cell_kernels_match_torch_fp32():
    B, H = 37, 512
    g = torch.Generator(device="cpu").manual_seed(5)
    gx, gh = torch.randn(B, 4 * H, generator=g).to(DEV), torch.randn(B, 4 * H, generator=g).to(DEV)
    c_prev = torch.randn(B, H, generator=g).to(DEV).requires_grad_(True)
    a = (gx + gh).requires_grad_(True)
    ai, af, ag, ao = a.chunk(4, 1)
    c_ref = torch.sigmoid(af) * c_prev + torch.sigmoid(ai) * torch.tanh(ag)
    h_ref = torch.sigmoid(ao) * torch.tanh(c_ref)
    gates, c = torch.empty(B, 4 * H, device=DEV), torch.empty(B, H, device=DEV)
    hb, hf = torch.empty(B, H, dtype=torch.bfloat16, device=DEV), torch.empty(B, H, device=DEV)
    L.call("cb_lstm_cell_fwd", L.ptr(gx), L.ptr(gh), L.ptr(c_prev.detach()), L.ptr(gates), L.ptr(c), L.ptr(hb), L.ptr(hf),
           B, H, L.stream())
    assert torch.allclose(c, c_ref, rtol=1e-5, atol=1e-6) and torch.allclose(hf, h_ref, rtol=1e-5, atol=1e-6)
    assert torch.equal(hb, hf.to(torch.bfloat16))
    ref_gates = torch.cat([torch.sigmoid(ai), torch.sigmoid(af), torch.tanh(ag), torch.sigmoid(ao)], 1)
    assert torch.allclose(gates, ref_gates, rtol=1e-5, atol=1e-6)
    # in-place form (gates aliasing gx) gives the same result
    gx2 = gx.clone()
    L.call("cb_lstm_cell_fwd", L.ptr(gx2), L.ptr(gh), L.ptr(c_prev.detach()), L.ptr(gx2), L.ptr(c), L.ptr(hb), None, B, H,
           L.stream())
    assert torch.equal(gx2, gates)
    # backward
    dh = torch.randn(B, H, generator=g).to(DEV).to(torch.bfloat16)
    dc_next = torch.randn(B, H, generator=g).to(DEV)
    (h_ref * dh.float()).sum().backward(retain_graph=True)
    (c_ref * dc_next).sum().backward()
    dg, dcp = torch.empty(B, 4 * H, dtype=torch.bfloat16, device=DEV), torch.empty(B, H, device=DEV)
    L.call("cb_lstm_cell_bwd", L.ptr(dh), L.ptr(dc_next), L.ptr(gates), L.ptr(c_prev.detach()), L.ptr(c), L.ptr(dg), L.ptr(dcp),
           B, H, L.stream())
    assert torch.allclose(dcp, c_prev.grad, rtol=1e-4, atol=1e-6)
    assert torch.allclose(dg.float(), a.grad, rtol=1e-2, atol=1e-4)          # bf16 output


def test_ppo_loss_kernel_matches_oracle():
    n, A = 1000, 7
    g = torch.Generator(device="cpu").manual_seed(9)
    logits = (torch.randn(n, A, generator=g) * 2).requires_grad_(True)
    value = torch.randn(n, generator=g).requires_grad_(True)
    action = torch.randint(0, A, (n,), generator=g)
    old = torch.softmax(logits.detach() + 0.3 * torch.randn(n, A, generator=g), -1)    # ratios on both sides of the clip
    adv, ret = torch.randn(n, generator=g), torch.randn(n, generator=g)
    valid = (torch.rand(n, generator=g) > 0.2).float()
    pi = torch.softmax(logits, -1)
    ref = O.ppo_loss(pi.view(n, 1, A), value.view(n, 1), action.view(n, 1), old.view(n, 1, A), ret.view(n, 1),
                     adv.view(n, 1), valid.view(n, 1), 0.1, 0.5, 0.01)
    ref[0].backward()
    d = lambda t: t.detach().to(DEV).contiguous()
    coef = d(valid / valid.sum())
    terms = torch.empty(4, n, device=DEV)
    prob, dl, dv = torch.empty(n, A, device=DEV), torch.empty(n, A, device=DEV), torch.empty(n, device=DEV)
    lg_d, v_d, a_d, o_d, adv_d, ret_d = d(logits), d(value), d(action), d(old), d(adv), d(ret)   # keep the operands alive
    L.call("cb_ppo_loss", L.ptr(lg_d), L.ptr(v_d), L.ptr(a_d), L.ptr(o_d), L.ptr(adv_d), L.ptr(ret_d),
           L.ptr(coef), 0.1, 0.5, 0.01, L.ptr(prob), L.ptr(terms), L.ptr(dl), L.ptr(dv), n, A, L.stream())
    means = (terms * coef).sum(1).cpu()
    loss = -means[0] + 0.5 * means[1] - 0.01 * means[2]
    assert abs(float(loss) - float(ref[0])) < 1e-5 * max(1.0, abs(float(ref[0])))
    assert abs(float(means[2]) - float(ref[4])) < 1e-5 and abs(float(means[3]) - float(ref[5])) < 1e-4
    assert torch.allclose(prob.cpu(), pi.detach(), rtol=1e-5, atol=1e-7)
    assert torch.allclose(dl.cpu(), logits.grad, rtol=1e-4, atol=1e-8)
    assert torch.allclose(dv.cpu(), value.grad, rtol=1e-5, atol=1e-9)
    # softmax forward / backward pair used by the autograd drop-in
    p2, dp = torch.empty(n, A, device=DEV), torch.randn(n, A, generator=g).to(DEV)
    L.call("cb_softmax_fwd", L.ptr(lg_d), L.ptr(p2), n, A, L.stream())
    assert torch.allclose(p2.cpu(), pi.detach(), rtol=1e-5, atol=1e-7)
    lg = d(logits).requires_grad_(True)
    (torch.softmax(lg, -1) * dp).sum().backward()
    dl2 = torch.empty(n, A, device=DEV)
    L.call("cb_softmax_bwd", L.ptr(p2), L.ptr(dp), L.ptr(dl2), n, A, L.stream())
    assert torch.allclose(dl2, lg.grad, rtol=1e-4, atol=2e-6)      # p*(dp - sum p dp): cancellation at fp32 epsilon


def test_fused_heads_match_torch_fp32():
    n, H, A = 1003, 512, 6
    g = torch.Generator(device="cpu").manual_seed(17)
    h = torch.randn(n, H, generator=g).to(DEV).to(torch.bfloat16)
    w_pi, b_pi = (torch.randn(A, H, generator=g) * 0.05).to(DEV), torch.randn(A, generator=g).to(DEV)
    w_v, b_v = (torch.randn(1, H, generator=g) * 0.05).to(DEV), torch.randn(1, generator=g).to(DEV)
    assert L.load().cb_policy_heads_supported(H, A + 1) == 1 and L.load().cb_policy_heads_supported(H, 9) == 0
    logits, value, prob = torch.empty(n, A, device=DEV), torch.empty(n, device=DEV), torch.empty(n, A, device=DEV)
    L.call("cb_policy_heads_fwd", L.ptr(h), L.ptr(w_pi), L.ptr(b_pi), L.ptr(w_v), L.ptr(b_v), L.ptr(logits), L.ptr(value),
           L.ptr(prob), n, H, A, L.stream())
    wq, vq = w_pi.to(torch.bfloat16).float(), w_v.to(torch.bfloat16).float()        # the kernel's operand rounding
    ref_l = h.float() @ wq.t() + b_pi
    ref_v = (h.float() @ vq.t()).squeeze(1) + b_v
    assert torch.allclose(logits, ref_l, rtol=1e-4, atol=1e-4) and torch.allclose(value, ref_v, rtol=1e-4, atol=1e-4)
    assert torch.allclose(prob, torch.softmax(ref_l, -1), rtol=1e-4, atol=1e-6)
    assert rel_err(logits.cpu(), (h.float() @ w_pi.t() + b_pi).cpu()) < 1e-2       # vs unrounded fp32 weights
    dl, dvv = torch.randn(n, A, generator=g).to(DEV), torch.randn(n, generator=g).to(DEV)
    dh = torch.empty(n, H, dtype=torch.bfloat16, device=DEV)
    dw_pi, db_pi = torch.full((A, H), 7.0, device=DEV), torch.full((A,), 7.0, device=DEV)     # overwritten, not accumulated
    dw_v, db_v = torch.full((1, H), 7.0, device=DEV), torch.full((1,), 7.0, device=DEV)
    L.call("cb_policy_heads_bwd", L.ptr(h), L.ptr(dl), L.ptr(dvv), L.ptr(w_pi), L.ptr(w_v), L.ptr(dh), L.ptr(dw_pi),
           L.ptr(db_pi), L.ptr(dw_v), L.ptr(db_v), n, H, A, L.ACT_NONE, L.stream())
    ref_dh = dl @ wq + dvv.unsqueeze(1) @ vq
    assert torch.allclose(dh.float(), ref_dh, rtol=1e-2, atol=1e-2)                 # bf16 output
    assert torch.allclose(dw_pi, dl.t() @ h.float(), rtol=1e-3, atol=1e-3)
    assert torch.allclose(dw_v, (dvv.unsqueeze(0) @ h.float()), rtol=1e-3, atol=1e-3)
    assert torch.allclose(db_pi, dl.sum(0), rtol=1e-4, atol=1e-3) and torch.allclose(db_v, dvv.sum().reshape(1), rtol=1e-4, atol=1e-3)


def test_fused_narrow_linear_without_value_head():
    """The same kernels as a plain Linear(H -> A) over a ReLU output (ICM's inverse-model head, icm.py:88-92)."""
    n, H, A = 777, 512, 7
    g = torch.Generator(device="cpu").manual_seed(23)
    h = torch.relu(torch.randn(n, H, generator=g)).to(DEV).to(torch.bfloat16)
    w, b = (torch.randn(A, H, generator=g) * 0.05).to(DEV), torch.randn(A, generator=g).to(DEV)
    logits = torch.empty(n, A, device=DEV)
    L.call("cb_policy_heads_fwd", L.ptr(h), L.ptr(w), L.ptr(b), None, None, L.ptr(logits), None, None, n, H, A, L.stream())
    wq = w.to(torch.bfloat16).float()
    assert torch.allclose(logits, h.float() @ wq.t() + b, rtol=1e-4, atol=1e-4)
    dl = torch.randn(n, A, generator=g).to(DEV)
    dh = torch.empty(n, H, dtype=torch.bfloat16, device=DEV)
    dw, db = torch.full((A, H), 3.0, device=DEV), torch.full((A,), 3.0, device=DEV)
    L.call("cb_policy_heads_bwd", L.ptr(h), L.ptr(dl), None, L.ptr(w), None, L.ptr(dh), L.ptr(dw), L.ptr(db), None, None,
           n, H, A, L.ACT_RELU, L.stream())
    ref_dh = (dl @ wq) * (h.float() > 0)
    assert torch.allclose(dh.float(), ref_dh, rtol=1e-2, atol=1e-2)
    assert torch.allclose(dw, dl.t() @ h.float(), rtol=1e-3, atol=1e-3) and torch.allclose(db, dl.sum(0), rtol=1e-4, atol=1e-3)


def test_generic_heads_path_for_large_action_sets():
    """A = 9 (> 7): the heads run through the generic Linear layers; same parity gate against the oracle."""
    T, B, A, seed = 6, 8, 9, 77
    params = R.make_policy_params(A, seed)
    icm_params = R.make_params(R.icm_shapes(A, "idf_burda"), seed + 2)
    m = make_model(A, params, icm_params)
    assert not m._fused_heads()
    obs = R.make_frames(seed + 3, T, B)
    prev_action, action = R.make_actions(seed + 5, T, B, A), R.make_actions(seed + 6, T, B, A)
    old_prob, ret, adv = R.make_probs(seed + 7, T, B, A), R.make_normal(seed + 8, T, B), R.make_normal(seed + 9, T, B)
    valid = 1 - R.make_suffix_done(seed + 10, T, B, frac=0.3)
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    pi, v, _ = O.policy_forward(p, obs, R.onehot(prev_action, A), None)
    ref = O.ppo_loss(pi, v, action, old_prob, ret, adv, valid, 0.1, 1.0, 0.01)
    ref[0].backward()
    algo = PPO(curiosity_type="none", clip_grad_norm=1e9)
    algo.initialize(m, n_itr=10)
    dv = lambda t: t.to(DEV)
    info = algo.loss_and_grads_(dv(obs), dv(prev_action), dv(action), dv(ret), dv(adv), dv(valid), dv(old_prob), None)
    assert abs(float(info.loss) - float(ref[0])) <= 1e-2 * abs(float(ref[0]))
    named = dict(m.named_parameters())
    for k in ("pi.weight", "pi.bias", "value.weight", "value.bias", "lstm.weight_hh_l0", "lstm.weight_ih_l0"):
        assert cosine(named[k].grad.cpu(), p[k].grad) > 0.995, (k, cosine(named[k].grad.cpu(), p[k].grad))


def test_lstm_gemms_run_on_tcgen05():
    """The three GEMM shapes of the LSTM (x W_ih^T with the one-hot k-block, h W_hh^T, their dgrad / wgrad) have
    tcgen05 specialisations: forcing the engine must not raise."""
    m = AtariLstmModel((4, 84, 84), 4, curiosity_kwargs=ICM_KW).to(DEV)
    ex = m._build()
    n, B, H = 256, 32, 512
    feat = torch.randn(n, 512, device=DEV).to(torch.bfloat16)
    side = R.onehot(R.make_actions(1, n, 1, 4), 4).reshape(n, 4).to(DEV)
    LY.set_engine("umma")
    try:
        _, gx = ex["ih"].forward(feat, n, side=side, side_pad=LY.pad_side(side), out_bf16=False, out_f32=True)
        _, gh = ex["hh"].forward(feat[:B].contiguous(), B, out_bf16=False, out_f32=True)
        dg = torch.randn(n, 4 * H, device=DEV).to(torch.bfloat16)
        ex["hh"].dgrad(dg[:B].contiguous(), B, dx_add=feat[:B].contiguous())
        ex["ih"].dgrad(dg, n)
        ex["hh"].wgrad(feat, dg, n)
        ex["ih"].wgrad(feat, dg, n, side=side)
    finally:
        LY.set_engine("auto")
    ls = m.lstm
    ref = feat.float() @ ls.weight_ih_l0[:, :512].t().to(torch.bfloat16).float() + side @ ls.weight_ih_l0[:, 512:].t() + ls.bias_ih_l0
    assert rel_err(gx.cpu(), ref.detach().cpu()) < 1e-2
    ref_dw = dg.float().t() @ feat.float()
    assert rel_err(ls.weight_hh_l0.grad.cpu(), ref_dw.cpu()) < 1e-2
    assert rel_err(ls.bias_hh_l0.grad.cpu(), dg.float().sum(0).cpu()) < 1e-2


def test_policy_and_ppo_loss_golden():
    g = golden("policy_ppo_icm")
    c = policy_case(g)
    T, B, A = c["T"], c["B"], c["A"]
    m = make_model(A, c["params"], c["icm_params"])
    dv = lambda t: t.to(DEV)
    with torch.no_grad():
        pi, v, rnn = m(dv(c["obs"]), dv(R.onehot(c["prev_action"], A)), torch.zeros(T, B), (dv(c["h0"]), dv(c["c0"])))
    assert rel_err(pi.cpu(), g["pi"]) < 1e-2 and rel_err(v.cpu(), g["value"]) < 1e-2
    assert rel_err(rnn.h.cpu(), g["hn"]) < 1e-2 and rel_err(rnn.c.cpu(), g["cn"]) < 1e-2
    algo = PPO(curiosity_type="icm", ratio_clip=float(g["ratio_clip"]), value_loss_coeff=float(g["value_loss_coeff"]),
               entropy_loss_coeff=float(g["entropy_loss_coeff"]), clip_grad_norm=1e9)
    algo.initialize(m, n_itr=10)
    h0b, c0b = dv(c["h0"].transpose(0, 1).contiguous()), dv(c["c0"].transpose(0, 1).contiguous())     # [B,1,H]
    cur = (dv(c["obs"]), dv(c["nxt"]), dv(R.onehot(c["action"], A)))
    info = algo.loss_and_grads_(dv(c["obs"]), dv(c["prev_action"]), dv(c["action"]), dv(c["ret"]), dv(c["adv"]),
                                dv(c["valid"]), dv(c["old_prob"]), (h0b, c0b), cur, None)
    for name, val in [("pi_loss", info.pi_loss), ("value_loss", info.value_loss), ("entropy_loss", info.entropy_loss),
                      ("entropy", info.entropy), ("perplexity", info.perplexity), ("inv_loss", info.inv_loss),
                      ("fwd_loss", info.forward_loss), ("loss", info.loss)]:
        ref = float(g[name])
        assert abs(float(val) - ref) <= 1e-2 * max(abs(ref), 1e-2), (name, float(val), ref)
    grads = {k: p.grad for k, p in m.named_parameters()}
    assert all(v is not None for v in grads.values())
    bad = check_grads(grads, g, cos_min=0.99, norm_tol=4e-2)      # 18 samples: same allowance as the ICM golden case
    assert not bad, bad
    gn = float(torch.sqrt(sum((p.grad.double() ** 2).sum() for p in m.parameters())))
    assert abs(gn - float(g["grad_norm"])) / float(g["grad_norm"]) < 5e-2


def test_policy_gradients_vs_oracle_larger_batch():
    """T=12, B=16 (192 samples): per-parameter cosine >= 0.999 and norm within 1e-2 of the oracle's autograd
    (SURVEY 8d parity gate) for the policy parameters under the PPO loss alone."""
    T, B, A, seed = 12, 16, 4, 123
    params = R.make_policy_params(A, seed)
    icm_params = R.make_params(R.icm_shapes(A, "idf_burda"), seed + 2)
    m = make_model(A, params, icm_params)
    obs = R.make_frames(seed + 3, T, B)
    prev_action, action = R.make_actions(seed + 5, T, B, A), R.make_actions(seed + 6, T, B, A)
    old_prob, ret, adv = R.make_probs(seed + 7, T, B, A), R.make_normal(seed + 8, T, B), R.make_normal(seed + 9, T, B)
    valid = 1 - R.make_suffix_done(seed + 10, T, B, frac=0.3)
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    pi, v, _ = O.policy_forward(p, obs, R.onehot(prev_action, A), None)
    ref = O.ppo_loss(pi, v, action, old_prob, ret, adv, valid, 0.1, 1.0, 0.01)
    ref[0].backward()
    algo = PPO(curiosity_type="none", clip_grad_norm=1e9)
    algo.initialize(m, n_itr=10)
    dv = lambda t: t.to(DEV)
    info = algo.loss_and_grads_(dv(obs), dv(prev_action), dv(action), dv(ret), dv(adv), dv(valid), dv(old_prob), None)
    assert abs(float(info.loss) - float(ref[0])) <= 1e-2 * abs(float(ref[0]))
    bad = []
    for k, t in p.items():
        gq = dict(m.named_parameters())[k].grad.cpu()
        cs = cosine(gq, t.grad)
        nr = abs(float(gq.norm()) - float(t.grad.norm())) / float(t.grad.norm())
        # the encoder's gradient crosses T bf16 recurrent dgrad steps before its own bf16 conv dgrads: 0.998 / 2e-2 there
        conv = k.startswith("conv.")
        if cs < (0.998 if conv else 0.999) or nr > (2e-2 if conv else 1e-2):
            bad.append((k, cs, nr))
    assert not bad, bad


def test_autograd_dropin_matches_fused_path():
    """model.forward under autograd + the reference's loss written in torch ops (what rlpyt's PPO.loss does after
    integration.install()) gives the same gradients as the fused cb_ppo_loss path."""
    T, B, A, seed = 5, 4, 4, 321
    params = R.make_policy_params(A, seed)
    icm_params = R.make_params(R.icm_shapes(A, "idf_burda"), seed + 2)
    obs = R.make_frames(seed + 3, T, B).to(DEV)
    prev_action, action = R.make_actions(seed + 5, T, B, A).to(DEV), R.make_actions(seed + 6, T, B, A).to(DEV)
    old_prob, ret, adv = R.make_probs(seed + 7, T, B, A).to(DEV), R.make_normal(seed + 8, T, B).to(DEV), R.make_normal(seed + 9, T, B).to(DEV)
    valid = (1 - R.make_suffix_done(seed + 10, T, B, frac=0.3)).to(DEV)
    m1 = make_model(A, params, icm_params)
    pi, v, _ = m1(obs, R.onehot(prev_action.cpu(), A).to(DEV), None, None)
    loss = O.ppo_loss(pi, v, action, old_prob, ret, adv, valid, 0.1, 1.0, 0.01)[0]      # torch ops on CUDA tensors
    loss.backward()
    m2 = make_model(A, params, icm_params)
    algo = PPO(curiosity_type="none", clip_grad_norm=1e9)
    algo.initialize(m2, n_itr=10)
    info = algo.loss_and_grads_(obs, prev_action, action, ret, adv, valid, old_prob, None)
    assert abs(float(info.loss) - float(loss)) <= 1e-4 * abs(float(loss))
    g2 = dict(m2.named_parameters())
    for k, prm in m1.named_parameters():
        if k.startswith("curiosity_model."):
            continue
        assert cosine(prm.grad.cpu(), g2[k].grad.cpu()) > 0.9999, k


def test_ppo_optimize_agent_two_iterations_vs_oracle():
    """Full iteration (ICM bonus -> returns -> epochs x minibatches of policy + ICM loss -> one clip + Adam over
    all parameters -> LR / ratio-clip decay) against the oracle trained with torch.optim.Adam on the same
    minibatch order."""
    T, B, A, seed = 8, 8, 4, 555
    params = R.make_policy_params(A, seed)
    icm_params = R.make_params(R.icm_shapes(A, "idf_burda"), seed + 2)
    m = make_model(A, params, icm_params)
    LR = 2e-5          # small enough that Adam's sign-like first steps do not amplify bf16 noise into the next iteration
    algo = PPO(curiosity_type="icm", learning_rate=LR, gae_lambda=0.95, minibatches=2, epochs=2)
    algo.initialize(m, n_itr=4, rng=np.random.RandomState(3))
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    q = {k: v.clone().requires_grad_(True) for k, v in icm_params.items()}
    allp = list(p.values()) + list(q.values())
    opt = torch.optim.Adam(allp, lr=LR)
    rng = np.random.RandomState(3)
    ratio_clip, lr = 0.1, LR
    dv = lambda t: t.to(DEV)
    for itr in range(2):
        s = seed + 100 * itr
        obs, nxt = R.make_frames(s + 3, T, B), R.make_frames(s + 4, T, B)
        prev_action, action = R.make_actions(s + 5, T, B, A), R.make_actions(s + 6, T, B, A)
        old_prob = R.make_probs(s + 7, T, B, A)
        done = R.make_suffix_done(s + 10, T, B, frac=0.3)
        value, boot = R.make_normal(s + 11, T, B), R.make_normal(s + 12, B)
        h0, c0 = R.make_normal(s + 13, B, 1, 512) * 0.5, R.make_normal(s + 14, B, 1, 512) * 0.5
        batch = PpoBatch(dv(obs), dv(nxt), dv(prev_action), dv(action), dv(old_prob), torch.zeros(T, B, device=DEV),
                         dv(done), dv(value), dv(boot), (dv(h0), dv(c0)))
        infos = algo.optimize_agent(itr, batch, dropout_masks=None)
        # ---- oracle
        with torch.no_grad():
            r_int = O.icm_bonus(q, obs, nxt, R.onehot(action, A), "idf_burda", 1.0)
        adv, ret = O.generalized_advantage_estimation(r_int.clone(), value, done, boot, 0.99, 0.95)
        valid = O.valid_from_done(done)
        for g_ in opt.param_groups:
            g_["lr"] = lr
        k = 0
        for _ in range(2):
            idx = np.arange(B)
            rng.shuffle(idx)
            for start in range(0, B - B // 2 + 1, B // 2):
                i = torch.as_tensor(idx[start:start + B // 2])
                opt.zero_grad()
                pi, v, _ = O.policy_forward(p, obs[:, i], R.onehot(prev_action[:, i], A),
                                            (h0[i].transpose(0, 1), c0[i].transpose(0, 1)))
                pl = O.ppo_loss(pi, v, action[:, i], old_prob[:, i], ret[:, i], adv[:, i], valid[:, i], ratio_clip, 1.0, 0.01)
                inv, fwd = O.icm_loss(q, obs[:, i], nxt[:, i], R.onehot(action[:, i], A), valid[:, i], "idf_burda", 0.2)
                total = pl[0] + inv + fwd
                total.backward()
                gn = float(torch.nn.utils.clip_grad_norm_(allp, 1.0))
                opt.step()
                for ours, ref_ in [(infos[k].loss, total), (infos[k].value_loss, pl[2]), (infos[k].inv_loss, inv),
                                   (infos[k].forward_loss, fwd), (infos[k].entropy, pl[4])]:
                    assert abs(float(ours) - float(ref_)) <= 5e-2 * max(abs(float(ref_)), 1e-2), (itr, k, float(ours), float(ref_))
                assert abs(float(infos[k].gradNorm) - gn) / gn < 1e-1, (itr, k, float(infos[k].gradNorm), gn)
                k += 1
        assert k == len(infos) == 4
        assert all(layer._version is None for layer in m._exec["chain"].layers + [m._exec["ih"], m._exec["hh"]])
        lr = LR * (4 - (itr + 1)) / 4
        ratio_clip = 0.1 * (4 - itr) / 4
        assert abs(algo.opt.lr - lr) < 1e-12 and abs(algo.ratio_clip - ratio_clip) < 1e-12
    named = dict(m.named_parameters())
    for k_, t in list(p.items()) + [("curiosity_model." + a, b) for a, b in q.items()]:
        init = params[k_] if k_ in params else icm_params[k_[len("curiosity_model."):]]
        d_ours = (named[k_].detach().cpu() - init).flatten()
        d_ref = (t.detach() - init).flatten()
        assert cosine(d_ours, d_ref) > 0.8, (k_, cosine(d_ours, d_ref))


def test_batch_prefetcher_delivers_identical_tensors():
    from curiosity_baselines_b200.agents.hooks import BatchPrefetcher
    pre = BatchPrefetcher(DEV)
    host = [R.make_frames(1, 3, 2).pin_memory(), R.make_actions(2, 3, 2, 4).pin_memory(), R.make_normal(3, 3, 2).pin_memory()]
    for _ in range(3):
        pre.prefetch(host)
        got = pre.get()
        torch.cuda.synchronize()
        assert all(torch.equal(a.cpu(), b) for a, b in zip(got, host))


def test_forward_leading_dims_like_the_reference():
    """atari_lstm_model.py:117-154 accepts [T,B], [B] and [] leading dims (the sampler calls it with [B] every env
    step); outputs keep them, the rnn state stays [1,B,H]."""
    A, seed, Bn = 4, 808, 5
    params = R.make_policy_params(A, seed)
    m = make_model(A, params, R.make_params(R.icm_shapes(A, "idf_burda"), seed + 2))
    obs = R.make_frames(seed + 1, 1, Bn)                      # [1,B,4,84,84]
    prev = R.onehot(R.make_actions(seed + 2, 1, Bn, A), A)
    h0, c0 = R.make_normal(seed + 3, 1, Bn, 512) * 0.3, R.make_normal(seed + 4, 1, Bn, 512) * 0.3
    with torch.no_grad():
        pi_ref, v_ref, (hn_ref, cn_ref) = O.policy_forward(params, obs, prev, (h0, c0))
        pi, v, rnn = m(obs[0].to(DEV), prev[0].to(DEV), torch.zeros(Bn), (h0.to(DEV), c0.to(DEV)))      # [B] leading dim
        assert pi.shape == (Bn, A) and v.shape == (Bn,) and rnn.h.shape == (1, Bn, 512)
        assert rel_err(pi.cpu(), pi_ref[0]) < 1e-2 and rel_err(v.cpu(), v_ref[0]) < 1e-2
        assert rel_err(rnn.h.cpu(), hn_ref) < 1e-2 and rel_err(rnn.c.cpu(), cn_ref) < 1e-2
        # feeding the returned state back in continues the sequence (what the collector does step by step)
        pi2, v2, rnn2 = m(obs[0].to(DEV), prev[0].to(DEV), torch.zeros(Bn), rnn)
        _, v2_ref, _ = O.policy_forward(params, obs, prev, (hn_ref, cn_ref))
        assert rel_err(v2.cpu(), v2_ref[0]) < 1e-2
        # no leading dims at all: one observation, zero state
        pi1, v1, rnn1 = m(obs[0, 0].to(DEV), prev[0, 0].to(DEV), torch.zeros(()), None)
        pi1_ref, v1_ref, _ = O.policy_forward(params, obs[:, :1], prev[:, :1], None)
        assert pi1.shape == (A,) and v1.shape == () and rnn1.h.shape == (1, 1, 512)
        assert rel_err(pi1.cpu(), pi1_ref[0, 0]) < 1e-2 and abs(float(v1) - float(v1_ref[0, 0])) < 1e-2 * max(1.0, abs(float(v1_ref[0, 0])))


def test_cpu_inputs_and_cpu_model_are_rejected():
    m = AtariLstmModel((4, 84, 84), 4, curiosity_kwargs=ICM_KW)           # parameters on the CPU
    with pytest.raises(L.CuriosityLibError):
        m(R.make_frames(1, 2, 2), R.onehot(R.make_actions(2, 2, 2, 4), 4), torch.zeros(2, 2), None)


def test_twin_first_conv_gives_the_same_losses_and_gradients():
    """The policy's conv1 and ICM's conv1 over `observation` as ONE launch (PPO.twin_conv1) vs two separate launches."""
    T, B, A, seed = 4, 6, 4, 4321
    params = R.make_policy_params(A, seed)
    icm_params = R.make_params(R.icm_shapes(A, "idf_burda"), seed + 2)
    dv = lambda t: t.to(DEV)
    obs, nxt = dv(R.make_frames(seed + 3, T, B)), dv(R.make_frames(seed + 4, T, B))
    prev, act = dv(R.make_actions(seed + 5, T, B, A)), dv(R.make_actions(seed + 6, T, B, A))
    old_prob, ret, adv = dv(R.make_probs(seed + 7, T, B, A)), dv(R.make_normal(seed + 8, T, B)), dv(R.make_normal(seed + 9, T, B))
    valid = dv(1 - R.make_suffix_done(seed + 10, T, B, frac=0.3))
    results = []
    for twin in (True, False):
        m = make_model(A, params, icm_params)
        algo = PPO(curiosity_type="icm", clip_grad_norm=1e9)
        algo.twin_conv1 = twin
        algo.initialize(m, n_itr=10)
        calls0 = L.launch_count()
        info = algo.loss_and_grads_(obs, prev, act, ret, adv, valid, old_prob, None, (obs, nxt, dv(R.onehot(act.cpu(), A))), None)
        results.append((info, {k: p.grad.clone() for k, p in m.named_parameters()}, L.launch_count() - calls0))
        assert "_first_out_obs" not in m.curiosity_model.__dict__            # consumed
    (ia, ga, ca), (ib, gb, cb_) = results
    assert ca == cb_ - 1                       # the twin launch replaces two first-conv launches
    for name in ("loss", "pi_loss", "value_loss", "inv_loss", "forward_loss"):
        # not bit-identical: the N = 64 twin MMA and the N = 32 single one round a few bf16 activations differently
        # (observed 2e-5 relative on the inverse loss); both are far inside the 1e-2 parity gate
        assert abs(float(getattr(ia, name)) - float(getattr(ib, name))) <= 1e-3 * max(1.0, abs(float(getattr(ib, name)))), name
    for k in ga:
        assert cosine(ga[k].cpu(), gb[k].cpu()) > 0.999, k


def test_ppo_with_disagreement_curiosity_vs_oracle():
    """PPO.loss with the Disagreement terms (ppo.py:226-229): policy + inverse + four-model ensemble loss in one
    fused step, twin first conv included; dropout masks off on both sides."""
    from curiosity_baselines_b200.curiosity.icm import Disagreement
    T, B, A, seed = 6, 5, 7, 999
    dis_kw = dict(ICM_KW, curiosity_alg="disagreement", device="cuda")
    params = R.make_policy_params(A, seed)
    dparams = R.make_params(R.disagreement_shapes(A, "idf_burda"), seed + 2)
    m = AtariLstmModel((4, 84, 84), A, curiosity_kwargs=dis_kw).to(DEV)
    assert isinstance(m.curiosity_model, Disagreement)
    sd = dict(params)
    sd.update({"curiosity_model." + k: v for k, v in dparams.items()})
    m.load_state_dict(sd)
    obs, nxt = R.make_frames(seed + 3, T, B), R.make_frames(seed + 4, T, B)
    prev, act = R.make_actions(seed + 5, T, B, A), R.make_actions(seed + 6, T, B, A)
    old_prob, ret, adv = R.make_probs(seed + 7, T, B, A), R.make_normal(seed + 8, T, B), R.make_normal(seed + 9, T, B)
    valid = 1 - R.make_suffix_done(seed + 10, T, B, frac=0.3)
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    q = {k: v.clone().requires_grad_(True) for k, v in dparams.items()}
    pi, v, _ = O.policy_forward(p, obs, R.onehot(prev, A), None)
    pl = O.ppo_loss(pi, v, act, old_prob, ret, adv, valid, 0.1, 1.0, 0.01)
    inv, fwd = O.disagreement_loss(q, obs, nxt, R.onehot(act, A), valid, "idf_burda", 0.2, None)
    total = pl[0] + inv + fwd
    total.backward()
    algo = PPO(curiosity_type="disagreement", clip_grad_norm=1e9)
    algo.initialize(m, n_itr=10)
    dv = lambda t: t.to(DEV)
    o = dv(obs)
    info = algo.loss_and_grads_(o, dv(prev), dv(act), dv(ret), dv(adv), dv(valid), dv(old_prob), None,
                                (o, dv(nxt), dv(R.onehot(act, A))), None)
    for ours, ref_ in [(info.loss, total), (info.inv_loss, inv), (info.forward_loss, fwd), (info.pi_loss, pl[1])]:
        assert abs(float(ours) - float(ref_)) <= 1e-2 * max(abs(float(ref_)), 1e-2), (float(ours), float(ref_))
    named = dict(m.named_parameters())
    gn_ref = float(torch.sqrt(sum((t.grad.double() ** 2).sum() for t in list(p.values()) + list(q.values()))))
    gn = float(torch.sqrt(sum((t.grad.double() ** 2).sum() for t in named.values())))
    assert abs(gn - gn_ref) / gn_ref < 5e-2
    for k in ("lstm.weight_hh_l0", "pi.weight", "curiosity_model.inverse_model.2.weight",
              "curiosity_model.forward_model_3.lin_last.weight"):
        ref_g = p[k].grad if k in p else q[k[len("curiosity_model."):]].grad
        assert cosine(named[k].grad.cpu(), ref_g) > 0.99, k


@pytest.mark.parametrize("enc,shape,gain", [("idf", (4, 84, 84), 0.5), ("idf_maze", (3, 5, 5), 1.0)])
def test_policy_with_the_other_curiosity_encoders(enc, shape, gain):
    """atari_lstm_model.py:80-93: the policy's conv is UniverseHead ('idf', 288 features in CHW order) or MazeHead
    ('idf_maze', 256 features, float {0,1} frames); outputs, loss and gradients against the oracle."""
    T, B, A, seed = 5, 6, 5, 2024
    kw = dict(ICM_KW, feature_encoding=enc)
    params = R.make_policy_params(A, seed, feature_encoding=enc, conv_gain=gain, c_in=shape[0])
    m = AtariLstmModel(shape, A, curiosity_kwargs=kw).to(DEV)
    sd = dict(m.state_dict())
    assert {k: tuple(v.shape) for k, v in sd.items() if not k.startswith("curiosity_model.")} == \
        dict(R.policy_shapes(A, enc, c_in=shape[0]))
    sd.update(params)
    m.load_state_dict(sd)
    obs = R.make_frames(seed + 3, T, B) if enc == "idf" else R.make_binary_frames(seed + 3, T, B)
    prev, act = R.make_actions(seed + 5, T, B, A), R.make_actions(seed + 6, T, B, A)
    old_prob, ret, adv = R.make_probs(seed + 7, T, B, A), R.make_normal(seed + 8, T, B), R.make_normal(seed + 9, T, B)
    valid = 1 - R.make_suffix_done(seed + 10, T, B, frac=0.3)
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    pi, v, _ = O.policy_forward(p, obs, R.onehot(prev, A), None, feature_encoding=enc)
    ref = O.ppo_loss(pi, v, act, old_prob, ret, adv, valid, 0.1, 1.0, 0.01)
    ref[0].backward()
    dv = lambda t: t.to(DEV)
    with torch.no_grad():
        pi_g, v_g, _ = m(dv(obs), dv(R.onehot(prev, A)), torch.zeros(T, B), None)
    assert rel_err(pi_g.cpu(), pi.detach()) < 1e-2 and rel_err(v_g.cpu(), v.detach()) < 1e-2
    algo = PPO(curiosity_type="none", clip_grad_norm=1e9)
    algo.initialize(m, n_itr=10)
    info = algo.loss_and_grads_(dv(obs), dv(prev), dv(act), dv(ret), dv(adv), dv(valid), dv(old_prob), None)
    assert abs(float(info.loss) - float(ref[0])) <= 1e-2 * abs(float(ref[0]))
    named = dict(m.named_parameters())
    for k, t in p.items():
        assert cosine(named[k].grad.cpu(), t.grad) > 0.99, (k, cosine(named[k].grad.cpu(), t.grad))


def _lstm_seq(T, B, H, gates, w, w_t, b, h0, c0, dh_out):
    """cb_lstm_sequence_fwd + _bwd on raw buffers -> (h_all, c_all, activated gates, dgates)."""
    gates = gates.clone()
    h_all = torch.empty(T + 1, B, H, dtype=torch.bfloat16, device=DEV)
    c_all = torch.empty(T + 1, B, H, device=DEV)
    h_all[0], c_all[0] = h0, c0
    gh, hn = torch.empty(B, 4 * H, device=DEV), torch.empty(B, H, device=DEV)
    L.call("cb_lstm_sequence_fwd", T, B, H, L.ptr(gates), L.ptr(w), L.ptr(b), L.ptr(h_all), L.ptr(c_all), L.ptr(gh), L.ptr(hn),
           L.ENGINE_AUTO, L.stream())
    dgates = torch.empty(T, B, 4 * H, dtype=torch.bfloat16, device=DEV)
    dc, dh = torch.empty(2, B, H, device=DEV), torch.empty(2, B, H, dtype=torch.bfloat16, device=DEV)
    L.call("cb_lstm_sequence_bwd", T, B, H, L.ptr(gates), L.ptr(c_all), L.ptr(dh_out.contiguous()), L.ptr(w_t), L.ptr(dgates),
           L.ptr(dc), L.ptr(dh), L.ENGINE_AUTO, L.stream())
    torch.cuda.synchronize()
    assert torch.equal(hn.to(torch.bfloat16), h_all[T])
    return h_all, c_all, gates, dgates


def test_recurrence_rows_are_independent_at_large_batch():
    """cb_lstm_sequence_* at B = 512 equals running each half of the env rows on its own, bit for bit (rows are
    independent and the kernels identical), and matches the fp32 recurrence in torch.  (Splitting the rows over two
    streams so that one half's GEMM overlaps the other half's cell kernel was tried and measured no faster: 3.92 vs
    3.96 ms forward, 5.77 vs 5.89 ms backward at T=128, B=2048 -- the step is bound by its fp32 gate traffic.)"""
    T, B, H = 6, 512, 512
    g = torch.Generator(device="cpu").manual_seed(31)
    gates = torch.randn(T, B, 4 * H, generator=g).to(DEV)
    w = (torch.randn(4 * H, H, generator=g) * 0.04).to(DEV).to(torch.bfloat16)
    w_t = w.t().contiguous()
    b = (torch.randn(4 * H, generator=g) * 0.1).to(DEV)
    h0 = (torch.randn(B, H, generator=g) * 0.3).to(DEV).to(torch.bfloat16)
    c0 = (torch.randn(B, H, generator=g) * 0.3).to(DEV)
    dh_out = (torch.randn(T, B, H, generator=g) * 0.1).to(DEV).to(torch.bfloat16)
    full = _lstm_seq(T, B, H, gates, w, w_t, b, h0, c0, dh_out)
    for half in (slice(0, B // 2), slice(B // 2, B)):
        part = _lstm_seq(T, B // 2, H, gates[:, half].contiguous(), w, w_t, b, h0[half].contiguous(), c0[half].contiguous(),
                         dh_out[:, half].contiguous())
        for a, p_ in zip(full, part):
            assert torch.equal(a[:, half], p_)
    # and against the fp32 recurrence in torch (bf16 GEMM operands: 1e-2)
    h, c = h0.float(), c0
    for t in range(T):
        a = gates[t] + h.to(torch.bfloat16).float() @ w.float().t() + b
        i, f, gg, o = a.chunk(4, 1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        assert rel_err(full[1][t + 1].cpu(), c.cpu()) < 1e-2 and rel_err(full[0][t + 1].float().cpu(), h.cpu()) < 2e-2


def test_ppo_icm_on_a_collector_layout_rollout_matches_separate_tensors():
    """PPO + ICM when next_observation is observation shifted by one step (two views of one T+1 buffer, what
    BatchStaging.get_shifted builds from rlpyt's sampler buffers): the curiosity encoder runs once over T+1 steps and its
    first convolution rides the policy's conv1 launch (twin, N = 64 MMAs over one converted image).  Every loss term equals
    the run on two separate tensors with the same content to 1e-4 (the twin and the single-network first convolution add
    their bias in different places of the fp32 pipeline); gradients agree up to summation order."""
    T, B, A, seed = 6, 4, 4, 4242
    params = R.make_policy_params(A, seed)
    icm_params = R.make_params(R.icm_shapes(A, "idf_burda"), seed + 2)
    frames = R.make_frames(seed + 3, T + 1, B).to(DEV)
    prev, act = R.make_actions(seed + 5, T, B, A).to(DEV), R.make_actions(seed + 6, T, B, A).to(DEV)
    old_prob, ret, adv = R.make_probs(seed + 7, T, B, A).to(DEV), R.make_normal(seed + 8, T, B).to(DEV), R.make_normal(seed + 9, T, B).to(DEV)
    valid = (1 - R.make_suffix_done(seed + 10, T, B, frac=0.3)).to(DEV)

    def run(shifted):
        m = make_model(A, params, icm_params)
        obs, nxt = (frames[:T], frames[1:]) if shifted else (frames[:T].clone(), frames[1:].clone())
        algo = PPO(curiosity_type="icm", clip_grad_norm=1e9)
        algo.initialize(m, n_itr=10)
        info = algo.loss_and_grads_(obs, prev, act, ret, adv, valid, old_prob, None, (obs, nxt, R.onehot(act.cpu(), A).to(DEV)), None)
        return m, info, {k: p.grad.clone() for k, p in m.named_parameters()}

    m0, i0, g0 = run(False)
    m1, i1, g1 = run(True)
    assert getattr(m1.curiosity_model, "dedup_hits", 0) == 1 and getattr(m0.curiosity_model, "dedup_hits", 0) == 0
    for f in ("loss", "pi_loss", "value_loss", "inv_loss", "forward_loss"):
        assert abs(float(getattr(i0, f)) - float(getattr(i1, f))) <= 1e-4 * max(1.0, abs(float(getattr(i0, f)))), f
    for k in g0:
        a, b = g0[k].double().flatten(), g1[k].double().flatten()
        assert float(a @ b / (a.norm() * b.norm())) > 0.9999, k


def test_twelve_actions_run_entirely_on_tcgen05():
    """SuperMarioBros' COMPLEX_MOVEMENT has 12 actions (envs/gym.py:190): too wide for the fused heads kernel.  The policy's
    pi / value heads and ICM's inverse-model output then run as zero-padded tcgen05 GEMMs (layers.NarrowHead), not on
    the SIMT cross-check engine: with the engine forced to 'umma' (which raises on any shape without a tcgen05
    specialisation) a full PPO + ICM loss / gradient pass goes through and matches the oracle."""
    T, B, A, seed = 6, 8, 12, 1212
    params = R.make_policy_params(A, seed)
    icm_params = R.make_params(R.icm_shapes(A, "idf_burda"), seed + 2)
    obs, nxt = R.make_frames(seed + 3, T, B), R.make_frames(seed + 4, T, B)
    prev_action, action = R.make_actions(seed + 5, T, B, A), R.make_actions(seed + 6, T, B, A)
    old_prob, ret, adv = R.make_probs(seed + 7, T, B, A), R.make_normal(seed + 8, T, B), R.make_normal(seed + 9, T, B)
    valid = 1 - R.make_suffix_done(seed + 10, T, B, frac=0.3)
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    q = {k: v.clone().requires_grad_(True) for k, v in icm_params.items()}
    pi, v, _ = O.policy_forward(p, obs, R.onehot(prev_action, A), None)
    ref = O.ppo_loss(pi, v, action, old_prob, ret, adv, valid, 0.1, 1.0, 0.01)
    inv, fwd = O.icm_loss(q, obs, nxt, R.onehot(action, A), valid, "idf_burda", 0.2)
    (ref[0] + inv + fwd).backward()
    LY.set_engine("umma")
    try:
        m = make_model(A, params, icm_params)
        assert not m._fused_heads()
        algo = PPO(curiosity_type="icm", clip_grad_norm=1e9)
        algo.initialize(m, n_itr=10)
        dv = lambda t: t.to(DEV)
        o = dv(obs)
        info = algo.loss_and_grads_(o, dv(prev_action), dv(action), dv(ret), dv(adv), dv(valid), dv(old_prob), None,
                                    (o, dv(nxt), dv(R.onehot(action, A))), None)
    finally:
        LY.set_engine("auto")
    assert abs(float(info.loss) - float(ref[0] + inv + fwd)) <= 1e-2 * abs(float(ref[0] + inv + fwd))
    assert abs(float(info.inv_loss) - float(inv)) <= 1e-2 * float(inv)
    named = dict(m.named_parameters())
    for k in ("pi.weight", "pi.bias", "value.weight", "value.bias", "lstm.weight_hh_l0"):
        assert cosine(named[k].grad.cpu(), p[k].grad) > 0.995, (k, cosine(named[k].grad.cpu(), p[k].grad))
    for k in ("inverse_model.2.weight", "inverse_model.2.bias", "inverse_model.0.weight"):
        assert cosine(named["curiosity_model." + k].grad.cpu(), q[k].grad) > 0.995, k
