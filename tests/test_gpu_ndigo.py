"""NDIGO on the GPU vs the reference's golden vectors (bonus, loss, gradients)."""
import pytest
import torch

import recipes as R
from helpers import golden, check_grads
from curiosity_baselines_b200 import _lib as L
from curiosity_baselines_b200.curiosity.ndigo import NDIGO, _GruFn

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_state_dict_keys_match_reference():
    m = NDIGO(image_shape=(3, 5, 5), action_size=5, horizon=1)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == dict(R.ndigo_shapes(5))


@pytest.mark.parametrize("H", [1, 3])
def test_golden(H):
    g = golden(f"ndigo_h{H}")
    T, B, A, seed = int(g["T"]), int(g["B"]), int(g["A"]), int(g["seed"])
    m = NDIGO(image_shape=(3, 5, 5), action_size=A, horizon=H, feature_encoding="idf_maze", num_predictors=10).to(DEV)
    m.load_state_dict(R.make_params(R.ndigo_shapes(A), seed))
    obs = R.make_binary_frames(seed + 1, T, B).to(DEV)
    prev = R.onehot(R.make_actions(seed + 2, T, B, A), A).to(DEV)
    act = R.onehot(R.make_actions(seed + 3, T, B, A), A).to(DEV)
    bonus = m.compute_bonus(obs, prev, act).cpu()
    ref = torch.from_numpy(g["bonus"])
    # the bonus is a difference of two BCE means of order 0.7: absolute tolerance 1e-2 of that scale
    assert float((bonus - ref).abs().max()) < 7e-3
    assert torch.equal(bonus == 0, ref == 0)
    loss = m.compute_loss(obs, prev, act, torch.ones(T, B, device=DEV))
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) / float(g["loss"]) < 1e-2
    grads = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    assert len(grads) == len(list(m.parameters()))
    bad = check_grads(grads, g, cos_min=0.99, norm_tol=4e-2)      # few samples; strict gate: test_gpu_parity_scale.py
    assert not bad, bad


def test_horizon_ten_fails_like_the_reference():
    m = NDIGO(image_shape=(3, 5, 5), action_size=5, horizon=10).to(DEV)
    obs = R.make_binary_frames(1, 14, 2).to(DEV)
    a = R.onehot(R.make_actions(2, 14, 2, 5), 5).to(DEV)
    with pytest.raises(AttributeError):
        m.compute_bonus(obs, a, a)


def test_gru_sequence_matches_torch_gru_fp32():
    """cb_gru_sequence_fwd/_bwd (through NDIGO's _GruFn) against torch.nn.GRU in fp32 on the same device: outputs within
    1e-2 (bf16 GEMM operands), parameter / input gradients cosine >= 0.999."""
    T, B, A, F, G, seed = 9, 33, 5, 256, 128, 5
    m = NDIGO(image_shape=(3, 5, 5), action_size=A, horizon=1).to(DEV)
    m._build()
    gen = torch.Generator(device="cpu").manual_seed(seed)
    phi = torch.randn(T * B, F, generator=gen).to(DEV).requires_grad_(True)
    prev = R.onehot(R.make_actions(seed, T, B, A), A).to(DEV)
    dout = torch.randn(T, B, G, generator=gen).to(DEV)
    ref = torch.nn.GRU(F + A, G).to(DEV)
    ref.load_state_dict(m.gru.state_dict())
    x = torch.cat([phi.detach().reshape(T, B, F), prev], 2).requires_grad_(True)
    with torch.backends.cudnn.flags(enabled=False):
        out_ref, _ = ref(x, None)
    (out_ref * dout).sum().backward()
    g = m.gru
    out = _GruFn.apply(m, phi, prev, g.weight_ih_l0, g.weight_hh_l0, g.bias_ih_l0, g.bias_hh_l0)
    (out * dout).sum().backward()
    assert float((out - out_ref).abs().max()) < 1e-2
    cos = lambda a, b: float((a.flatten() @ b.flatten()) / (a.norm() * b.norm()))
    assert cos(phi.grad, x.grad[:, :, :F].reshape(T * B, F)) > 0.999
    for name in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"):
        a, b = getattr(g, name).grad, getattr(ref, name).grad
        assert cos(a, b) > 0.999 and abs(float(a.norm()) - float(b.norm())) / float(b.norm()) < 1e-2, name


def test_gru_cell_kernels_match_torch_fp32():
    B, G = 19, 128
    gen = torch.Generator(device="cpu").manual_seed(11)
    gi, gh = torch.randn(B, 3 * G, generator=gen).to(DEV), torch.randn(B, 3 * G, generator=gen).to(DEV)
    hp = torch.randn(B, G, generator=gen).to(DEV)
    gi_r, gh_r, hp_r = gi.clone().requires_grad_(True), gh.clone().requires_grad_(True), hp.clone().requires_grad_(True)
    ir, iz, inn = gi_r.chunk(3, 1)
    hr, hz, hn = gh_r.chunk(3, 1)
    r, z = torch.sigmoid(ir + hr), torch.sigmoid(iz + hz)
    n = torch.tanh(inn + r * hn)
    h_ref = (1 - z) * n + z * hp_r
    gates, ghn, h = torch.empty(B, 3 * G, device=DEV), torch.empty(B, G, device=DEV), torch.empty(B, G, device=DEV)
    hb = torch.empty(B, G, dtype=torch.bfloat16, device=DEV)
    L.call("cb_gru_cell_fwd", L.ptr(gi), L.ptr(gh), L.ptr(hp), L.ptr(gates), L.ptr(ghn), L.ptr(h), L.ptr(hb), B, G, L.stream())
    assert torch.allclose(h, h_ref, rtol=1e-5, atol=1e-6) and torch.equal(hb, h.to(torch.bfloat16))
    assert torch.allclose(gates, torch.cat([r, z, n], 1), rtol=1e-5, atol=1e-6) and torch.equal(ghn, gh[:, 2 * G:])
    dout = torch.randn(B, G, generator=gen).to(DEV)
    drec = torch.randn(B, G, generator=gen).to(DEV).to(torch.bfloat16)
    (h_ref * (dout + drec.float())).sum().backward()
    dgi, dgh = torch.empty(B, 3 * G, dtype=torch.bfloat16, device=DEV), torch.empty(B, 3 * G, dtype=torch.bfloat16, device=DEV)
    carry = torch.empty(B, G, dtype=torch.bfloat16, device=DEV)
    L.call("cb_gru_cell_bwd", L.ptr(dout), L.ptr(drec), L.ptr(gates), L.ptr(ghn), L.ptr(hp), L.ptr(dgi), L.ptr(dgh), L.ptr(carry),
           B, G, L.stream())
    assert torch.allclose(dgi.float(), gi_r.grad, rtol=1e-2, atol=1e-3)
    assert torch.allclose(dgh.float(), gh_r.grad, rtol=1e-2, atol=1e-3)
    assert torch.allclose(carry.float(), (dout + drec.float()) * z.detach(), rtol=1e-2, atol=1e-3)


def test_gru_sequence_vs_torch_gru_at_large_batch():
    """B = 1024 (several GEMM row tiles per step): same gate as the small-batch test."""
    T, B, A, F, G, seed = 5, 1024, 5, 256, 128, 6
    m = NDIGO(image_shape=(3, 5, 5), action_size=A, horizon=1).to(DEV)
    m._build()
    gen = torch.Generator(device="cpu").manual_seed(seed)
    phi = torch.randn(T * B, F, generator=gen).to(DEV).requires_grad_(True)
    prev = R.onehot(R.make_actions(seed, T, B, A), A).to(DEV)
    dout = torch.randn(T, B, G, generator=gen).to(DEV)
    ref = torch.nn.GRU(F + A, G).to(DEV)
    ref.load_state_dict(m.gru.state_dict())
    x = torch.cat([phi.detach().reshape(T, B, F), prev], 2).requires_grad_(True)
    with torch.backends.cudnn.flags(enabled=False):
        out_ref, _ = ref(x, None)
    (out_ref * dout).sum().backward()
    g = m.gru
    out = _GruFn.apply(m, phi, prev, g.weight_ih_l0, g.weight_hh_l0, g.bias_ih_l0, g.bias_hh_l0)
    (out * dout).sum().backward()
    assert float((out - out_ref).abs().max()) < 1e-2
    cos = lambda a, b: float((a.flatten() @ b.flatten()) / (a.norm() * b.norm()))
    assert cos(phi.grad, x.grad[:, :, :F].reshape(T * B, F)) > 0.999
    for name in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"):
        assert cos(getattr(g, name).grad, getattr(ref, name).grad) > 0.999, name
