"""cb_rnd_conv12_fwd: the first two convolutions of an RND network in one launch (rnd.py:40-46 / 55-61), checked
against the two stand-alone kernels (cb_conv1_fwd + the window conv2) layer by layer, and the whole RND bonus /
loss / gradients with the fusion on against the fusion off and against the oracle."""
import numpy as np
import pytest
import torch

import recipes as R
from curiosity_baselines_b200 import _lib as L
from curiosity_baselines_b200.curiosity.rnd import RND

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _model(seed=0):
    torch.manual_seed(seed)
    m = RND(image_shape=(4, 84, 84)).to(DEV)
    m._build()
    # non-trivial statistics so the normalisation and the +-5 clamp both matter
    g = torch.Generator().manual_seed(seed + 1)
    m._norm[0].copy_((torch.rand(84 * 84, generator=g) * 200 + 20).to(DEV))
    m._norm[1].copy_((torch.rand(84 * 84, generator=g) * 0.05 + 0.005).to(DEV))
    m.fuse_conv12 = True          # (an option: off by default)
    return m


def _frames(T, B, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (T, B, 4, 84, 84), dtype=torch.uint8, generator=g).to(DEV)


@pytest.mark.parametrize("T,B", [(1, 1), (1, 5), (2, 74), (3, 100), (7, 149)])
@pytest.mark.parametrize("which", [0, 1])
def test_fused_layers_equal_the_separate_kernels(T, B, which):
    m = _model()
    obs = _frames(T, B, seed=T * 100 + B)
    obs, T, B, base, strides = m._frame_view(obs)
    n = T * B
    chain = m._chains[which]
    assert m._can_fuse12(base, strides)
    a1_ref, _ = chain.layers[0].forward(base, n, in_dtype=L.U8, strides=strides, norm=m._norm)
    a2_ref, _ = chain.layers[1].forward(a1_ref, n)
    img = m._normalized(base, n, strides[0])
    a1, a2 = m._conv12(chain, img, n, True)
    _, a2_only = m._conv12(chain, img, n, False)
    torch.cuda.synchronize()
    assert torch.equal(a2, a2_only)                   # writing a1 out does not change a2
    a1f, r1 = a1.float(), a1_ref.float()
    # same operands, fp32 accumulation in a different order, one rounding to bf16: at most one bf16 ulp apart, rarely
    d1 = (a1f - r1).abs()
    assert float((d1 / r1.abs().clamp_min(1e-2)).max()) <= 2 ** -6
    assert float((d1 > 0).float().mean()) < 0.02
    d2 = (a2.float() - a2_ref.float()).abs()
    scale = float(a2_ref.float().abs().max())
    assert float(d2.max()) <= 2e-2 * scale
    assert float(d2.mean()) <= 1e-3 * scale
    # conv2 of the fused kernel on ITS OWN a1 against the stand-alone conv2 on the same a1: the same MMAs; the bias is added
    # in fp32 by the epilogue here and through the accumulator there -- at most one bf16 ulp apart, rarely
    a2_chk, _ = chain.layers[1].forward(a1, n)
    # (the negative LeakyReLU branch is rounded once here and twice there: many last-bit differences, nothing larger)
    dd = (a2.float() - a2_chk.float()).abs()
    assert float((dd / a2_chk.float().abs().clamp_min(1e-2)).max()) <= 2 ** -6
    assert float((dd > 0).float().mean()) < 0.4


@pytest.mark.parametrize("T,B", [(1, 1), (1, 5), (2, 74), (3, 100), (7, 149)])
def test_twin_kernel_equals_the_separate_kernels(T, B):
    """cb_rnd_conv12_twin_fwd: target and predictor of every sample in one CTA (shared image, N = 64 first layer)."""
    m = _model()
    obs = _frames(T, B, seed=T * 100 + B + 1)
    obs, T, B, base, strides = m._frame_view(obs)
    n = T * B
    pred, targ = m._chains
    img = m._normalized(base, n, strides[0])
    a2_t, a1_p, a2_p = m._conv12_twin(img, n, True)
    a2_t2, none, a2_p2 = m._conv12_twin(img, n, False)
    torch.cuda.synchronize()
    assert none is None and torch.equal(a2_t, a2_t2) and torch.equal(a2_p, a2_p2)
    # the one-network kernel (predictor alone, feature-reuse path) gives bit-identical results
    a1_s, a2_s = m._conv12(pred, img, n, True)
    assert torch.equal(a1_s, a1_p) and torch.equal(a2_s, a2_p)
    for chain, a1, a2 in ((targ, None, a2_t), (pred, a1_p, a2_p)):
        a1_ref, _ = chain.layers[0].forward(base, n, in_dtype=L.U8, strides=strides, norm=m._norm)
        a2_ref, _ = chain.layers[1].forward(a1_ref, n)
        if a1 is not None:
            d1 = (a1.float() - a1_ref.float()).abs()
            assert float((d1 / a1_ref.float().abs().clamp_min(1e-2)).max()) <= 2 ** -6
            assert float((d1 > 0).float().mean()) < 0.02
        d2 = (a2.float() - a2_ref.float()).abs()
        scale = float(a2_ref.float().abs().max())
        assert float(d2.max()) <= 2e-2 * scale
        assert float(d2.mean()) <= 1e-3 * scale


def test_unaligned_frames_fall_back_to_the_separate_kernels():
    m = _model()
    obs = _frames(2, 3)
    _, _, _, base, strides = m._frame_view(obs)
    assert not m._can_fuse12(base + 2, strides)
    m.fuse_conv12 = False
    assert not m._can_fuse12(base, strides)


def _run(m, obs, done, valid, mask):
    bonus = m.compute_bonus(obs, done)
    for p in m.forward_model.parameters():
        p.grad = None
    loss = m.compute_loss(obs, valid, mask)
    loss.backward()
    grads = torch.cat([p.grad.flatten() for p in m.forward_model.parameters()])
    return bonus, loss.detach(), grads


def test_rnd_step_with_fusion_matches_without():
    T, B = 16, 64
    obs = _frames(T, B, seed=5)
    g = torch.Generator().manual_seed(9)
    done = (torch.rand(T, B, generator=g) < 0.05).float().to(DEV)
    valid = torch.ones(T, B, device=DEV)
    mask = (torch.rand(T, B, generator=g) > 0.25).float().to(DEV)
    out = []
    for fuse in (False, True):
        m = _model(seed=3)
        m.reuse_features = False
        m.fuse_conv12 = fuse
        out.append(_run(m, obs, done, valid, mask))
    (b0, l0, g0), (b1, l1, g1) = out
    assert float((b0 - b1).abs().max()) <= 1e-3 * float(b0.abs().max())
    assert abs(float(l0) - float(l1)) <= 1e-3 * abs(float(l0))
    cos = float(torch.dot(g0, g1) / (g0.norm() * g1.norm()))
    assert cos >= 0.9995 and abs(float(g1.norm() / g0.norm()) - 1) <= 5e-3


def test_first_conv_kernels_are_repeatable():
    """Regression for the raw-frame-slot race (a slot handed back to the loader before the loads that read it had
    returned gave one warp's worth of stale pixels once in a few thousand samples): repeated launches on the same
    frames must agree bit for bit, stand-alone and fused."""
    m = _model()
    obs = _frames(1, 2500, seed=11)
    obs, T, B, base, strides = m._frame_view(obs)
    n = T * B
    pred, targ = m._chains
    first = None
    for _ in range(6):
        a1_p, _, a1_t = pred.layers[0].forward(base, n, in_dtype=L.U8, strides=strides, norm=m._norm, twin=targ.layers[0])
        img = m._normalized(base, n, strides[0])
        _, a2 = m._conv12(pred, img, n, False)
        tw = m._conv12_twin(img, n, True)
        torch.cuda.synchronize()
        if first is None:
            first = (a1_p, a1_t, a2) + tuple(tw)
        else:
            assert all(torch.equal(a, b) for a, b in zip((a1_p, a1_t, a2) + tuple(tw), first))
