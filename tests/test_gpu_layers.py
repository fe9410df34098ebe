"""Layer-level parity of the library's GEMM-shaped entry points against torch fp32 ops fed with
the same bf16-rounded operands (so the only difference is fp32 summation order), for every
layer shape on the path, through both engines when a tcgen05 specialisation exists."""
import pytest
import torch
import torch.nn.functional as F
from torch import nn

from curiosity_baselines_b200 import _lib as L
from curiosity_baselines_b200 import layers as LY

pytestmark = pytest.mark.gpu
DEV = "cuda"


def bf(x):
    return x.to(torch.bfloat16).float()


def rel(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-20))


ACTS = {"none": lambda x: x, "relu": F.relu, "lrelu": F.leaky_relu, "elu": F.elu}

CONV_CASES = [  # (in_c, h, w, out_c, k, stride, pad, act, n)
    (1, 84, 84, 32, 8, 4, 0, "lrelu", 3),      # RND conv1
    (4, 84, 84, 32, 8, 4, 0, "lrelu", 2),      # Burda conv1
    (32, 20, 20, 64, 4, 2, 0, "lrelu", 5),     # conv2
    (64, 9, 9, 64, 3, 1, 0, "lrelu", 7),       # conv3
    (4, 84, 84, 32, 3, 2, 1, "elu", 2),        # Universe l0
    (32, 42, 42, 32, 3, 2, 1, "elu", 2),       # Universe l1
    (32, 6, 6, 32, 3, 2, 1, "elu", 9),         # Universe l4
    (3, 5, 5, 16, 3, 1, 1, "relu", 11),        # Maze conv1
    (16, 5, 5, 16, 3, 2, 2, "relu", 11),       # Maze conv2
]


def engines():
    out = ["simt"]
    return out + ["auto"]


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("engine", engines())
def test_conv_fwd_dgrad_wgrad(case, engine):
    in_c, h, w, out_c, k, s, p, act, n = case
    torch.manual_seed(0)
    LY.set_engine(engine)
    try:
        conv = nn.Conv2d(in_c, out_c, k, s, p).to(DEV)
        with torch.no_grad():
            conv.bias.uniform_(-0.5, 0.5)
        layer = LY.conv_layer(conv, (h, w, in_c), act)
        x = bf(torch.randn(n, in_c, h, w, device=DEV))
        x_nhwc = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
        yb, yf = layer.forward(x_nhwc, n, out_f32=True)
        ref = ACTS[act](F.conv2d(x, bf(conv.weight), conv.bias, stride=s, padding=p))
        ref_nhwc = ref.permute(0, 2, 3, 1).reshape(-1, out_c)
        assert rel(yf, ref_nhwc) < 2e-5
        assert rel(yb, ref_nhwc) < 5e-3
        # backward
        dy = bf(torch.randn(n, out_c, layer.out_h, layer.out_w, device=DEV))
        dy_nhwc = dy.permute(0, 2, 3, 1).reshape(-1, out_c).contiguous().to(torch.bfloat16)
        xs = x.clone().requires_grad_(True)
        wq = bf(conv.weight).detach().requires_grad_(True)
        bq = conv.bias.detach().clone().requires_grad_(True)
        F.conv2d(xs, wq, bq, stride=s, padding=p).backward(dy)
        dx = layer.dgrad(dy_nhwc, n)
        assert rel(dx, xs.grad.permute(0, 2, 3, 1).reshape(-1, in_c)) < 5e-3
        conv.weight.grad = None; conv.bias.grad = None
        layer.wgrad(x_nhwc, dy_nhwc, n)
        assert rel(conv.weight.grad, wq.grad) < 1e-4
        assert rel(conv.bias.grad, bq.grad) < 1e-4
        layer.wgrad(x_nhwc, dy_nhwc, n, accumulate=True)
        assert rel(conv.weight.grad, 2 * wq.grad) < 1e-4
        assert rel(conv.bias.grad, 2 * bq.grad) < 1e-4
    finally:
        LY.set_engine("auto")


@pytest.mark.parametrize("engine", engines())
def test_conv1_uint8_nchw_with_normalisation(engine):
    """RND conv1 reads the newest uint8 frame in place and normalises in the loader."""
    torch.manual_seed(1)
    LY.set_engine(engine)
    try:
        n, C = 4, 4
        frames = torch.randint(0, 256, (n, C, 84, 84), dtype=torch.uint8, device=DEV)
        mean = torch.rand(84 * 84, device=DEV) * 255
        rstd = 1.0 / (torch.rand(84 * 84, device=DEV) * 60 + 20)
        conv = nn.Conv2d(1, 32, 8, 4).to(DEV)
        layer = LY.conv_layer(conv, (84, 84, 1), "lrelu")
        base = frames.data_ptr() + (C - 1) * 7056
        strides = (C * 7056, 84, 1, 7056)
        _, yf = layer.forward(base, n, in_dtype=L.U8, strides=strides, norm=(mean, rstd), out_f32=True)
        xn = ((frames[:, -1:].float() - mean.view(1, 1, 84, 84)) * rstd.view(1, 1, 84, 84)).clamp(-5, 5)
        ref = F.leaky_relu(F.conv2d(bf(xn), bf(conv.weight), conv.bias, stride=4))
        assert rel(yf, ref.permute(0, 2, 3, 1).reshape(-1, 32)) < 2e-5
        dy = bf(torch.randn(n, 32, 20, 20, device=DEV))
        dyb = dy.permute(0, 2, 3, 1).reshape(-1, 32).contiguous().to(torch.bfloat16)
        wq = bf(conv.weight).detach().requires_grad_(True)
        F.conv2d(bf(xn), wq, None, stride=4).backward(dy)
        layer.wgrad(base, dyb, n, in_dtype=L.U8, strides=strides, norm=(mean, rstd))
        assert rel(conv.weight.grad, wq.grad) < 1e-4
    finally:
        LY.set_engine("auto")


LIN_CASES = [  # (K, out, act, side_dim, residual, n)
    (512, 512, "relu", 0, False, 37),
    (512, 512, "none", 0, False, 300),
    (1024, 512, "relu", 0, False, 50),
    (512, 4, "none", 0, False, 50),
    (512, 512, "lrelu", 4, False, 129),
    (512, 512, "none", 7, True, 64),
    (288, 288, "lrelu", 12, True, 33),
    (128 + 15, 64, "relu", 0, False, 40),
    (64, 75, "none", 0, False, 40),
]


@pytest.mark.parametrize("case", LIN_CASES)
@pytest.mark.parametrize("engine", engines())
def test_linear_fwd_bwd(case, engine):
    K, out, act, side_dim, use_res, n = case
    torch.manual_seed(2)
    LY.set_engine(engine)
    try:
        lin = nn.Linear(K + side_dim, out).to(DEV)
        layer = LY.linear_layer(lin, act, side_dim)
        x = bf(torch.randn(n, K, device=DEV))
        side = F.one_hot(torch.randint(0, max(side_dim, 1), (n,), device=DEV), max(side_dim, 1)).float() if side_dim else None
        res = bf(torch.randn(n, out, device=DEV)) if use_res else None
        yb, yf = layer.forward(x.to(torch.bfloat16), n, side=side,
                               residual=res.to(torch.bfloat16) if use_res else None, out_f32=True)
        W = lin.weight.detach()
        ref = x @ bf(W[:, :K]).t() + lin.bias
        if side_dim:
            ref = ref + side @ W[:, K:].t()
        ref = ACTS[act](ref)
        if use_res:
            ref = ref + res
        assert rel(yf, ref) < 2e-5
        dy = bf(torch.randn(n, out, device=DEV))
        dyb = dy.to(torch.bfloat16)
        saved = bf(torch.randn(n, K, device=DEV))
        add = bf(torch.randn(n, K, device=DEV))
        dx = layer.dgrad(dyb, n, x_saved=saved.to(torch.bfloat16), x_act="lrelu", dx_add=add.to(torch.bfloat16))
        ref_dx = (dy @ bf(W[:, :K]) + add) * torch.where(saved > 0, 1.0, 0.01)
        assert rel(dx, ref_dx) < 5e-3
        lin.weight.grad = None; lin.bias.grad = None
        layer.wgrad(x.to(torch.bfloat16), dyb, n, side=side)
        ref_dw = dy.t() @ x
        if side_dim:
            ref_dw = torch.cat([ref_dw, dy.t() @ side], 1)
        assert rel(lin.weight.grad, ref_dw) < 1e-4
        assert rel(lin.bias.grad, dy.sum(0)) < 1e-4
    finally:
        LY.set_engine("auto")


def test_fc_after_flatten_matches_chw_order():
    """Burda's Linear(3136->512) consumes a CHW flatten; we run it as a 7x7 conv over NHWC."""
    torch.manual_seed(3)
    n = 6
    lin = nn.Linear(3136, 512).to(DEV)
    layer = LY.linear_layer(lin, "none", spatial=(7, 64))
    a = bf(torch.randn(n, 64, 7, 7, device=DEV))
    a_nhwc = a.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    _, yf = layer.forward(a_nhwc, n, out_f32=True)
    ref = a.reshape(n, -1) @ bf(lin.weight).t() + lin.bias
    assert rel(yf, ref) < 2e-5
    dy = bf(torch.randn(n, 512, device=DEV))
    layer.wgrad(a_nhwc, dy.to(torch.bfloat16), n)
    assert rel(lin.weight.grad, dy.t() @ a.reshape(n, -1)) < 1e-4
    dx = layer.dgrad(dy.to(torch.bfloat16), n)
    ref_dx = (dy @ bf(lin.weight)).reshape(n, 64, 7, 7).permute(0, 2, 3, 1).reshape(-1, 64)
    assert rel(dx, ref_dx) < 5e-3
