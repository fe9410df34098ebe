"""The tcgen05 engine forced on (engine='umma' fails loudly if a shape has no specialisation):
same torch-fp32-on-bf16-operands reference as test_gpu_layers, at sizes that span several tiles
and more than one wave of the persistent grid."""
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


@pytest.fixture(autouse=True)
def _umma_engine():
    LY.set_engine("umma")
    yield
    LY.set_engine("auto")


@pytest.mark.parametrize("K,out,act,side_dim,use_res,n", [
    (512, 512, "relu", 0, False, 37),
    (512, 512, "none", 0, False, 300),
    (512, 512, "lrelu", 4, False, 1000),
    (512, 512, "none", 7, True, 40000),
    (1024, 512, "relu", 0, False, 513),
    (512, 64, "none", 0, False, 700),
    (512, 128, "relu", 0, False, 700),
    (64, 32, "none", 0, False, 5000),
])
def test_linear_forward(K, out, act, side_dim, use_res, n):
    torch.manual_seed(0)
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
    ref = {"none": lambda t: t, "relu": F.relu, "lrelu": F.leaky_relu}[act](ref)
    if use_res:
        ref = ref + res
    assert rel(yf, ref) < 2e-5
    assert rel(yb, ref) < 5e-3


@pytest.mark.parametrize("n", [3, 260, 1500])
def test_fc_over_flattened_map(n):
    torch.manual_seed(1)
    lin = nn.Linear(3136, 512).to(DEV)
    layer = LY.linear_layer(lin, "relu", spatial=(7, 64))
    a = bf(torch.randn(n, 64, 7, 7, device=DEV))
    _, yf = layer.forward(a.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16), n, out_f32=True)
    ref = F.relu(a.reshape(n, -1) @ bf(lin.weight).t() + lin.bias)
    assert rel(yf, ref) < 2e-5


@pytest.mark.parametrize("case", [
    (32, 20, 20, 64, 4, 2, 0, 5), (32, 20, 20, 64, 4, 2, 0, 1000),      # conv2
    (64, 9, 9, 64, 3, 1, 0, 7), (64, 9, 9, 64, 3, 1, 0, 2001),          # conv3
    (64, 9, 9, 128, 3, 1, 1, 33),                                       # padded stride-1 conv
])
def test_conv_forward(case):
    in_c, h, w, out_c, k, s, p, n = case
    torch.manual_seed(2)
    conv = nn.Conv2d(in_c, out_c, k, s, p).to(DEV)
    layer = LY.conv_layer(conv, (h, w, in_c), "lrelu")
    x = bf(torch.randn(n, in_c, h, w, device=DEV))
    yb, yf = layer.forward(x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16), n, out_f32=True)
    ref = F.leaky_relu(F.conv2d(x, bf(conv.weight), conv.bias, stride=s, padding=p))
    ref = ref.permute(0, 2, 3, 1).reshape(-1, out_c)
    assert rel(yf, ref) < 2e-5
    assert rel(yb, ref) < 5e-3


@pytest.mark.parametrize("K,out,n", [(512, 512, 129), (512, 512, 3000), (1024, 512, 77), (512, 64, 300)])
def test_linear_dgrad(K, out, n):
    torch.manual_seed(3)
    lin = nn.Linear(K, out).to(DEV)
    layer = LY.linear_layer(lin, "none")
    dy = bf(torch.randn(n, out, device=DEV))
    saved = bf(torch.randn(n, K, device=DEV))
    add = bf(torch.randn(n, K, device=DEV))
    dx = layer.dgrad(dy.to(torch.bfloat16), n, x_saved=saved.to(torch.bfloat16), x_act="relu",
                     dx_add=add.to(torch.bfloat16))
    ref = (dy @ bf(lin.weight) + add) * (saved > 0)
    assert rel(dx, ref) < 5e-3


@pytest.mark.parametrize("n", [4, 700])
def test_conv3_dgrad(n):
    torch.manual_seed(4)
    conv = nn.Conv2d(64, 64, 3, 1, 0).to(DEV)
    layer = LY.conv_layer(conv, (9, 9, 64), "lrelu")
    dy = bf(torch.randn(n, 64, 7, 7, device=DEV))
    xs = torch.zeros(n, 64, 9, 9, device=DEV, requires_grad=True)
    F.conv2d(xs, bf(conv.weight), None).backward(dy)
    saved = bf(torch.randn(n * 81, 64, device=DEV))
    dx = layer.dgrad(dy.permute(0, 2, 3, 1).reshape(-1, 64).contiguous().to(torch.bfloat16), n,
                     x_saved=saved.to(torch.bfloat16), x_act="lrelu")
    ref = xs.grad.permute(0, 2, 3, 1).reshape(-1, 64) * torch.where(saved > 0, 1.0, 0.01)
    assert rel(dx, ref) < 5e-3


@pytest.mark.parametrize("n", [2, 5, 777])
def test_conv2_dgrad_depth_to_space(n):
    torch.manual_seed(5)
    conv = nn.Conv2d(32, 64, 4, 2, 0).to(DEV)
    layer = LY.conv_layer(conv, (20, 20, 32), "lrelu")
    dy = bf(torch.randn(n, 64, 9, 9, device=DEV))
    xs = torch.zeros(n, 32, 20, 20, device=DEV, requires_grad=True)
    F.conv2d(xs, bf(conv.weight), None, stride=2).backward(dy)
    saved = bf(torch.randn(n * 400, 32, device=DEV))
    add = bf(torch.randn(n * 400, 32, device=DEV))
    dx = layer.dgrad(dy.permute(0, 2, 3, 1).reshape(-1, 64).contiguous().to(torch.bfloat16), n,
                     x_saved=saved.to(torch.bfloat16), x_act="lrelu", dx_add=add.to(torch.bfloat16))
    ref = (xs.grad.permute(0, 2, 3, 1).reshape(-1, 32) + add) * torch.where(saved > 0, 1.0, 0.01)
    assert rel(dx, ref) < 5e-3


@pytest.mark.parametrize("n", [6, 1000])
def test_fc_dgrad_flat(n):
    torch.manual_seed(6)
    lin = nn.Linear(3136, 512).to(DEV)
    layer = LY.linear_layer(lin, "none", spatial=(7, 64))
    dy = bf(torch.randn(n, 512, device=DEV))
    saved = bf(torch.randn(n * 49, 64, device=DEV))
    dx = layer.dgrad(dy.to(torch.bfloat16), n, x_saved=saved.to(torch.bfloat16), x_act="lrelu")
    ref = (dy @ bf(lin.weight)).reshape(n, 64, 7, 7).permute(0, 2, 3, 1).reshape(-1, 64)
    ref = ref * torch.where(saved > 0, 1.0, 0.01)
    assert rel(dx, ref) < 5e-3


@pytest.mark.parametrize("K,out,n", [(512, 512, 100), (512, 512, 5000), (1024, 512, 333), (512, 128, 70)])
def test_linear_wgrad(K, out, n):
    torch.manual_seed(7)
    lin = nn.Linear(K, out).to(DEV)
    layer = LY.linear_layer(lin, "none")
    x = bf(torch.randn(n, K, device=DEV))
    dy = bf(torch.randn(n, out, device=DEV))
    layer.wgrad(x.to(torch.bfloat16), dy.to(torch.bfloat16), n)
    assert rel(lin.weight.grad, dy.t() @ x) < 1e-4
    assert rel(lin.bias.grad, dy.sum(0)) < 1e-4
    layer.wgrad(x.to(torch.bfloat16), dy.to(torch.bfloat16), n, accumulate=True)
    assert rel(lin.weight.grad, 2 * (dy.t() @ x)) < 1e-4


@pytest.mark.parametrize("n", [5, 1000])
def test_fc_wgrad(n):
    torch.manual_seed(8)
    lin = nn.Linear(3136, 512).to(DEV)
    layer = LY.linear_layer(lin, "none", spatial=(7, 64))
    a = bf(torch.randn(n, 64, 7, 7, device=DEV))
    dy = bf(torch.randn(n, 512, device=DEV))
    layer.wgrad(a.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16), dy.to(torch.bfloat16), n)
    assert rel(lin.weight.grad, dy.t() @ a.reshape(n, -1)) < 1e-4


@pytest.mark.parametrize("case", [(64, 9, 9, 64, 3, 1, 0, 4), (64, 9, 9, 64, 3, 1, 0, 1001),
                                  (32, 20, 20, 64, 4, 2, 0, 3), (32, 20, 20, 64, 4, 2, 0, 777),
                                  (64, 7, 7, 64, 3, 1, 1, 50)])
def test_conv_wgrad(case):
    in_c, h, w, out_c, k, s, p, n = case
    torch.manual_seed(9)
    conv = nn.Conv2d(in_c, out_c, k, s, p).to(DEV)
    layer = LY.conv_layer(conv, (h, w, in_c), "lrelu")
    x = bf(torch.randn(n, in_c, h, w, device=DEV))
    dy = bf(torch.randn(n, out_c, layer.out_h, layer.out_w, device=DEV))
    wq = bf(conv.weight).detach().requires_grad_(True)
    bq = conv.bias.detach().clone().requires_grad_(True)
    F.conv2d(x, wq, bq, stride=s, padding=p).backward(dy)
    layer.wgrad(x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16),
                dy.permute(0, 2, 3, 1).reshape(-1, out_c).contiguous().to(torch.bfloat16), n)
    assert rel(conv.weight.grad, wq.grad) < 1e-4
    assert rel(conv.bias.grad, bq.grad) < 1e-4


@pytest.mark.parametrize("cin,n,norm", [(1, 3, True), (1, 700, True), (4, 2, False), (4, 301, False), (1, 5, False)])
def test_conv1_uint8_fwd_wgrad(cin, n, norm):
    """Burda / RND conv1 on tcgen05: uint8 NCHW frames read in place, normalisation in the loader."""
    torch.manual_seed(10)
    C = 4
    frames = torch.randint(0, 256, (n, C, 84, 84), dtype=torch.uint8, device=DEV)
    mean = torch.rand(84 * 84, device=DEV) * 255
    rstd = 1.0 / (torch.rand(84 * 84, device=DEV) * 60 + 20)
    conv = nn.Conv2d(cin, 32, 8, 4).to(DEV)
    layer = LY.conv_layer(conv, (84, 84, cin), "lrelu")
    if cin == 1:
        base = frames.data_ptr() + (C - 1) * 7056
        strides = (C * 7056, 84, 1, 7056)
        xin = frames[:, -1:].float()
    else:
        base = frames.data_ptr()
        strides = (C * 7056, 84, 1, 7056)
        xin = frames.float()
    nm = (mean, rstd) if norm else None
    if norm:
        xin = ((xin - mean.view(1, 1, 84, 84)) * rstd.view(1, 1, 84, 84)).clamp(-5, 5)
    yb, _ = layer.forward(base, n, in_dtype=L.U8, strides=strides, norm=nm)
    ref = F.leaky_relu(F.conv2d(bf(xin), bf(conv.weight), conv.bias, stride=4))
    assert rel(yb, ref.permute(0, 2, 3, 1).reshape(-1, 32)) < 5e-3
    dy = bf(torch.randn(n, 32, 20, 20, device=DEV))
    dyb = dy.permute(0, 2, 3, 1).reshape(-1, 32).contiguous().to(torch.bfloat16)
    wq = bf(conv.weight).detach().requires_grad_(True)
    bq = conv.bias.detach().clone().requires_grad_(True)
    F.conv2d(bf(xin), wq, bq, stride=4).backward(dy)
    layer.wgrad(base, dyb, n, in_dtype=L.U8, strides=strides, norm=nm)
    assert rel(conv.weight.grad, wq.grad) < 1e-4
    assert rel(conv.bias.grad, bq.grad) < 1e-4
    layer.wgrad(base, dyb, n, in_dtype=L.U8, strides=strides, norm=nm, accumulate=True)
    assert rel(conv.weight.grad, 2 * wq.grad) < 1e-4


@pytest.mark.parametrize("E,n,shared", [(4, 300, True), (4, 1000, False), (2, 40, False)])
def test_grouped_linear(E, n, shared):
    """cb_grouped_linear_{fwd,dgrad,wgrad}: E Linear(512+A -> 512) problems in one launch against E
    separate fp32 matmuls on the same bf16 operands (disagreement.py:97-100)."""
    import ctypes as C
    torch.manual_seed(5)
    Fd, A = 512, 7
    W = bf(torch.randn(E, Fd, Fd, device=DEV) * 0.05)
    Ws = torch.randn(E, Fd, A, device=DEV) * 0.1
    b = torch.randn(E, Fd, device=DEV)
    x = bf(torch.randn(n, Fd if shared else E * Fd, device=DEV))
    res = bf(torch.randn(n, E * Fd, device=DEV))
    side = F.one_hot(torch.randint(0, A, (n,), device=DEV), A).float()
    yb = torch.empty(n, E * Fd, dtype=torch.bfloat16, device=DEV)
    yf = torch.empty(n, E * Fd, dtype=torch.float32, device=DEV)
    e = L.Epilogue()
    w_all = W.reshape(E * Fd, Fd).to(torch.bfloat16).contiguous()
    b_all, sw_all = b.reshape(-1).contiguous(), Ws.reshape(E * Fd, A).contiguous()
    xb, resb = x.to(torch.bfloat16).contiguous(), res.to(torch.bfloat16).contiguous()
    e.bias, e.residual, e.side, e.side_w, e.side_dim = L.ptr(b_all), L.ptr(resb), L.ptr(side), L.ptr(sw_all), A
    e.out_bf16, e.out_f32 = L.ptr(yb), L.ptr(yf)
    L.call("cb_grouped_linear_fwd", n, E, Fd, Fd, int(shared), L.ptr(xb), L.ptr(w_all), L.ACT_LRELU, C.byref(e), L.stream())
    xs = [x if shared else x[:, g * Fd:(g + 1) * Fd] for g in range(E)]
    ref = torch.cat([F.leaky_relu(xs[g] @ W[g].t() + b[g] + side @ Ws[g].t()) + res[:, g * Fd:(g + 1) * Fd]
                     for g in range(E)], 1)
    assert rel(yf, ref) < 2e-5
    assert rel(yb, ref) < 5e-3
    # dgrad: dx_g = (dy_g W_g + add_g) * lrelu'(saved_g)
    dy = bf(torch.randn(n, E * Fd, device=DEV))
    add = bf(torch.randn(n, E * Fd, device=DEV))
    saved = bf(torch.randn(n, E * Fd, device=DEV))
    w_t_all = W.transpose(1, 2).reshape(E * Fd, Fd).to(torch.bfloat16).contiguous()     # [g*K + i][j] = W_g[j][i]
    dx = torch.empty(n, E * Fd, dtype=torch.bfloat16, device=DEV)
    dyb, savedb, addb = (t.to(torch.bfloat16).contiguous() for t in (dy, saved, add))    # keep the operands alive
    L.call("cb_grouped_linear_dgrad", n, E, Fd, Fd, L.ptr(dyb), L.ptr(w_t_all), L.ptr(savedb), L.ACT_LRELU,
           L.ptr(addb), L.ptr(dx), L.stream())
    ref = torch.cat([(dy[:, g * Fd:(g + 1) * Fd] @ W[g] + add[:, g * Fd:(g + 1) * Fd]) *
                     torch.where(saved[:, g * Fd:(g + 1) * Fd] > 0, 1.0, 0.01) for g in range(E)], 1)
    assert rel(dx, ref) < 5e-3
    # wgrad: dW_g = dy_g^T x_g, db_g = column sums
    dw = torch.empty(E * Fd, Fd, device=DEV)
    db = torch.empty(E * Fd, device=DEV)
    L.call("cb_grouped_linear_wgrad", n, E, Fd, Fd, int(shared), L.ptr(xb), L.ptr(dyb), L.ptr(dw), L.ptr(db), 0.0,
           L.stream())
    ref = torch.cat([dy[:, g * Fd:(g + 1) * Fd].t() @ xs[g] for g in range(E)], 0)
    assert rel(dw, ref) < 1e-4
    assert rel(db, dy.sum(0)) < 1e-4


@pytest.mark.parametrize("A,n,use_res", [(4, 1000, False), (7, 513, True), (12, 37, False)])
def test_linear_side_as_kblock(A, n, use_res):
    """cb_epilogue.side_bf16: the one-hot action concatenated on the tensor cores (icm.py:19-21) must equal
    the fp32-epilogue path up to the bf16 rounding of the side weights."""
    torch.manual_seed(11)
    lin = nn.Linear(512 + A, 512).to(DEV)
    layer = LY.linear_layer(lin, "lrelu", A)
    x = bf(torch.randn(n, 512, device=DEV))
    side = F.one_hot(torch.randint(0, A, (n,), device=DEV), A).float()
    res = bf(torch.randn(n, 512, device=DEV)) if use_res else None
    resb = res.to(torch.bfloat16) if use_res else None
    sp = LY.pad_side(side)
    _, yf = layer.forward(x.to(torch.bfloat16), n, side=side, residual=resb, out_f32=True, side_pad=sp)
    W = lin.weight.detach()
    ref = F.leaky_relu(x @ bf(W[:, :512]).t() + lin.bias + side @ bf(W[:, 512:]).t())
    if use_res:
        ref = ref + res
    assert rel(yf, ref) < 2e-5
    _, y_epi = layer.forward(x.to(torch.bfloat16), n, side=side, residual=resb, out_f32=True)
    assert rel(yf, y_epi) < 5e-3


@pytest.mark.parametrize("n,act", [(80001, "relu"), (131072 + 77, "none")])
def test_linear_big_tiles(n, act):
    """Above 2*148*256 rows the Linear layers switch to 256x256 tiles with a single-buffered TMEM accumulator
    (umma_gemm_kernel<256, 2>); ragged last tile included.  Forward, fp32 output, dgrad and the two-accumulator
    row weight gradient (K >= 1024) against fp32 matmuls on the same bf16 operands."""
    torch.manual_seed(3)
    lin = nn.Linear(512, 512).to(DEV)
    layer = LY.linear_layer(lin, act)
    x = bf(torch.randn(n, 512, device=DEV))
    xb = x.to(torch.bfloat16)
    yb, yf = layer.forward(xb, n, out_f32=True)
    ref = x @ bf(lin.weight.detach()).t() + lin.bias
    ref = F.relu(ref) if act == "relu" else ref
    assert rel(yf, ref) < 2e-5
    assert rel(yb, ref) < 5e-3
    dy = bf(torch.randn(n, 512, device=DEV))
    dx = layer.dgrad(dy.to(torch.bfloat16), n, x_saved=xb, x_act="lrelu")
    ref_dx = (dy @ bf(lin.weight.detach())) * torch.where(x > 0, 1.0, 0.01)
    assert rel(dx, ref_dx) < 5e-3
    fc = nn.Linear(1024, 512).to(DEV)
    lfc = LY.linear_layer(fc, "none")
    x2 = bf(torch.randn(n, 1024, device=DEV))
    lfc.wgrad(x2.to(torch.bfloat16), dy.to(torch.bfloat16), n)
    torch.cuda.synchronize()
    assert rel(fc.weight.grad, dy.t() @ x2) < 2e-4
    assert rel(fc.bias.grad, dy.sum(0)) < 2e-4
