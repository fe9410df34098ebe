"""Host-side driver for the GEMM-shaped layers of the curiosity models.

A ``GemmLayer`` binds one ``torch.nn.Conv2d`` / ``torch.nn.Linear`` parameter holder (kept only
so that ``state_dict`` keys match the reference) to the library's conv entry points:
prepared bf16 weights in the kernel layout, forward, dgrad and wgrad.  ``torch`` is used for
allocation and for copying gradients into ``param.grad`` -- no arithmetic of the path runs in
torch.
"""
import ctypes as C

import torch

from . import _lib as L
from . import dist

ACT = {"none": L.ACT_NONE, "relu": L.ACT_RELU, "lrelu": L.ACT_LRELU, "elu": L.ACT_ELU}

_engine = L.ENGINE_AUTO


def set_engine(engine):
    """Force the engine of every subsequent GEMM-shaped call: 'auto' | 'simt' | 'umma'."""
    global _engine
    _engine = {"auto": L.ENGINE_AUTO, "simt": L.ENGINE_SIMT, "umma": L.ENGINE_UMMA}[engine]


def get_engine():
    return _engine


class GemmLayer:
    """One conv / linear layer.

    in_shape  (h, w, c) of the NHWC input (linear: (1, 1, K))
    k, stride, pad, out_c, act
    side_dim  extra fp32 input columns concatenated after the c channels (one-hot action,
              icm.py:19-21); only for linear layers.
    """

    def __init__(self, module, in_shape, k, stride, pad, out_c, act="none", side_dim=0,
                 trainable=True):
        self.module = module
        self.in_h, self.in_w, self.in_c = in_shape
        self.k, self.stride, self.pad, self.out_c = k, stride, pad, out_c
        self.out_h = (self.in_h + 2 * pad - k) // stride + 1
        self.out_w = (self.in_w + 2 * pad - k) // stride + 1
        self.act = ACT[act]
        self.side_dim = side_dim
        self.trainable = trainable
        self.K = k * k * self.in_c
        # a Linear over a flattened [c,k,k] map: its dgrad is a plain [*, K] linear dgrad
        self.flat_linear = k > 1 and k == self.in_h == self.in_w and pad == 0
        self.w_fwd = self.w_t = self.w_shuffle = self.w_native = self.bias = self.side_w = None
        self._version = None

    # ---------------------------------------------------------------- weights
    def _torch_weight_4d(self):
        """View of the holder's weight as [out_c, in_c, k, k] (+ the side columns)."""
        w = self.module.weight
        if w.dim() == 4:
            return w, None
        main = w[:, : self.K] if self.side_dim else w
        side = w[:, self.K:] if self.side_dim else None
        # a Linear after a CHW flatten is a kxk conv over the (k,k,c) NHWC map: [o, c*k*k]
        return main.reshape(self.out_c, self.in_c, self.k, self.k), side

    def prepare(self, force=False):
        """(Re)build the bf16 kernel-layout copies when the fp32 master weights changed."""
        w = self.module.weight
        ver = (w._version, w.data_ptr(), self.module.bias._version)
        if not force and ver == self._version:
            return
        dev = w.device
        w4, side = self._torch_weight_4d()
        w4 = w4.detach().contiguous()
        if self.w_fwd is None:
            self.w_fwd = torch.empty(self.out_c, self.K, dtype=torch.bfloat16, device=dev)
            self.w_t = torch.empty(self.in_c, self.k * self.k * self.out_c, dtype=torch.bfloat16,
                                   device=dev)
        L.call("cb_prepare_conv_weight", L.ptr(w4), L.ptr(self.w_fwd), L.ptr(self.w_t),
               self.out_c, self.in_c, self.k, self.k, L.stream())
        if self.k == 8 and self.stride == 4 and self.pad == 0 and self.out_c == 32 and self.in_c <= 4:
            # first Burda conv: the uint8 tcgen05 kernel consumes torch's own (c, kh, kw) order
            self.w_native = w4.reshape(self.out_c, -1).to(torch.bfloat16).contiguous()
        if self.flat_linear:        # [K, out_c] with K in (h, w, c) order
            self.w_t = self.w_fwd.t().contiguous()
        if self.stride == 2 and self.k == 4 and self.pad == 0 and self.in_c == 32:
            # depth-to-space dgrad operand: [(ph,pw,ci)][(u,v,co)] = W[co][ci][2u+ph][2v+pw]
            w6 = w4.view(self.out_c, self.in_c, 2, 2, 2, 2)          # co ci u ph v pw
            self.w_shuffle = w6.permute(3, 5, 1, 2, 4, 0).reshape(4 * self.in_c, 4 * self.out_c) \
                .to(torch.bfloat16).contiguous()
        self.bias = self.module.bias.detach()
        self.side_w = side.detach().contiguous() if side is not None else None
        self.w_ext = None
        if side is not None and self.k == 1 and self.K % 64 == 0 and self.side_dim <= 64:
            # [out, K + 64]: the side weights as one more 64-wide k-block (cb_epilogue.side_bf16)
            self.w_ext = torch.zeros(self.out_c, self.K + 64, dtype=torch.bfloat16, device=dev)
            self.w_ext[:, : self.K] = self.w_fwd
            self.w_ext[:, self.K: self.K + self.side_dim] = self.side_w.to(torch.bfloat16)
        self._version = ver

    # ---------------------------------------------------------------- descriptors
    def desc(self, n, in_dtype=L.BF16, strides=None, act=None):
        d = L.ConvDesc()
        d.n = n
        d.in_h, d.in_w, d.in_c = self.in_h, self.in_w, self.in_c
        d.k_h = d.k_w = self.k
        d.stride, d.pad = self.stride, self.pad
        d.out_h, d.out_w, d.out_c = self.out_h, self.out_w, self.out_c
        d.in_dtype = in_dtype
        if strides is None:   # dense NHWC
            strides = (self.in_h * self.in_w * self.in_c, self.in_w * self.in_c, self.in_c, 1)
        d.in_stride_n, d.in_stride_h, d.in_stride_w, d.in_stride_c = strides
        d.act = self.act if act is None else act
        d.engine = _engine
        return d

    # ---------------------------------------------------------------- forward
    def forward(self, x, n, in_dtype=L.BF16, strides=None, side=None, residual=None, norm=None,
                out_bf16=True, out_f32=False, twin=None, side_pad=None):
        """Returns (bf16 NHWC output or None, fp32 output or None), each [n*out_h*out_w, out_c].
        ``twin``: a second first-conv layer of identical shape reading the same uint8 input (RND
        target + predictor); its bf16 output is then returned as a third element."""
        self.prepare()
        m = n * self.out_h * self.out_w
        dev = self.w_fwd.device
        yb = torch.empty(m, self.out_c, dtype=torch.bfloat16, device=dev) if out_bf16 else None
        yf = torch.empty(m, self.out_c, dtype=torch.float32, device=dev) if out_f32 else None
        e = L.Epilogue()
        e.bias = L.ptr(self.bias)
        e.residual = L.ptr(residual)
        w_fwd = self.w_fwd
        if self.side_dim:
            assert side is not None and side.shape[-1] == self.side_dim
            e.side, e.side_w, e.side_dim = L.ptr(side), L.ptr(self.side_w), self.side_dim
            if side_pad is not None and self.w_ext is not None and _engine != L.ENGINE_SIMT:
                e.side_bf16 = L.ptr(side_pad)        # concatenation on the tensor cores (one more k-block)
                w_fwd = self.w_ext
        if norm is not None:
            e.norm_mean, e.norm_rstd = L.ptr(norm[0]), L.ptr(norm[1])
        e.out_bf16, e.out_f32 = L.ptr(yb), L.ptr(yf)
        d = self.desc(n, in_dtype, strides)
        xp = x if isinstance(x, int) else L.ptr(x)
        if self._use_conv1(d) and yb is not None and yf is None:
            y2 = None
            if twin is not None:
                twin.prepare()
                y2 = torch.empty_like(yb)
                L.call("cb_conv1_fwd", C.byref(d), xp, L.ptr(self.w_native), C.byref(e), L.ptr(twin.w_native),
                       L.ptr(twin.bias), L.ptr(y2), L.stream())
                return yb, yf, y2
            L.call("cb_conv1_fwd", C.byref(d), xp, L.ptr(self.w_native), C.byref(e), None, None, None, L.stream())
        else:
            L.call("cb_conv_fwd", C.byref(d), xp, L.ptr(w_fwd), C.byref(e), L.stream())
        if twin is not None:      # no fused kernel for this shape/engine: run the twin separately
            y2, _ = twin.forward(x, n, in_dtype=in_dtype, strides=strides, norm=norm)
            return yb, yf, y2
        return yb, yf

    def _use_conv1(self, d):
        return (self.w_native is not None and d.in_dtype == L.U8 and _engine != L.ENGINE_SIMT
                and bool(L.load().cb_conv1_supported(C.byref(d))))

    # ---------------------------------------------------------------- backward
    def dgrad(self, dy, n, x_saved=None, x_act="none", dx_add=None):
        """dx (bf16 NHWC [n*in_h*in_w, in_c]) = (dy (*) w + dx_add) .* act'(x_saved)."""
        self.prepare()
        dx = torch.empty(n * self.in_h * self.in_w, self.in_c, dtype=torch.bfloat16,
                         device=dy.device)
        if self.flat_linear:
            d = L.ConvDesc()
            d.n, d.in_h, d.in_w, d.in_c = n, 1, 1, self.K
            d.k_h = d.k_w = d.stride = 1
            d.out_h = d.out_w = 1
            d.out_c = self.out_c
            d.in_dtype = L.BF16
            d.in_stride_n, d.in_stride_h, d.in_stride_w, d.in_stride_c = self.K, self.K, self.K, 1
            d.engine = _engine
        else:
            d = self.desc(n)
            if (self.w_shuffle is not None and _engine != L.ENGINE_SIMT
                    and L.load().cb_conv_dgrad_s2_supported(C.byref(d))):
                L.call("cb_conv_dgrad_s2", C.byref(d), L.ptr(dy), L.ptr(self.w_shuffle), L.ptr(x_saved),
                       ACT[x_act], L.ptr(dx_add), L.ptr(dx), L.stream())
                return dx
        L.call("cb_conv_dgrad", C.byref(d), L.ptr(dy), L.ptr(self.w_t), L.ptr(x_saved), ACT[x_act],
               L.ptr(dx_add), L.ptr(dx), L.stream())
        return dx

    def wgrad(self, x, dy, n, in_dtype=L.BF16, strides=None, side=None, norm=None, accumulate=False, final=True):
        """Weight / bias gradients written (or accumulated) into the holder's ``.grad`` in the
        torch layout.  ``final``: no later pass of this step accumulates into them any more, so the data-parallel
        reducer may start summing them across ranks (dist.BucketReducer)."""
        if not self.trainable:
            return
        self._wgrad_impl(x, dy, n, in_dtype=in_dtype, strides=strides, side=side, norm=norm, accumulate=accumulate)
        if final:
            dist.grads_ready((self.module.weight.grad, self.module.bias.grad))

    def _wgrad_impl(self, x, dy, n, in_dtype=L.BF16, strides=None, side=None, norm=None, accumulate=False):
        self.prepare()
        dev = dy.device
        w, b = self.module.weight, self.module.bias
        if w.grad is None:
            w.grad = torch.zeros_like(w)
        if b.grad is None:
            b.grad = torch.zeros_like(b)
        beta = 1.0 if accumulate else 0.0
        direct = self.k == 1 and not self.side_dim and w.dim() == 2
        dw = w.grad if direct else torch.empty(self.out_c, self.K, dtype=torch.float32, device=dev)
        d = self.desc(n, in_dtype, strides)
        xp = x if isinstance(x, int) else L.ptr(x)
        mean, rstd = (L.ptr(norm[0]), L.ptr(norm[1])) if norm is not None else (None, None)
        if self._use_conv1(d):      # gradient lands directly in the torch layout
            L.call("cb_conv1_wgrad", C.byref(d), xp, L.ptr(dy), mean, rstd, L.ptr(w.grad), L.ptr(b.grad), beta,
                   L.stream())
            return
        if direct or not accumulate:
            L.call("cb_conv_wgrad", C.byref(d), xp, L.ptr(dy), mean, rstd, L.ptr(dw),
                   L.ptr(b.grad), beta if direct else 0.0, L.stream())
        else:   # dw is scratch (beta 0); the bias gradient must still accumulate
            db = torch.empty_like(b.grad)
            L.call("cb_conv_wgrad", C.byref(d), xp, L.ptr(dy), mean, rstd, L.ptr(dw), L.ptr(db),
                   0.0, L.stream())
            b.grad.add_(db)
        if direct:
            return
        if not self.side_dim:      # conv, or Linear over a flattened [c,k,k] map
            g4 = w.grad.view(self.out_c, self.in_c, self.k, self.k)
            L.call("cb_unprepare_conv_grad", L.ptr(dw), L.ptr(g4), self.out_c, self.in_c, self.k,
                   self.k, beta, L.stream())
            return
        m = n * self.out_h * self.out_w
        dsw = torch.empty(self.out_c, self.side_dim, dtype=torch.float32, device=dev)
        L.call("cb_side_wgrad", L.ptr(dy), L.ptr(side), L.ptr(dsw), m, self.out_c, self.side_dim,
               0.0, L.stream())
        if accumulate:
            w.grad[:, : self.K].add_(dw)
            w.grad[:, self.K:].add_(dsw)
        else:
            w.grad[:, : self.K].copy_(dw)
            w.grad[:, self.K:].copy_(dsw)

class DenseConvLayer(GemmLayer):
    """A convolution on a tiny fixed grid (PyColab 5x5 mazes, encoders.py:37-66) run as ONE dense Linear over the
    flattened NHWC map: the conv weight is expanded into its block-Toeplitz matrix
    [round_up(oh*ow*oc, 128), round_up(ih*iw*ic, 8)], so these 3- and 16-channel layers run on the tcgen05 GEMM
    (2.8x redundant MACs on ~0.1 MFLOP per sample) instead of the SIMT engine.  Gradients of the expanded matrix
    are folded back onto the conv weight with one index_add_."""

    def __init__(self, module, in_shape, act="none", trainable=True, in_features=None):
        h, w, c = in_shape
        k, s, pd = module.kernel_size[0], module.stride[0], module.padding[0]
        oc = module.out_channels
        oh, ow = (h + 2 * pd - k) // s + 1, (w + 2 * pd - k) // s + 1
        self.real_in, self.real_out = h * w * c, oh * ow * oc
        k_in = in_features if in_features is not None else (self.real_in + 7) // 8 * 8
        n_out = (self.real_out + 127) // 128 * 128
        super().__init__(module, (1, 1, k_in), 1, 1, 0, n_out, act, 0, trainable)
        self.conv_out_h, self.conv_out_w, self.conv_out_c = oh, ow, oc    # geometry seen by the next layer
        rows, cols, src = [], [], []
        for oy in range(oh):
            for ox in range(ow):
                for ky in range(k):
                    for kx in range(k):
                        iy, ix = oy * s + ky - pd, ox * s + kx - pd
                        if 0 <= iy < h and 0 <= ix < w:
                            for o in range(oc):
                                for i in range(c):
                                    rows.append((oy * ow + ox) * oc + o)
                                    cols.append((iy * w + ix) * c + i)
                                    src.append(((o * c + i) * k + ky) * k + kx)
        self._flat = torch.tensor(rows, dtype=torch.int64) * k_in + torch.tensor(cols, dtype=torch.int64)
        self._src = torch.tensor(src, dtype=torch.int64)
        self._brow = torch.arange(self.real_out, dtype=torch.int64) % oc          # conv output channel of each dense row
        self._dense_w = self._dense_b = None

    def prepare(self, force=False):
        w, b = self.module.weight, self.module.bias
        ver = (w._version, w.data_ptr(), b._version)
        if not force and ver == self._version:
            return
        dev = w.device
        if self._flat.device != dev:
            self._flat, self._src, self._brow = self._flat.to(dev), self._src.to(dev), self._brow.to(dev)
        dense = torch.zeros(self.out_c * self.K, dtype=torch.float32, device=dev)
        dense[self._flat] = w.detach().reshape(-1)[self._src]
        dense = dense.view(self.out_c, self.K)
        if self.w_fwd is None:
            self.w_fwd = torch.empty(self.out_c, self.K, dtype=torch.bfloat16, device=dev)
            self.w_t = torch.empty(self.K, self.out_c, dtype=torch.bfloat16, device=dev)
        L.call("cb_prepare_conv_weight", L.ptr(dense), L.ptr(self.w_fwd), L.ptr(self.w_t), self.out_c, self.K, 1, 1,
               L.stream())
        bias = torch.zeros(self.out_c, dtype=torch.float32, device=dev)
        bias[: self.real_out] = b.detach()[self._brow]
        self.bias, self.side_w, self.w_ext = bias, None, None
        self._dense_w = dense                      # keeps the fp32 source alive until the conversion kernel ran
        self._version = ver

    def _flat_input(self, x, n):
        """bf16 NHWC map [n, h, w, c] (or already flat) -> [n, K] with zero padding columns."""
        x = x.reshape(n, -1)
        if x.dtype != torch.bfloat16:
            x = x.to(torch.bfloat16)
        if x.shape[1] != self.K:
            x = torch.nn.functional.pad(x, (0, self.K - x.shape[1]))
        return x.contiguous()

    def forward(self, x, n, in_dtype=L.BF16, strides=None, **kw):
        kw.pop("norm", None)
        return super().forward(self._flat_input(x, n), n, **kw)

    def _wgrad_impl(self, x, dy, n, in_dtype=L.BF16, strides=None, side=None, norm=None, accumulate=False):
        self.prepare()
        w, b = self.module.weight, self.module.bias
        dev = dy.device
        dw = torch.empty(self.out_c, self.K, dtype=torch.float32, device=dev)
        db = torch.empty(self.out_c, dtype=torch.float32, device=dev)
        d = self.desc(n)
        L.call("cb_conv_wgrad", C.byref(d), L.ptr(self._flat_input(x, n)), L.ptr(dy), None, None, L.ptr(dw), L.ptr(db), 0.0,
               L.stream())
        if w.grad is None or not accumulate:
            w.grad = torch.zeros_like(w) if w.grad is None else w.grad.zero_()
        if b.grad is None or not accumulate:
            b.grad = torch.zeros_like(b) if b.grad is None else b.grad.zero_()
        w.grad.view(-1).index_add_(0, self._src, dw.view(-1)[self._flat])          # fold the Toeplitz entries
        b.grad.index_add_(0, self._brow, db[: self.real_out])


class PaddedLinearLayer(GemmLayer):
    """A Linear layer whose sizes the tcgen05 kernels do not tile (NDIGO's 128+kA -> 64 -> 75 predictors,
    ndigo.py:19-34) run with zero-padded weights: outputs padded to a multiple of 128 (zero rows -> exact zeros
    after ReLU), inputs to ``in_features``.  Callers see the padded width; ``real_out`` columns are meaningful.
    Gradients are sliced back onto the holder's parameters."""

    def __init__(self, module, act="none", side_dim=0, trainable=True, in_features=None):
        self.real_out = module.out_features
        self.real_K = module.in_features - side_dim
        k_in = in_features if in_features is not None else self.real_K
        out_pad = (self.real_out + 127) // 128 * 128
        super().__init__(module, (1, 1, k_in), 1, 1, 0, out_pad, act, side_dim, trainable)

    def prepare(self, force=False):
        w, b = self.module.weight, self.module.bias
        ver = (w._version, w.data_ptr(), b._version)
        if not force and ver == self._version:
            return
        dev = w.device
        dense = torch.zeros(self.out_c, self.K, dtype=torch.float32, device=dev)
        dense[: self.real_out, : self.real_K] = w.detach()[:, : self.real_K]
        if self.w_fwd is None:
            self.w_fwd = torch.empty(self.out_c, self.K, dtype=torch.bfloat16, device=dev)
            self.w_t = torch.empty(self.K, self.out_c, dtype=torch.bfloat16, device=dev)
        L.call("cb_prepare_conv_weight", L.ptr(dense), L.ptr(self.w_fwd), L.ptr(self.w_t), self.out_c, self.K, 1, 1,
               L.stream())
        self._dense_w = dense
        bias = torch.zeros(self.out_c, dtype=torch.float32, device=dev)
        bias[: self.real_out] = b.detach()
        self.bias = bias
        self.side_w = self.w_ext = None
        if self.side_dim:
            sw = torch.zeros(self.out_c, self.side_dim, dtype=torch.float32, device=dev)
            sw[: self.real_out] = w.detach()[:, self.real_K:]
            self.side_w = sw
            if self.K % 64 == 0 and self.side_dim <= 64:
                self.w_ext = torch.zeros(self.out_c, self.K + 64, dtype=torch.bfloat16, device=dev)
                self.w_ext[:, : self.K] = self.w_fwd
                self.w_ext[:, self.K: self.K + self.side_dim] = sw.to(torch.bfloat16)
        self._version = ver

    def _wgrad_impl(self, x, dy, n, in_dtype=L.BF16, strides=None, side=None, norm=None, accumulate=False):
        self.prepare()
        w, b = self.module.weight, self.module.bias
        dev = dy.device
        dw = torch.empty(self.out_c, self.K, dtype=torch.float32, device=dev)
        db = torch.empty(self.out_c, dtype=torch.float32, device=dev)
        d = self.desc(n)
        L.call("cb_conv_wgrad", C.byref(d), L.ptr(x), L.ptr(dy), None, None, L.ptr(dw), L.ptr(db), 0.0, L.stream())
        if w.grad is None or not accumulate:
            w.grad = torch.zeros_like(w) if w.grad is None else w.grad.zero_()
        if b.grad is None or not accumulate:
            b.grad = torch.zeros_like(b) if b.grad is None else b.grad.zero_()
        w.grad[:, : self.real_K].add_(dw[: self.real_out, : self.real_K])
        b.grad.add_(db[: self.real_out])
        if self.side_dim:
            dsw = torch.empty(self.out_c, self.side_dim, dtype=torch.float32, device=dev)
            L.call("cb_side_wgrad", L.ptr(dy), L.ptr(side), L.ptr(dsw), n, self.out_c, self.side_dim, 0.0, L.stream())
            w.grad[:, self.real_K:].add_(dsw[: self.real_out])


class Im2colConvLayer(GemmLayer):
    """Any convolution without an implicit-GEMM specialisation (UniverseHead's 3x3 stride-2 pad-1 layers,
    encoders.py:7-35; the policy's Conv2dHeadModel convolutions, models/conv2d.py:67-118) on the tensor cores through
    an explicit patch matrix: ``cb_im2col`` gathers [rows, k_pad] bf16 rows (rows = samples x out_h x out_w), the row
    GEMM kernels do forward / dgrad / wgrad on it, ``cb_col2im`` gathers the data gradient back.  The patch matrix
    costs HBM traffic the implicit kernels avoid, so it is built per chunk of samples in a bounded scratch buffer and
    recomputed in the backward pass instead of being kept.

    ``in_row_c``: channels per pixel row of the input tensor when it is wider than ``in_c`` (the policy's conv2 reads
    the first 16 of the 32 channels the padded first convolution writes)."""

    SCRATCH_BYTES = 1 << 30

    def __init__(self, module, in_shape, act="none", trainable=True, in_row_c=None):
        k, s, pd = module.kernel_size[0], module.stride[0], module.padding[0]
        super().__init__(module, in_shape, k, s, pd, module.out_channels, act, 0, trainable)
        if self.out_c % 32:
            raise L.CuriosityLibError("Im2colConvLayer: out_channels must be a multiple of 32")
        self.in_row_c = in_row_c or self.in_c
        self.k_pad = (self.K + 63) // 64 * 64           # forward / dgrad: k-blocks of 64
        self.k_pad_wg = (self.K + 127) // 128 * 128     # wgrad: the patch matrix is the 128-row MMA operand
        self.P = self.out_h * self.out_w
        self._scratch = None

    def prepare(self, force=False):
        w, b = self.module.weight, self.module.bias
        ver = (w._version, w.data_ptr(), b._version)
        if not force and ver == self._version:
            return
        dev = w.device
        flat = w.detach().permute(0, 2, 3, 1).reshape(self.out_c, self.K)            # [o][kh][kw][ci]
        self.w_fwd = torch.zeros(self.out_c, self.k_pad, dtype=torch.bfloat16, device=dev)
        self.w_fwd[:, : self.K] = flat
        self.w_t = self.w_fwd.t().contiguous()                                       # [k_pad, out_c]
        self.bias = b.detach()
        self.side_w = self.w_ext = None
        self._version = ver

    def _chunks(self, n, k_pad):
        per = max(1, self.SCRATCH_BYTES // (self.P * k_pad * 2))
        return [(a, min(n, a + per)) for a in range(0, n, per)]

    def _buf(self, rows, k_pad, dev):
        need = rows * k_pad
        if self._scratch is None or self._scratch.numel() < need or self._scratch.device != dev:
            self._scratch = torch.empty(need, dtype=torch.bfloat16, device=dev)
        return self._scratch[:need]

    def _rows_desc(self, rows, k, out_c, act=L.ACT_NONE):
        """The [rows, k] x [out_c, k]^T product as a Linear layer."""
        d = L.ConvDesc()
        d.n, d.in_h, d.in_w, d.in_c = rows, 1, 1, k
        d.k_h = d.k_w = d.stride = 1
        d.pad = 0
        d.out_h = d.out_w = 1
        d.out_c = out_c
        d.in_dtype = L.BF16
        d.in_stride_n, d.in_stride_h, d.in_stride_w, d.in_stride_c = k, k, k, 1
        d.act = act
        d.engine = _engine
        return d

    def _conv_desc(self, n, in_dtype, strides):
        if strides is None:
            rc = self.in_row_c
            strides = (self.in_h * self.in_w * rc, self.in_w * rc, rc, 1)
        return self.desc(n, in_dtype, strides)

    def forward(self, x, n, in_dtype=L.BF16, strides=None, side=None, residual=None, norm=None,
                out_bf16=True, out_f32=False, twin=None, side_pad=None):
        if norm is not None or residual is not None or side is not None or twin is not None:
            raise L.CuriosityLibError("Im2colConvLayer: plain convolution only")
        self.prepare()
        dev = self.w_fwd.device
        m = n * self.P
        yb = torch.empty(m, self.out_c, dtype=torch.bfloat16, device=dev) if out_bf16 else None
        yf = torch.empty(m, self.out_c, dtype=torch.float32, device=dev) if out_f32 else None
        cd = self._conv_desc(n, in_dtype, strides)
        xp = x if isinstance(x, int) else L.ptr(x)
        st = L.stream()
        for a, b in self._chunks(n, self.k_pad):
            rows = (b - a) * self.P
            col = self._buf(rows, self.k_pad, dev)
            L.call("cb_im2col", C.byref(cd), xp, L.ptr(col), self.k_pad, a, b, st)
            e = L.Epilogue()
            e.bias = L.ptr(self.bias)
            e.out_bf16 = L.ptr(yb[a * self.P:]) if yb is not None else None
            e.out_f32 = L.ptr(yf[a * self.P:]) if yf is not None else None
            rd = self._rows_desc(rows, self.k_pad, self.out_c, self.act)
            L.call("cb_conv_fwd", C.byref(rd), L.ptr(col), L.ptr(self.w_fwd), C.byref(e), st)
        return yb, yf

    def dgrad(self, dy, n, x_saved=None, x_act="none", dx_add=None):
        self.prepare()
        dev, rc = dy.device, self.in_row_c
        alloc = torch.zeros if rc != self.in_c else torch.empty
        dx = alloc(n * self.in_h * self.in_w, rc, dtype=torch.bfloat16, device=dev)
        cd = self._conv_desc(n, L.BF16, None)
        st = L.stream()
        for a, b in self._chunks(n, self.k_pad):
            rows = (b - a) * self.P
            dcol = self._buf(rows, self.k_pad, dev)
            rd = self._rows_desc(rows, self.k_pad, self.out_c)
            L.call("cb_conv_dgrad", C.byref(rd), L.ptr(dy[a * self.P:]), L.ptr(self.w_t), None, L.ACT_NONE, None,
                   L.ptr(dcol), st)
            L.call("cb_col2im", C.byref(cd), L.ptr(dcol), self.k_pad, L.ptr(x_saved), ACT[x_act], L.ptr(dx_add),
                   L.ptr(dx), self.in_h * self.in_w * rc, self.in_w * rc, rc, a, b, st)
        return dx

    def _wgrad_impl(self, x, dy, n, in_dtype=L.BF16, strides=None, side=None, norm=None, accumulate=False):
        self.prepare()
        w, b = self.module.weight, self.module.bias
        dev = dy.device
        if w.grad is None:
            w.grad = torch.zeros_like(w)
        if b.grad is None:
            b.grad = torch.zeros_like(b)
        kp = self.k_pad_wg
        # dW^T [k_pad, out_c] = col^T dy: the patch matrix plays the 128-row operand of the split-K wgrad kernel
        dwt = torch.empty(kp, self.out_c, dtype=torch.float32, device=dev)
        cd = self._conv_desc(n, in_dtype, strides)
        xp = x if isinstance(x, int) else L.ptr(x)
        st = L.stream()
        for i, (a, c) in enumerate(self._chunks(n, kp)):
            rows = (c - a) * self.P
            col = self._buf(rows, kp, dev)
            L.call("cb_im2col", C.byref(cd), xp, L.ptr(col), kp, a, c, st)
            rd = self._rows_desc(rows, self.out_c, kp)
            L.call("cb_conv_wgrad", C.byref(rd), L.ptr(dy[a * self.P:]), L.ptr(col), None, None, L.ptr(dwt), None,
                   1.0 if i else 0.0, st)
        L.call("cb_colsum_bf16", L.ptr(dy), n * self.P, self.out_c, L.ptr(b.grad), 1.0 if accumulate else 0.0, st)
        g = dwt[: self.K].t().reshape(self.out_c, self.k, self.k, self.in_c).contiguous()     # [o][kh][kw][ci]
        L.call("cb_unprepare_conv_grad", L.ptr(g), L.ptr(w.grad), self.out_c, self.in_c, self.k, self.k,
               1.0 if accumulate else 0.0, st)


class PaddedConv1Layer(GemmLayer):
    """A first convolution 8x8 stride 4 over uint8 frames with FEWER than 32 output channels (the policy's
    Conv2dHeadModel: 4 -> 16, models/pg/atari_lstm_model.py:94-101) on the 32-channel tcgen05 first-conv kernel:
    weights and bias are zero-padded to 32 outputs, the padded channels come out as act(0) = 0 and their gradient
    rows are dropped.  Callers see 32-channel NHWC rows of which ``real_out`` are meaningful."""

    def __init__(self, module, in_shape, act="none", trainable=True):
        self.real_out = module.out_channels
        super().__init__(module, in_shape, 8, 4, 0, 32, act, 0, trainable)

    def prepare(self, force=False):
        w, b = self.module.weight, self.module.bias
        ver = (w._version, w.data_ptr(), b._version)
        if not force and ver == self._version:
            return
        dev = w.device
        self.w_native = torch.zeros(32, self.K, dtype=torch.bfloat16, device=dev)
        self.w_native[: self.real_out] = w.detach().reshape(self.real_out, -1)
        self.bias = torch.zeros(32, dtype=torch.float32, device=dev)
        self.bias[: self.real_out] = b.detach()
        self.w_fwd = self.w_native                         # (only the first-conv kernel runs this layer)
        self.side_w = self.w_ext = self.w_t = self.w_shuffle = None
        self._version = ver

    def forward(self, x, n, in_dtype=L.BF16, strides=None, **kw):
        d = self.desc(n, in_dtype, strides)
        self.prepare()
        if not self._use_conv1(d):
            raise L.CuriosityLibError("PaddedConv1Layer: needs uint8 NCHW frames (8x8 stride 4, <= 4 channels) on sm_100")
        return super().forward(x, n, in_dtype=in_dtype, strides=strides, **kw)

    def _wgrad_impl(self, x, dy, n, in_dtype=L.BF16, strides=None, side=None, norm=None, accumulate=False):
        self.prepare()
        w, b = self.module.weight, self.module.bias
        if w.grad is None:
            w.grad = torch.zeros_like(w)
        if b.grad is None:
            b.grad = torch.zeros_like(b)
        dev = dy.device
        dw = torch.empty(32, self.K, dtype=torch.float32, device=dev)
        db = torch.empty(32, dtype=torch.float32, device=dev)
        d = self.desc(n, in_dtype, strides)
        xp = x if isinstance(x, int) else L.ptr(x)
        L.call("cb_conv1_wgrad", C.byref(d), xp, L.ptr(dy), None, None, L.ptr(dw), L.ptr(db), 0.0, L.stream())
        r = self.real_out
        if accumulate:
            w.grad.add_(dw[:r].view_as(w)); b.grad.add_(db[:r])
        else:
            w.grad.copy_(dw[:r].view_as(w)); b.grad.copy_(db[:r])


class NarrowHead:
    """A Linear(K -> few outputs) head (the policy's pi / value for action sets the fused heads kernel does not cover,
    ICM's inverse-model output for A > 8: e.g. the 12 actions of SuperMarioBros' COMPLEX_MOVEMENT) on the tcgen05 GEMM:
    the outputs are zero-padded to one 128-wide tile (``PaddedLinearLayer``); callers pass and receive fp32 tensors of
    the real width.  No SIMT launch on this path."""

    def __init__(self, module, act="none"):
        self.layer = PaddedLinearLayer(module, act)
        self.real_out = module.out_features

    def prepare(self):
        self.layer.prepare()

    @property
    def _version(self):
        return self.layer._version

    @_version.setter
    def _version(self, v):
        self.layer._version = v

    def forward(self, x_bf16, n):
        _, yf = self.layer.forward(x_bf16, n, out_bf16=False, out_f32=True)
        return yf[:, : self.real_out].contiguous()

    def backward(self, x_bf16, dy_f32, n, x_saved=None, x_act="none", dx_add=None):
        """Weight / bias gradients into the holder's ``.grad``; returns d x (bf16 [n, K])."""
        dy = torch.zeros(n, self.layer.out_c, dtype=torch.bfloat16, device=dy_f32.device)
        dy[:, : self.real_out] = dy_f32.reshape(n, self.real_out)
        self.layer.wgrad(x_bf16, dy, n)
        return self.layer.dgrad(dy, n, x_saved=x_saved, x_act=x_act, dx_add=dx_add)


def narrow_head(module, act="none"):
    """``NarrowHead`` on the tensor cores, or the generic layer when the SIMT cross-check engine is forced."""
    return NarrowHead(module, act) if _engine != L.ENGINE_SIMT else linear_layer(module, act)


class BatchNormLayer:
    """nn.BatchNorm2d between two layers of an encoder ``Chain`` (rlpyt/models/curiosity/encoders.py:23-25, 87-96):
    torch's arithmetic on bf16 NHWC activations.  Training mode (``module.training``): statistics of THIS batch over
    all rows of a channel, running statistics updated as torch does (momentum, unbiased variance,
    ``num_batches_tracked``); evaluation mode: the running statistics.

    It follows the Chain protocol of ``GemmLayer``: ``forward`` -> (bf16, fp32), ``wgrad`` leaves dgamma / dbeta in the
    holder's ``.grad``, ``dgrad`` returns the gradient w.r.t. the PRE-activation output of the layer in front (the
    activation's derivative is folded into the same pass, as the conv dgrad epilogues do).  What the backward pass
    needs from the forward pass (mean, 1/sqrt(var+eps)) travels as attributes of the INPUT activation tensor, so an
    encoder that is run twice before its backward passes (ICM: obs and next_obs) keeps the two apart."""

    def __init__(self, module, shape, trainable=True):
        self.module = module
        self.in_h, self.in_w, self.in_c = shape
        self.out_h, self.out_w, self.out_c = shape
        self.k, self.stride, self.pad = 1, 1, 0
        self.act = ACT["none"]
        self.trainable = trainable
        self._version = None
        if self.in_c % 8 or 256 % (self.in_c // 8):
            raise L.CuriosityLibError(f"BatchNorm2d over {self.in_c} channels is not supported (8 * a power of two)")

    def prepare(self, force=False):
        pass

    def _rows(self, n):
        return n * self.in_h * self.in_w

    def forward(self, x, n, out_bf16=True, out_f32=False, **kw):
        m, C, rows = self.module, self.in_c, self._rows(n)
        dev = x.device
        st = L.stream()
        if m.training or m.running_mean is None:
            s = torch.zeros(2, C, dtype=torch.float64, device=dev)
            L.call("cb_bn_moments", L.ptr(x), rows, C, L.ptr(s[0]), L.ptr(s[1]), st)
            mean = s[0] / rows
            var = (s[1] / rows - mean * mean).clamp_min_(0.0)          # biased, as torch normalises with
            if m.training and m.track_running_stats and m.running_mean is not None:
                with torch.no_grad():
                    m.num_batches_tracked += 1
                    mom = m.momentum if m.momentum is not None else 1.0 / float(m.num_batches_tracked)
                    m.running_mean.mul_(1 - mom).add_((mom * mean).to(m.running_mean.dtype))
                    m.running_var.mul_(1 - mom).add_((mom * var * (rows / max(rows - 1, 1))).to(m.running_var.dtype))
            batch_stats = True
        else:
            mean, var = m.running_mean.double(), m.running_var.double()
            batch_stats = False
        invstd = (var + m.eps).rsqrt()
        gamma = m.weight.detach().double() if m.weight is not None else torch.ones(C, dtype=torch.float64, device=dev)
        beta = m.bias.detach().double() if m.bias is not None else torch.zeros(C, dtype=torch.float64, device=dev)
        scale = (gamma * invstd).float()
        shift = (beta - mean * gamma * invstd).float()
        yb = torch.empty(rows, C, dtype=torch.bfloat16, device=dev) if out_bf16 else None
        yf = torch.empty(rows, C, dtype=torch.float32, device=dev) if out_f32 else None
        L.call("cb_bn_apply", L.ptr(x), rows, C, L.ptr(scale), L.ptr(shift), L.ptr(yb), L.ptr(yf), st)
        x._bn_saved = (mean.float(), invstd.float(), batch_stats)
        x._bn_sums = None
        return yb, yf

    def _sums(self, x, dy, n):
        if getattr(x, "_bn_sums", None) is None:
            C, rows = self.in_c, self._rows(n)
            mean, invstd, _ = x._bn_saved
            s = torch.zeros(2, C, dtype=torch.float64, device=dy.device)
            L.call("cb_bn_bwd_moments", L.ptr(dy), L.ptr(x), rows, C, L.ptr(mean), L.ptr(invstd), L.ptr(s[0]), L.ptr(s[1]),
                   L.stream())
            x._bn_sums = s
        return x._bn_sums

    def wgrad(self, x, dy, n, accumulate=False, final=True, **kw):
        m = self.module
        if not self.trainable or m.weight is None:
            return
        s = self._sums(x, dy, n)
        for p, g in ((m.weight, s[1]), (m.bias, s[0])):
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            if accumulate:
                p.grad.add_(g.to(p.grad.dtype))
            else:
                p.grad.copy_(g)
        if final:
            dist.grads_ready((m.weight.grad, m.bias.grad))

    def dgrad(self, dy, n, x_saved=None, x_act="none", dx_add=None):
        assert dx_add is None
        m, C, rows = self.module, self.in_c, self._rows(n)
        x = x_saved
        mean, invstd, batch_stats = x._bn_saved
        mean, invstd = mean.double(), invstd.double()
        gamma = m.weight.detach().double() if m.weight is not None else torch.ones(C, dtype=torch.float64, device=dy.device)
        a = gamma * invstd
        if batch_stats:      # dx = gamma*invstd * (dz - mean(dz) - xhat * mean(dz*xhat)), xhat = (x - mean)*invstd
            s = self._sums(x, dy, n)
            m_dz, m_dzx = s[0] / rows, s[1] / rows
            b = -a * invstd * m_dzx
            c = -a * m_dz - b * mean
        else:
            b = torch.zeros_like(a)
            c = torch.zeros_like(a)
        dx = torch.empty(rows, C, dtype=torch.bfloat16, device=dy.device)
        af, bf, cf = a.float(), b.float(), c.float()         # (named: the pointers must outlive the argument list)
        L.call("cb_bn_bwd_apply", L.ptr(dy), L.ptr(x), rows, C, L.ptr(af), L.ptr(bf), L.ptr(cf), ACT[x_act], L.ptr(dx),
               L.stream())
        return dx


class ParamPair:
    """A (weight, bias) pair of an ``nn.LSTM`` / ``nn.GRU`` holder presented like a ``Linear`` to ``GemmLayer``."""

    def __init__(self, weight, bias):
        self.weight, self.bias = weight, bias
        self.in_features, self.out_features = weight.shape[1], weight.shape[0]


def pad_side(side):
    """fp32 [n, A] side input (one-hot action) -> bf16 [n, 64] zero padded: the extra k-block operand."""
    n, a = side.shape
    out = torch.zeros(n, 64, dtype=torch.bfloat16, device=side.device)
    out[:, :a] = side
    return out


def _has_implicit_kernel(in_shape, k, s, p, out_c):
    """Shapes the implicit-GEMM tcgen05 kernels cover (umma_conv1.cu / umma_window.cuh / umma_conv.cu)."""
    h, w, c = in_shape
    if (k, s, p, out_c) == (8, 4, 0, 32) and c <= 4:
        return True                                                   # uint8 first convolution
    if s == 2 and k == 4 and p == 0 and c == 32 and out_c in (64, 128) and h % 2 == 0 and w % 2 == 0:
        return True                                                   # Burda conv2
    if s == 1 and c == 64 and out_c in (64, 128) and k > 1 and (h + 2 * p) * (w + 2 * p) <= 256:
        return True                                                   # Burda conv3
    return False


def conv_layer(module, in_shape, act, trainable=True, in_row_c=None):
    """The layer object that runs ``module`` (a Conv2d holder): the implicit-GEMM kernels where a specialisation
    exists, else the explicit patch-matrix path -- every convolution of the path runs its MACs on tcgen05.  The SIMT
    engine is only reachable by forcing it (``set_engine('simt')``, the cross-check of the tensor-core kernels)."""
    k, s, p = module.kernel_size[0], module.stride[0], module.padding[0]
    oc = module.out_channels
    if _engine != L.ENGINE_SIMT and not _has_implicit_kernel(in_shape, k, s, p, oc):
        if (k, s, p) == (8, 4, 0) and in_shape[2] <= 4 and oc < 32:
            return PaddedConv1Layer(module, in_shape, act, trainable)
        if oc % 32 == 0:
            return Im2colConvLayer(module, in_shape, act, trainable, in_row_c=in_row_c)
    return GemmLayer(module, in_shape, k, s, p, oc, act, 0, trainable)


def linear_layer(module, act="none", side_dim=0, spatial=None, trainable=True):
    """``spatial=(k, c)``: the Linear consumes a flattened [c,k,k] map (Burda / Maze heads) and is
    run as a kxk convolution over the NHWC map; otherwise a plain [*, K] GEMM."""
    if spatial is not None:
        k, c = spatial
        return GemmLayer(module, (k, k, c), k, 1, 0, module.out_features, act, 0, trainable)
    K = module.in_features - side_dim
    return GemmLayer(module, (1, 1, K), 1, 1, 0, module.out_features, act, side_dim, trainable)


class Chain:
    """A feed-forward stack of GemmLayers (conv encoder + MLP head).

    ``forward`` keeps every layer's bf16 output (needed by wgrad / act' in ``backward``); the
    last layer can additionally emit fp32 (features entering a loss stay in fp32 so the
    squared-error / variance epilogues do not see bf16 cancellation).
    """

    ACT_NAME = {v: k for k, v in ACT.items()}

    def __init__(self, layers):
        self.layers = layers

    @property
    def last_act(self):
        """Name of the activation the chain's output has been through (whoever computes the gradient w.r.t. that output
        folds its derivative in): 'none' when the chain ends in a Linear without activation or a BatchNorm2d."""
        return self.ACT_NAME[self.layers[-1].act]

    def forward(self, x, n, in_dtype=L.BF16, strides=None, norm=None, last_f32=True,
                last_bf16=False, first_out=None):
        """``first_out``: bf16 output of layer 0 computed elsewhere (fused twin first conv), or a list with the
        outputs of the first few layers (fused first two convolutions; an entry that is not needed may be None)."""
        acts = []
        cur = x
        last = len(self.layers) - 1
        out_f32 = None
        lead = [] if first_out is None else (list(first_out) if isinstance(first_out, (list, tuple)) else [first_out])
        for i, layer in enumerate(self.layers):
            if i < len(lead):
                acts.append(lead[i])
                cur = lead[i]
                continue
            kw = dict(in_dtype=in_dtype, strides=strides, norm=norm) if i == 0 else {}
            want_f32 = last_f32 and i == last
            want_bf16 = (i != last) or last_bf16 or not last_f32
            yb, yf = layer.forward(cur, n, out_bf16=want_bf16, out_f32=want_f32, **kw)
            acts.append(yb)
            cur = yb
            if i == last:
                out_f32 = yf
        return acts, out_f32

    def backward(self, x, acts, dy, n, in_dtype=L.BF16, strides=None, norm=None, accumulate=False,
                 need_dx=False, x_saved=None, x_act="none", final=True):
        """dy: bf16 gradient w.r.t. the last layer's (pre-loss) output.  Fills ``.grad`` of all
        trainable holders; returns the gradient w.r.t. the chain input when ``need_dx``."""
        for i in range(len(self.layers) - 1, -1, -1):
            layer = self.layers[i]
            if i == 0:
                layer.wgrad(x, dy, n, in_dtype=in_dtype, strides=strides, norm=norm,
                            accumulate=accumulate, final=final)
                if need_dx:
                    return layer.dgrad(dy, n, x_saved=x_saved, x_act=x_act)
                return None
            prev = self.layers[i - 1]
            layer.wgrad(acts[i - 1], dy, n, accumulate=accumulate, final=final)
            dy = layer.dgrad(dy, n, x_saved=acts[i - 1], x_act=self.ACT_NAME[prev.act])
        return None
