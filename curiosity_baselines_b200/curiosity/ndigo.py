"""NDIGO on B200 -- drop-in for rlpyt/models/curiosity/ndigo.py:36-274 (same constructor,
``state_dict`` keys, ``compute_bonus`` / ``compute_loss`` signatures and quirks: the GRU state is
reset every call, ``valid`` and ``prediction_beta`` are ignored, horizon 10 has no predictor).

The encoder, the 128-unit GRU (x W_ih^T for all T*B rows as one tcgen05 GEMM with the one-hot
prev_action as its extra k-block, the recurrence and its back-propagation through time as ONE library call each,
``cb_gru_sequence_fwd / _bwd``), the multi-horizon predictor MLPs (the action window enters as a side input of the
first Linear instead of a concatenation) and the BCE all run in libcuriosity_b200.so; ``self.gru`` is a
``torch.nn.GRU`` parameter holder that is never called.  The pieces are tied together through
``torch.autograd.Function`` wrappers.
"""
import ctypes as C

import torch
from torch import nn

from .. import _lib as L
from .. import layers as LY
from ..layers import linear_layer, pad_side, PaddedLinearLayer
from .encoders import make_encoder, frame_input


class _EncoderFn(torch.autograd.Function):
    """features = encoder(obs); obs carries no gradient."""

    @staticmethod
    def forward(ctx, model, frames, n, *params):
        ex = model._exec
        keep, x, dtype, strides = frame_input(frames)
        acts, phi = ex["chain"].forward(x, n, in_dtype=dtype, strides=strides, last_f32=True, last_bf16=True)
        ctx.model, ctx.n = model, n
        ctx.enc = dict(keep=keep, x=x, dtype=dtype, strides=strides, acts=acts)
        ctx.params = params
        return phi.reshape(n, -1)

    @staticmethod
    def backward(ctx, dphi):
        m, n, enc = ctx.model, ctx.n, ctx.enc
        chain = m._exec["chain"]
        last = chain.layers[-1]
        # gradient w.r.t. the last layer's pre-activation output
        dy = dphi.reshape(-1, last.out_c)
        y = enc["acts"][-1].float()
        if chain.last_act == "relu":
            dy = dy * (y > 0)
        elif chain.last_act == "elu":
            dy = dy * torch.where(y > 0, torch.ones_like(y), y + 1)
        saved = [p.grad for p in ctx.params]
        for p in ctx.params:
            p.grad = None
        chain.backward(enc["x"], enc["acts"], dy.to(torch.bfloat16).contiguous(), n, in_dtype=enc["dtype"],
                       strides=enc["strides"])
        grads = [p.grad for p in ctx.params]
        for p, g in zip(ctx.params, saved):
            p.grad = g
        return (None, None, None) + tuple(grads)


class _GruFn(torch.autograd.Function):
    """belief[T,B,G] = GRU([phi, prev_action]) from a zero state (ndigo.py:79,127,145)."""

    @staticmethod
    def forward(ctx, model, phi, prev_actions, *params):
        ex = model._exec
        T, B, A = prev_actions.shape
        n, G, dev = T * B, model.gru_size, phi.device
        ih, hh = ex["gru_ih"], ex["gru_hh"]
        xb = phi.to(torch.bfloat16).contiguous()
        side = prev_actions.reshape(n, A).contiguous()
        sp = pad_side(side) if A <= 64 and xb.shape[1] % 64 == 0 else None
        _, gates = ih.forward(xb, n, side=side, side_pad=sp, out_bf16=False, out_f32=True)       # gi, [n,3G] fp32
        h_all = torch.zeros(T + 1, B, G, dtype=torch.float32, device=dev)
        hb = torch.zeros(T + 1, B, G, dtype=torch.bfloat16, device=dev)
        ghn = torch.empty(T, B, G, dtype=torch.float32, device=dev)
        hh.prepare()
        if model.persistent_gru and LY.get_engine() != L.ENGINE_SIMT and L.load().cb_gru_persistent_supported(G):
            # one persistent launch for all T steps (csrc/recurrent.cu) instead of two launches per step
            L.call("cb_gru_persistent_fwd", T, B, G, L.ptr(gates), L.ptr(hh.w_fwd), L.ptr(hh.bias), L.ptr(h_all), L.ptr(hb),
                   L.ptr(ghn), L.stream())
        else:
            gh = torch.empty(B, 3 * G, dtype=torch.float32, device=dev)
            L.call("cb_gru_sequence_fwd", T, B, G, L.ptr(gates), L.ptr(hh.w_fwd), L.ptr(hh.bias), L.ptr(h_all), L.ptr(hb),
                   L.ptr(ghn), L.ptr(gh), LY.get_engine(), L.stream())
        ctx.model, ctx.saved, ctx.params = model, (xb, side, gates, h_all, hb, ghn, T, B, n), params
        out = h_all[1:]
        model._belief_pair = (out, hb[1:])          # the recurrence also wrote h in bf16: the predictors' GEMM operand
        return out

    @staticmethod
    def backward(ctx, dbelief):
        ex = ctx.model._exec
        xb, side, gates, h_all, hb, ghn, T, B, n = ctx.saved
        G, dev = ctx.model.gru_size, xb.device
        ih, hh = ex["gru_ih"], ex["gru_hh"]
        dout = dbelief.to(torch.float32).contiguous()
        dgi = torch.empty(n, 3 * G, dtype=torch.bfloat16, device=dev)
        dgh = torch.empty(n, 3 * G, dtype=torch.bfloat16, device=dev)
        hh.prepare()
        if ctx.model.persistent_gru and LY.get_engine() != L.ENGINE_SIMT and L.load().cb_gru_persistent_supported(G):
            L.call("cb_gru_persistent_bwd", T, B, G, L.ptr(gates), L.ptr(ghn), L.ptr(h_all), L.ptr(dout), L.ptr(hh.w_t),
                   L.ptr(dgi), L.ptr(dgh), L.stream())
        else:
            ws = torch.empty(3, B, G, dtype=torch.bfloat16, device=dev)
            L.call("cb_gru_sequence_bwd", T, B, G, L.ptr(gates), L.ptr(ghn), L.ptr(h_all), L.ptr(dout), L.ptr(hh.w_t),
                   L.ptr(dgi), L.ptr(dgh), L.ptr(ws), LY.get_engine(), L.stream())
        saved = [p.grad for p in ctx.params]
        for p in ctx.params:
            p.grad = None
        hh.wgrad(hb[:T].reshape(n, G), dgh, n)
        ih.wgrad(xb, dgi, n, side=side)
        dphi = ih.dgrad(dgi, n)
        grads = [p.grad for p in ctx.params]
        for p, g in zip(ctx.params, saved):
            p.grad = g
        return (None, dphi.float(), None) + tuple(grads)


class _PredictorFn(torch.autograd.Function):
    """logits = Linear(64->D)(relu(Linear(G + kA -> 64)([belief, action window])))."""

    @staticmethod
    def forward(ctx, layers, belief, window, *params):
        l1, l2 = layers
        n = belief.shape[0]
        xb = belief.to(torch.bfloat16).contiguous()
        window = window.contiguous()
        h, _ = l1.forward(xb, n, side=window, side_pad=pad_side(window) if window.shape[1] <= 64 else None)
        _, logits = l2.forward(h, n, out_bf16=False, out_f32=True)
        ctx.layers, ctx.saved, ctx.params = layers, (xb, window, h, n), params
        real = getattr(l2, "real_out", l2.out_c)
        return logits if real == l2.out_c else logits[:, :real].contiguous()      # padded layers: drop the zero columns

    @staticmethod
    def backward(ctx, dlogits):
        l1, l2 = ctx.layers
        xb, window, h, n = ctx.saved
        saved = [p.grad for p in ctx.params]
        for p in ctx.params:
            p.grad = None
        dy = dlogits.to(torch.bfloat16).contiguous()
        if getattr(l2, "real_out", l2.out_c) != l2.out_c:
            dy = torch.nn.functional.pad(dy, (0, l2.out_c - dy.shape[1])).contiguous()
        l2.wgrad(h, dy, n)
        dh = l2.dgrad(dy, n, x_saved=h, x_act="relu")
        l1.wgrad(xb, dh, n, side=window)
        dx = l1.dgrad(dh, n)
        grads = [p.grad for p in ctx.params]
        for p, g in zip(ctx.params, saved):
            p.grad = g
        return (None, dx.float(), None) + tuple(grads)


class _HorizonStack:
    """NDIGO's ten multi-horizon predictors (ndigo.py:19-34, 82-111) as TWO grouped layers over all (t,b) rows at once
    (SURVEY 8f rank 3): the reference -- and our round-1 code -- looped over k in Python, building each action window
    with torch.cat and launching every predictor separately (~25 launches per horizon).

      layer 1   h[n, G*64]   = relu([belief | window10] W1_all^T + b1)     ONE grouped tcgen05 launch: every predictor
                reads the SAME belief rows (x_shared) and the same 10-step action window as its extra k-block; predictor
                k's weights for the window steps j >= k are zero, which is exactly its k-step window
      layer 2   z[n, G*FP]   = h_g W2_g^T + b2_g                          ONE grouped launch (block diagonal)
      loss      cb_bce_horizons: predictor k's row (t,b) against obs[t+k,b], rows with t+k >= T masked, every horizon
                in one pass; it also writes d z in the bf16 operand layout of the backward GEMMs
    Backward: grouped wgrad / dgrad of layer 2, then layer 1's weight gradient and d belief as PLAIN GEMMs -- the
    contraction over the stacked G*64 outputs is the sum over horizons."""

    F1, WIN = 64, 10

    def __init__(self, model):
        self.m = model
        self.A, self.Gs, self.D = model.action_size, model.gru_size, model.obs_dim
        self.FP = (self.D + 63) // 64 * 64
        self.supported = (self.Gs % 64 == 0 and self.WIN * self.A <= 64 and LY.get_engine() != L.ENGINE_SIMT)
        self._key = None

    def _params(self, k):
        seq = getattr(self.m, f"forward_model_{k}").model
        return seq[0].weight, seq[0].bias, seq[2].weight, seq[2].bias

    def prepare(self):
        key = tuple((p._version, p.data_ptr()) for k in range(1, 11) for p in self._params(k))
        if key == self._key:
            return
        dev = self.m.gru.weight_ih_l0.device
        G, F1, Gs, D, FP, A = 10, self.F1, self.Gs, self.D, self.FP, self.A
        w1 = torch.zeros(G * F1, Gs + 64, dtype=torch.float32, device=dev)
        b1 = torch.zeros(G * F1, dtype=torch.float32, device=dev)
        w2 = torch.zeros(G * FP, F1, dtype=torch.float32, device=dev)
        b2 = torch.zeros(G * FP, dtype=torch.float32, device=dev)
        for g in range(G):
            a, ab, c, cb_ = self._params(g + 1)
            w1[g * F1:(g + 1) * F1, : Gs + (g + 1) * A] = a.detach()
            b1[g * F1:(g + 1) * F1] = ab.detach()
            w2[g * FP: g * FP + D] = c.detach()
            b2[g * FP: g * FP + D] = cb_.detach()
        self.w1 = w1.to(torch.bfloat16).contiguous()                          # [G*64, Gs + 64]
        self.side_w = w1[:, Gs: Gs + self.WIN * A].contiguous()               # fp32 [G*64, 50] (only read on fallback paths)
        self.b1, self.b2 = b1, b2
        self.w2 = w2.to(torch.bfloat16).contiguous()                          # [G*FP, 64]
        self.w2_t = torch.cat([w2[g * FP:(g + 1) * FP].t() for g in range(G)]).to(torch.bfloat16).contiguous()   # [G*64, FP]
        self._w1_t = {}
        self._w1_f32 = w1
        self._key = key

    def w1_t(self, G):
        if G not in self._w1_t:      # [Gs, G*64]: operand of the plain d-belief GEMM
            self._w1_t[G] = self._w1_f32[: G * self.F1, : self.Gs].t().to(torch.bfloat16).contiguous()
        return self._w1_t[G]

    def forward(self, belief_bf16, onehot_actions, T, B, G):
        """-> (packed input [n, Gs+64] bf16 = [belief | 10-step action window], h [n, G*64] bf16, logits [n, G*FP] fp32)."""
        self.prepare()
        n, dev, st = T * B, belief_bf16.device, L.stream()
        A, F1, FP, Gs = self.A, self.F1, self.FP, self.Gs
        xcat = torch.empty(n, Gs + 64, dtype=torch.bfloat16, device=dev)
        xcat[:, :Gs].copy_(belief_bf16)
        L.call("cb_action_windows", L.ptr(onehot_actions), T, B, A, self.WIN, xcat.data_ptr() + 2 * Gs, Gs + 64, None, st)
        h = torch.empty(n, G * F1, dtype=torch.bfloat16, device=dev)
        e = L.Epilogue()
        e.bias, e.out_bf16 = L.ptr(self.b1), L.ptr(h)
        L.call("cb_grouped_linear_fwd", n, G, Gs + 64, F1, 1, L.ptr(xcat), L.ptr(self.w1), L.ACT_RELU, C.byref(e), st)
        z = torch.empty(n, G * FP, dtype=torch.float32, device=dev)
        e2 = L.Epilogue()
        e2.bias, e2.out_f32 = L.ptr(self.b2), L.ptr(z)
        L.call("cb_grouped_linear_fwd", n, G, F1, FP, 0, L.ptr(h), L.ptr(self.w2), L.ACT_NONE, C.byref(e2), st)
        return xcat, h, z

    def backward(self, xcat, h, dz, T, B, G):
        """dz bf16 [n, G*FP] -> (d belief bf16 [n, Gs], per-predictor parameter gradients {k: (dW1, db1, dW2, db2)})."""
        n, dev, st = T * B, h.device, L.stream()
        A, F1, FP, Gs, D = self.A, self.F1, self.FP, self.Gs, self.D
        dw2 = torch.empty(G * FP, F1, dtype=torch.float32, device=dev)
        db2 = torch.empty(G * FP, dtype=torch.float32, device=dev)
        L.call("cb_grouped_linear_wgrad", n, G, F1, FP, 0, L.ptr(h), L.ptr(dz), L.ptr(dw2), L.ptr(db2), 0.0, st)
        dh = torch.empty(n, G * F1, dtype=torch.bfloat16, device=dev)
        L.call("cb_grouped_linear_dgrad", n, G, F1, FP, L.ptr(dz), L.ptr(self.w2_t), L.ptr(h), L.ACT_RELU, None, L.ptr(dh), st)
        # layer 1: the stacked outputs make both remaining products plain GEMMs; the packed input gives the belief AND
        # the action-window weight gradients in one of them
        def desc(k):
            d = L.ConvDesc()
            d.n, d.in_h, d.in_w, d.in_c = n, 1, 1, k
            d.k_h = d.k_w = d.stride = 1
            d.out_h = d.out_w = 1
            d.out_c = G * F1
            d.in_dtype = L.BF16
            d.in_stride_n, d.in_stride_h, d.in_stride_w, d.in_stride_c = k, k, k, 1
            d.engine = LY.get_engine()
            return d
        dw1 = torch.empty(G * F1, Gs + 64, dtype=torch.float32, device=dev)
        db1 = torch.empty(G * F1, dtype=torch.float32, device=dev)
        dk = desc(Gs + 64)
        L.call("cb_conv_wgrad", C.byref(dk), L.ptr(xcat), L.ptr(dh), None, None, L.ptr(dw1), L.ptr(db1), 0.0, st)
        dbel = torch.empty(n, Gs, dtype=torch.bfloat16, device=dev)
        dg = desc(Gs)
        L.call("cb_conv_dgrad", C.byref(dg), L.ptr(dh), L.ptr(self.w1_t(G)), None, L.ACT_NONE, None, L.ptr(dbel), st)
        grads = {}
        for g in range(G):
            k = g + 1
            grads[k] = (dw1[g * F1:(g + 1) * F1, : Gs + k * A], db1[g * F1:(g + 1) * F1], dw2[g * FP: g * FP + D],
                        db2[g * FP: g * FP + D])
        return dbel, grads


class _HorizonLossFn(torch.autograd.Function):
    """sum_{k <= K} mean BCE(sigmoid(M_k(belief[:T-k], window_k)), obs[k:])   (ndigo.py:234-268) on the grouped stack."""

    @staticmethod
    def forward(ctx, model, belief, onehot_actions, obs_flat, K, *params):
        hs = model._horizons
        T, B = onehot_actions.shape[:2]
        n = T * B
        G = (K + 1) // 2 * 2                               # layer 1's weight gradient tiles 128 stacked outputs
        xb = model._belief_bf16(belief, n)
        xcat, h, z = hs.forward(xb, onehot_actions.contiguous(), T, B, G)
        rows = torch.empty(G, n, dtype=torch.float32, device=xb.device)
        need = any(ctx.needs_input_grad)
        dz = torch.empty(n, G * hs.FP, dtype=torch.bfloat16, device=xb.device) if need else None
        L.call("cb_bce_horizons", L.ptr(z), L.ptr(obs_flat), T, B, hs.D, G, hs.FP, 1, None, L.ptr(rows), 1.0, L.ptr(dz),
               L.stream())
        if need and G > K:
            dz.view(n, G, hs.FP)[:, K:].zero_()            # the padding horizon is not part of the loss
        ctx.saved = (xcat, h, dz, T, B, G, K)
        ctx.model = model
        return rows[:K].sum()

    @staticmethod
    def backward(ctx, g):
        xcat, h, dz, T, B, G, K = ctx.saved
        hs = ctx.model._horizons
        dbel, grads = hs.backward(xcat, h, dz, T, B, G)
        out = []
        for k in range(1, 11):
            out += [t * g for t in grads[k]] if k <= K else [None] * 4
        return (None, (dbel.float() * g).reshape(T, B, -1), None, None, None) + tuple(out)


class _BceMeanFn(torch.autograd.Function):
    """mean over all elements of binary_cross_entropy(sigmoid(logits), target)."""

    @staticmethod
    def forward(ctx, logits, target):
        m, f = logits.shape
        rows = torch.empty(m, dtype=torch.float32, device=logits.device)
        dz = torch.empty_like(logits)
        L.call("cb_bce_logits", L.ptr(logits.contiguous()), L.ptr(target.contiguous()), 1.0 / (m * f), L.ptr(rows),
               1.0 / (m * f), L.ptr(dz), m, f, L.stream())
        ctx.save_for_backward(dz)
        return rows.sum()

    @staticmethod
    def backward(ctx, g):
        (dz,) = ctx.saved_tensors
        return dz * g, None


class NdigoForward(nn.Module):
    """ndigo.py:17-34 (holder)."""

    def __init__(self, feature_size, action_size, output_size):
        super().__init__()
        self.model = nn.Sequential(nn.Linear(feature_size + action_size, 64), nn.ReLU(), nn.Linear(64, output_size))


class NDIGO(nn.Module):
    persistent_gru = True        # the recurrence as one persistent launch when the hidden size is specialised (128)

    def __init__(self, image_shape, action_size, horizon, feature_encoding="idf_maze", gru_size=128, batch_norm=False,
                 obs_stats=None, num_predictors=10, device="cpu"):
        super().__init__()
        assert num_predictors >= horizon
        self.action_size, self.horizon = action_size, horizon
        self.feature_encoding, self.obs_stats, self.num_predictors = feature_encoding, obs_stats, num_predictors
        self.encoder, self.feature_size = make_encoder(feature_encoding, image_shape, batch_norm)
        self.gru_size = gru_size
        self.gru = nn.GRU(self.feature_size + action_size, gru_size)
        self.gru_states = None
        self.obs_dim = image_shape[0] * image_shape[1] * image_shape[2]
        for k in range(1, 11):          # ndigo.py:82-111: always ten predictors
            setattr(self, f"forward_model_{k}", NdigoForward(gru_size, action_size * k, self.obs_dim))
        self._exec = None

    def _build(self):
        dev = self.gru.weight_ih_l0.device
        if dev.type != "cuda":
            raise L.CuriosityLibError("NDIGO parameters must be on a CUDA device (no CPU fallback)")
        if self._exec is None or self._exec["dev"] != dev:
            chain, perm = self.encoder.build_chain(True)
            preds = {}
            for k in range(1, 11):
                seq = getattr(self, f"forward_model_{k}").model
                if LY.get_engine() != L.ENGINE_SIMT and self.gru_size % 64 == 0:
                    # 128+kA -> 64 -> obs_dim on the tcgen05 kernels with zero-padded widths (64 -> 128, obs_dim -> 128k)
                    l1 = PaddedLinearLayer(seq[0], "relu", side_dim=k * self.action_size)
                    preds[k] = (l1, PaddedLinearLayer(seq[2], "none", in_features=l1.out_c))
                else:
                    preds[k] = (linear_layer(seq[0], "relu", side_dim=k * self.action_size), linear_layer(seq[2], "none"))
            g = self.gru
            self._horizons = _HorizonStack(self)
            self._exec = dict(dev=dev, chain=chain, perm=None if perm is None else perm.to(dev), preds=preds,
                              gru_ih=linear_layer(LY.ParamPair(g.weight_ih_l0, g.bias_ih_l0), "none", self.action_size),
                              gru_hh=linear_layer(LY.ParamPair(g.weight_hh_l0, g.bias_hh_l0), "none"))
        return self._exec

    # ------------------------------------------------------------------
    def _beliefs(self, observations, prev_actions):
        ex = self._build()
        T, B = observations.shape[:2]
        n = T * B
        phi = _EncoderFn.apply(self, observations.reshape(n, *observations.shape[2:]), n, *self.encoder.parameters())
        if ex["perm"] is not None:
            phi = phi.index_select(1, ex["perm"])
        prev = prev_actions.to(ex["dev"], torch.float32).reshape(T, B, -1)
        g = self.gru                                        # zero state every call (ndigo.py:127,145)
        return _GruFn.apply(self, phi, prev, g.weight_ih_l0, g.weight_hh_l0, g.bias_ih_l0, g.bias_hh_l0)

    def _belief_bf16(self, belief, n):
        """bf16 rows of ``belief`` [T,B,G]: the copy the GRU kernel wrote alongside the fp32 one when it is that tensor."""
        pair = getattr(self, "_belief_pair", None)
        if pair is not None and pair[0].data_ptr() == belief.data_ptr() and pair[0].shape == belief.shape:
            return pair[1].reshape(n, -1)
        return belief.reshape(n, -1).to(torch.bfloat16).contiguous()

    def _logits(self, belief, actions, k):
        """predictor k on belief[:T-k] with the k-step action window starting at each step."""
        T, B = actions.shape[:2]
        window = torch.cat([actions[j:T - k + j] for j in range(k)], dim=2)            # ndigo.py:154-163
        l1, l2 = self._exec["preds"][k]
        params = list(getattr(self, f"forward_model_{k}").parameters())
        return _PredictorFn.apply((l1, l2), belief[:T - k].reshape((T - k) * B, -1),
                                  window.reshape((T - k) * B, -1).detach(), *params)

    @torch.no_grad()
    def compute_bonus(self, observations, prev_actions, actions):
        """ndigo.py:132-217: r[H:T-1] = L_{t-1} - L_t[1:], zero elsewhere."""
        self._build()
        dev = self._exec["dev"]
        if prev_actions.dim() == 1:
            prev_actions = prev_actions.view(1, 1, -1)
        if actions.dim() == 1:
            actions = actions.view(1, 1, -1)
        T, B = observations.shape[:2]
        H = self.horizon
        if H + 1 > 10:
            raise AttributeError("'NDIGO' object has no attribute 'forward_model_11'")    # ndigo.py:194
        actions = actions.to(dev, torch.float32)
        belief = self._beliefs(observations, prev_actions)
        flat = observations.to(dev, torch.float32).reshape(T, B, -1)

        # ndigo.py:194-206: M_H(bel[:T-H]) vs obs[H:], M_{H+1}(bel[:T-H-1]) vs obs[H:-1]
        hs = self._horizons
        if hs.supported and T > H + 1:
            n, G = T * B, (H + 2) // 2 * 2
            xb = self._belief_bf16(belief, n)
            _, _, z = hs.forward(xb, actions.reshape(T, B, -1).contiguous(), T, B, G)
            rows = torch.empty(G, n, dtype=torch.float32, device=dev)
            # predictor H+1 starts one step earlier and is scored against the SAME frame as predictor H (ndigo.py:199-206)
            offs = (C.c_int32 * G)(*[H if g == H else g + 1 for g in range(G)])
            L.call("cb_bce_horizons", L.ptr(z), L.ptr(flat.contiguous()), T, B, hs.D, G, hs.FP, 0, offs, L.ptr(rows), 0.0,
                   None, L.stream())
            rows = rows.view(G, T, B)
            loss_t, loss_tm1 = rows[H - 1, : T - H], rows[H, : T - H - 1]
        else:
            loss_t = _row_losses(self, belief, actions, H, flat[H:], dev, B)
            loss_tm1 = _row_losses(self, belief, actions, H + 1, flat[H:-1], dev, B)
        r = torch.zeros(T, B, dtype=torch.float32, device=dev)
        r[H:T - 1] = loss_tm1 - loss_t[1:]
        return r

    def compute_loss(self, observations, prev_actions, actions, valid):
        """ndigo.py:220-274: sum over k of the mean BCE of predictor k; ``valid`` is ignored."""
        self._build()
        dev = self._exec["dev"]
        if prev_actions.dim() == 2:
            prev_actions = prev_actions.unsqueeze(1)
        if actions.dim() == 2:
            actions = actions.unsqueeze(1)
        T, B = observations.shape[:2]
        actions = actions.to(dev, torch.float32)
        belief = self._beliefs(observations, prev_actions)
        flat = observations.to(dev, torch.float32).reshape(T, B, -1)
        hs = self._horizons
        if hs.supported and T > self.num_predictors:
            params = [p for k in range(1, 11) for p in hs._params(k)]
            return _HorizonLossFn.apply(self, belief, actions.reshape(T, B, -1), flat.contiguous(), self.num_predictors,
                                        *params)
        loss = None
        for k in range(1, self.num_predictors + 1):
            logits = self._logits(belief, actions, k)
            term = _BceMeanFn.apply(logits, flat[k:].reshape(logits.shape[0], -1))
            loss = term if loss is None else loss + term
        return loss


def _row_losses(model, belief, actions, k, target, dev, B):
    """Per-(t,b) mean BCE of predictor k against ``target`` ([T-k, B, D])."""
    logits = model._logits(belief, actions, k)
    m, f = logits.shape
    rows = torch.empty(m, dtype=torch.float32, device=dev)
    L.call("cb_bce_logits", L.ptr(logits), L.ptr(target.reshape(m, f).contiguous()), 1.0 / f, L.ptr(rows), 0.0, None, m,
           f, L.stream())
    return rows.reshape(-1, B)
