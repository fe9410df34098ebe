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
        if m._enc_last_act == "relu":
            dy = dy * (y > 0)
        elif m._enc_last_act == "elu":
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
        gh = torch.empty(B, 3 * G, dtype=torch.float32, device=dev)
        hh.prepare()
        L.call("cb_gru_sequence_fwd", T, B, G, L.ptr(gates), L.ptr(hh.w_fwd), L.ptr(hh.bias), L.ptr(h_all), L.ptr(hb),
               L.ptr(ghn), L.ptr(gh), LY.get_engine(), L.stream())
        ctx.model, ctx.saved, ctx.params = model, (xb, side, gates, h_all, hb, ghn, T, B, n), params
        return h_all[1:]

    @staticmethod
    def backward(ctx, dbelief):
        ex = ctx.model._exec
        xb, side, gates, h_all, hb, ghn, T, B, n = ctx.saved
        G, dev = ctx.model.gru_size, xb.device
        ih, hh = ex["gru_ih"], ex["gru_hh"]
        dout = dbelief.to(torch.float32).contiguous()
        dgi = torch.empty(n, 3 * G, dtype=torch.bfloat16, device=dev)
        dgh = torch.empty(n, 3 * G, dtype=torch.bfloat16, device=dev)
        ws = torch.empty(3, B, G, dtype=torch.bfloat16, device=dev)
        hh.prepare()
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
    def __init__(self, image_shape, action_size, horizon, feature_encoding="idf_maze", gru_size=128, batch_norm=False,
                 obs_stats=None, num_predictors=10, device="cpu"):
        super().__init__()
        assert num_predictors >= horizon
        self.action_size, self.horizon = action_size, horizon
        self.feature_encoding, self.obs_stats, self.num_predictors = feature_encoding, obs_stats, num_predictors
        self.encoder, self.feature_size = make_encoder(feature_encoding, image_shape, batch_norm)
        self._enc_last_act = {"idf": "elu", "idf_burda": "none", "idf_maze": "relu"}[feature_encoding]
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
