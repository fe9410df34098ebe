"""The recurrent policy network on B200 -- drop-in for rlpyt/models/pg/atari_lstm_model.py:19-154
(SURVEY 8a row a22 / 8f rank 1): same constructor, ``state_dict`` keys (``conv.model.*``,
``lstm.weight_ih_l0`` ..., ``pi.*``, ``value.*``, ``curiosity_model.*``) and ``forward(image, prev_action,
prev_reward, init_rnn_state) -> (pi, v, RnnState(h, c))``.  The ``torch.nn`` sub-modules are parameter
holders; everything runs in libcuriosity_b200.so:

  conv encoder ........ the curiosity encoders' ``Chain`` (tcgen05 implicit GEMMs, uint8 frames read in place;
                        raw 0..255 pixels like the reference, atari_lstm_model.py:130-133)
  LSTM(F+A -> H) ...... x W_ih^T for all T*B rows as ONE tcgen05 GEMM (the one-hot prev_action is its extra
                        k-block, like ``torch.cat`` at :136-139); the recurrence is ONE library call
                        (``cb_lstm_sequence_fwd``: per timestep h W_hh^T on the same GEMM kernel + the fused cell
                        kernel, looped in C++); backward through time likewise (``cb_lstm_sequence_bwd``: cell
                        backward + recurrent dgrad GEMM per step), then both weight gradients as single GEMMs
                        over all T*B rows
  pi / value heads .... ``cb_policy_heads_fwd/_bwd``: both heads + softmax in one HBM-bound pass over the LSTM
                        output (A + 1 <= 8; larger action sets: generic Linear layers + ``cb_softmax_fwd``)

There is no CPU fallback.
"""
from collections import namedtuple

import torch
from torch import nn

from .. import _lib as L
from .. import dist
from .. import layers as LY
from ..layers import linear_layer, pad_side
from ..curiosity.encoders import BurdaHead, UniverseHead, MazeHead, frame_input
from ..curiosity.icm import ICM, Disagreement
from ..curiosity.rnd import RND
from ..curiosity.ndigo import NDIGO

RnnState = namedtuple("RnnState", ["h", "c"])


class _Holder(nn.Module):
    pass


class Conv2dHeadModel(nn.Module):
    """models/conv2d.py:67-118 as used by the policy when the curiosity model brings no encoder of its own
    (``feature_encoding='none'``: RND, or no curiosity at all; atari_lstm_model.py:94-112): Conv2d(c -> 16, k8 s4),
    ReLU, Conv2d(16 -> 32, k4 s2 p1), ReLU, flatten (CHW), Linear(-> fc_sizes), ReLU.  Parameter holder with the
    reference's ``state_dict`` keys (``conv.conv.{0,2}.*``, ``head.model.0.*``); the layers run in the library:
    the first convolution on the uint8 tcgen05 kernel (outputs zero-padded to 32 channels), the second through the
    explicit patch matrix (layers.Im2colConvLayer), the Linear as a k x k convolution over the NHWC map."""

    def __init__(self, image_shape, channels, kernel_sizes, strides, hidden_sizes, paddings=None, use_maxpool=False):
        super().__init__()
        if use_maxpool:
            raise NotImplementedError("use_maxpool=True is outside the hot-path scope (launcher default is False)")
        if isinstance(hidden_sizes, (list, tuple)):
            if len(hidden_sizes) != 1:
                raise NotImplementedError("one hidden layer after the convolutions (launcher default fc_sizes=512)")
            hidden_sizes = hidden_sizes[0]
        paddings = paddings or [0] * len(channels)
        c, h, w = image_shape
        self.image_shape = image_shape
        seq, ic = [], c
        for oc, k, s, p in zip(channels, kernel_sizes, strides, paddings):
            seq += [nn.Conv2d(ic, oc, kernel_size=k, stride=s, padding=p), nn.ReLU()]
            h, w = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
            ic = oc
        self.conv = _Holder()
        self.conv.conv = nn.Sequential(*seq)
        self.head = _Holder()
        self.head.model = nn.Sequential(nn.Linear(h * w * ic, hidden_sizes), nn.ReLU())
        self._map = (h, w, ic)
        self.output_size = hidden_sizes

    def build_chain(self, trainable=True, last_act="relu"):
        c, h, w = self.image_shape
        layers, shape, row_c = [], (h, w, c), None
        convs = [m for m in self.conv.conv if isinstance(m, nn.Conv2d)]
        for m in convs:
            lay = LY.conv_layer(m, shape, "relu", trainable, in_row_c=row_c)
            layers.append(lay)
            shape = (lay.out_h, lay.out_w, getattr(lay, "real_out", lay.out_c))
            row_c = lay.out_c if lay.out_c != shape[2] else None       # padded rows: the next layer reads a prefix
        if row_c is not None:
            raise NotImplementedError("the last convolution must write dense rows")
        mh, mw, mc = self._map
        if mh != mw:
            raise NotImplementedError("square feature maps only")
        layers.append(LY.linear_layer(self.head.model[0], "relu", spatial=(mh, mc), trainable=trainable))
        return LY.Chain(layers), None


def _lead_dims(image):
    """utils/tensor.py:49-68 for 3 trailing image dims."""
    lead = image.dim() - 3
    if lead not in (0, 1, 2):
        raise ValueError("image must be [C,H,W], [B,C,H,W] or [T,B,C,H,W]")
    if lead == 2:
        T, B = image.shape[:2]
    else:
        T, B = 1, (1 if lead == 0 else image.shape[0])
    return lead, T, B


class _PolicyFn(torch.autograd.Function):
    """Autograd boundary for the reference's PPO.loss (ppo.py:192-242): forward runs the library and keeps its
    activations; backward turns (d pi, d v) into parameter gradients with the library's backward pass."""

    @staticmethod
    def forward(ctx, model, args, *params):
        image, prev_action, init_rnn_state, need = args
        st = model._run_forward(image, prev_action, init_rnn_state, save=need)
        ctx.model, ctx.st, ctx.n_params = model, st if need else None, len(params)
        T, B, A = st["T"], st["B"], model.action_size
        ctx.mark_non_differentiable(st["hn"], st["cn"])
        return st["prob"].view(T, B, A), st["value"].view(T, B), st["hn"], st["cn"]

    @staticmethod
    def backward(ctx, d_pi, d_v, _dh, _dc):
        model, st = ctx.model, ctx.st
        n, A = st["n"], model.action_size
        dev = st["prob"].device
        dlogits = torch.empty(n, A, dtype=torch.float32, device=dev)
        d_pi = (torch.zeros(n, A, device=dev) if d_pi is None else d_pi.reshape(n, A).float()).contiguous()
        d_v = (torch.zeros(n, device=dev) if d_v is None else d_v.reshape(n).float()).contiguous()
        L.call("cb_softmax_bwd", L.ptr(st["prob"]), L.ptr(d_pi), L.ptr(dlogits), n, A, L.stream())
        params = list(model.policy_parameters())
        saved = [p.grad for p in params]
        for p in params:
            p.grad = None
        model._run_backward(st, dlogits, d_v)
        grads = {}
        for p, old in zip(params, saved):
            grads[id(p)] = p.grad
            p.grad = old
        out = [grads.get(id(p)) for p in model._fn_params]
        return (None, None) + tuple(out)


class AtariLstmModel(nn.Module):
    """atari_lstm_model.py:19-154."""

    rnn_state_cls = RnnState        # integration.install(policy=True) swaps in rlpyt's namedarraytuple of the same name

    def __init__(self, image_shape, output_size, fc_sizes=512, lstm_size=512, use_maxpool=False, channels=None,
                 kernel_sizes=None, strides=None, paddings=None, curiosity_kwargs=dict(curiosity_alg="none"),
                 obs_stats=None):
        super().__init__()
        self.obs_stats = obs_stats
        if obs_stats is not None:
            raise NotImplementedError("obs_stats normalisation of the policy input is outside the hot-path scope "
                                      "(the launcher passes None)")
        ck = dict(curiosity_kwargs)
        alg = ck.get("curiosity_alg", "none")
        self.action_size = output_size
        if alg == "icm":
            self.curiosity_model = ICM(image_shape=image_shape, action_size=output_size,
                                       feature_encoding=ck["feature_encoding"], batch_norm=ck["batch_norm"],
                                       prediction_beta=ck["prediction_beta"], obs_stats=obs_stats,
                                       forward_loss_wt=ck["forward_loss_wt"])
        elif alg == "disagreement":
            self.curiosity_model = Disagreement(image_shape=image_shape, action_size=output_size,
                                                feature_encoding=ck["feature_encoding"], batch_norm=ck["batch_norm"],
                                                prediction_beta=ck["prediction_beta"], obs_stats=obs_stats,
                                                device=ck.get("device", "cpu"), forward_loss_wt=ck["forward_loss_wt"])
        elif alg == "ndigo":
            self.curiosity_model = NDIGO(image_shape=image_shape, action_size=output_size, obs_stats=obs_stats,
                                         horizon=ck["pred_horizon"], feature_encoding=ck["feature_encoding"],
                                         batch_norm=ck["batch_norm"], num_predictors=ck["num_predictors"],
                                         device=ck.get("device", "cpu"))
        elif alg == "rnd":
            self.curiosity_model = RND(image_shape=image_shape, prediction_beta=ck["prediction_beta"],
                                       drop_probability=ck["drop_probability"], gamma=ck["gamma"],
                                       device=ck.get("device", "cpu"))
        elif alg != "none":
            raise ValueError(f"unknown curiosity_alg {alg!r}")
        enc = ck.get("feature_encoding", "none") if alg != "none" else "none"
        bn = ck.get("batch_norm", False)
        if enc == "idf":                                           # atari_lstm_model.py:80-83
            self.conv = UniverseHead(image_shape=image_shape, batch_norm=bn)
        elif enc == "idf_burda":                                   # :84-88
            self.conv = BurdaHead(image_shape=image_shape, output_size=self.curiosity_model.feature_size,
                                  batch_norm=bn)
        elif enc == "idf_maze":                                    # :89-93
            self.conv = MazeHead(image_shape=image_shape, output_size=self.curiosity_model.feature_size,
                                 batch_norm=bn)
        else:                                                      # :94-112 ('none': RND, or no curiosity model)
            self.conv = Conv2dHeadModel(image_shape=image_shape, channels=channels or [16, 32],
                                        kernel_sizes=kernel_sizes or [8, 4], strides=strides or [4, 2],
                                        paddings=paddings or [0, 1], use_maxpool=use_maxpool, hidden_sizes=fc_sizes)
            enc = "none"
        if enc != "none":
            self.conv.output_size = self.curiosity_model.feature_size
        self.feature_encoding = enc
        self.lstm_size = lstm_size
        self.lstm = nn.LSTM(self.conv.output_size + output_size, lstm_size)
        self.pi = nn.Linear(lstm_size, output_size)
        self.value = nn.Linear(lstm_size, 1)
        self._exec = None

    # ------------------------------------------------------------------ plumbing
    def policy_parameters(self):
        """Parameters of the policy proper (everything but ``curiosity_model``)."""
        for name, p in self.named_parameters():
            if not name.startswith("curiosity_model."):
                yield p

    def _build(self):
        dev = self.pi.weight.device
        if dev.type != "cuda":
            raise L.CuriosityLibError("policy parameters must be on a CUDA device (no CPU fallback)")
        if self._exec is None or self._exec["dev"] != dev:
            chain, perm = self.conv.build_chain(True)
            ls = self.lstm
            self._exec = dict(
                dev=dev, chain=chain, perm=None if perm is None else perm.to(dev),
                inv_perm=None if perm is None else torch.argsort(perm).to(dev),
                ih=linear_layer(LY.ParamPair(ls.weight_ih_l0, ls.bias_ih_l0), "none", self.action_size),
                hh=linear_layer(LY.ParamPair(ls.weight_hh_l0, ls.bias_hh_l0), "none"),
                pi=LY.narrow_head(self.pi), value=LY.narrow_head(self.value))
        return self._exec

    def invalidate_weights(self):
        """Call after the fp32 master weights changed outside torch's version tracking (FlatAdam)."""
        if self._exec is not None:
            for layer in list(self._exec["chain"].layers) + [self._exec[k] for k in ("ih", "hh", "pi", "value")]:
                layer._version = None
        cm = getattr(self, "curiosity_model", None)
        if cm is not None and hasattr(cm, "invalidate_weights"):
            cm.invalidate_weights()

    # ------------------------------------------------------------------ forward
    def _run_forward(self, image, prev_action, init_rnn_state, save, twin=None, twin_steps=None):
        """image uint8/float [T,B,C,H,W] on the device, prev_action one-hot [T,B,A], init_rnn_state (h, c) each
        [1,B,H] or None.  Returns a dict with logits [n,A], prob [n,A], value [n], hn / cn [1,B,H] and, when
        ``save``, everything the backward pass needs.  ``twin``: the first-conv layer of another network that reads the
        same frames (the curiosity encoder); it is computed in the same conv1 launch and returned as ``twin_out``."""
        ex = self._build()
        dev, H, A = ex["dev"], self.lstm_size, self.action_size
        if not image.is_cuda:
            raise L.CuriosityLibError("policy inputs must be on a CUDA device (no CPU fallback)")
        T, B = image.shape[:2]
        n = T * B
        st = L.stream()
        keep, x, dtype, strides = frame_input(image.reshape(n, *image.shape[2:]))
        first_out = twin_out = None
        l0 = ex["chain"].layers[0]
        if (twin is not None and dtype == L.U8 and (l0.k, l0.stride, l0.pad, l0.out_c, l0.in_c) ==
                (twin.k, twin.stride, twin.pad, twin.out_c, twin.in_c) and (l0.in_h, l0.in_w) == (twin.in_h, twin.in_w)):
            # ``twin_steps`` = T + 1: the frames tensor is a view of a T+1-step rollout buffer (icm.py ``shifted``) and the
            # twin network wants every step; this network keeps the first T steps of its own output
            n_tw = (twin_steps or T) * B
            first_out, _, twin_out = l0.forward(x, n_tw, in_dtype=dtype, strides=strides, twin=twin)
            first_out = first_out[: n * l0.out_h * l0.out_w]
        acts, _ = ex["chain"].forward(x, n, in_dtype=dtype, strides=strides, last_f32=False, last_bf16=True,
                                      first_out=first_out)
        feat = acts[-1].reshape(n, -1)
        if ex["perm"] is not None:                       # NHWC -> the reference's CHW feature order
            feat = feat.index_select(1, ex["perm"])
        side = prev_action.to(dev, torch.float32).reshape(n, A).contiguous()
        sp = pad_side(side) if A <= 64 and feat.shape[1] % 64 == 0 else None
        _, gates = ex["ih"].forward(feat, n, side=side, side_pad=sp, out_bf16=False, out_f32=True)   # [n,4H] fp32
        gates = gates.view(T, B, 4 * H)
        h_all = torch.empty(T + 1, B, H, dtype=torch.bfloat16, device=dev)
        c_all = torch.empty(T + 1, B, H, dtype=torch.float32, device=dev)
        if init_rnn_state is None:
            h_all[0].zero_()
            c_all[0].zero_()
        else:
            h0, c0 = init_rnn_state
            h_all[0].copy_(h0.to(dev).reshape(B, H))
            c_all[0].copy_(c0.to(dev).reshape(B, H))
        hn = torch.empty(1, B, H, dtype=torch.float32, device=dev)
        self._lstm_forward(gates, h_all, c_all, hn)
        h_out = h_all[1:].reshape(n, H)
        prob = torch.empty(n, A, dtype=torch.float32, device=dev)
        if self._fused_heads():
            logits = torch.empty(n, A, dtype=torch.float32, device=dev)
            value = torch.empty(n, dtype=torch.float32, device=dev)
            L.call("cb_policy_heads_fwd", L.ptr(h_out), L.ptr(self.pi.weight), L.ptr(self.pi.bias),
                   L.ptr(self.value.weight), L.ptr(self.value.bias), L.ptr(logits), L.ptr(value), L.ptr(prob), n, H, A, st)
        else:
            if isinstance(ex["pi"], LY.NarrowHead):          # A + 1 > 8: padded tcgen05 GEMMs
                logits, value = ex["pi"].forward(h_out, n), ex["value"].forward(h_out, n)
            else:
                _, logits = ex["pi"].forward(h_out, n, out_bf16=False, out_f32=True)
                _, value = ex["value"].forward(h_out, n, out_bf16=False, out_f32=True)
            L.call("cb_softmax_fwd", L.ptr(logits), L.ptr(prob), n, A, st)
        out = dict(T=T, B=B, n=n, logits=logits, prob=prob, value=value.reshape(n), hn=hn,
                   cn=c_all[T].clone().unsqueeze(0), twin_out=twin_out)
        if save:
            out.update(keep=keep, x=x, dtype=dtype, strides=strides, acts=acts, feat=feat, side=side, gates=gates,
                       h_all=h_all, c_all=c_all)
        return out

    def _fused_heads(self):
        return bool(L.load().cb_policy_heads_supported(self.lstm_size, self.action_size + 1))

    def _heads_backward(self, h_out, dlogits, dvalue, n):
        """Gradients of the pi / value heads (into ``.grad``) and the gradient arriving at the LSTM output (bf16)."""
        ex, st = self._exec, L.stream()
        dev, H, A = ex["dev"], self.lstm_size, self.action_size
        if self._fused_heads():
            for prm in (self.pi.weight, self.pi.bias, self.value.weight, self.value.bias):
                if prm.grad is None:
                    prm.grad = torch.zeros_like(prm)
            dh = torch.empty(n, H, dtype=torch.bfloat16, device=dev)
            dlogits, dvalue = dlogits.contiguous(), dvalue.contiguous()
            L.call("cb_policy_heads_bwd", L.ptr(h_out), L.ptr(dlogits), L.ptr(dvalue), L.ptr(self.pi.weight),
                   L.ptr(self.value.weight), L.ptr(dh), L.ptr(self.pi.weight.grad), L.ptr(self.pi.bias.grad),
                   L.ptr(self.value.weight.grad), L.ptr(self.value.bias.grad), n, H, A, L.ACT_NONE, st)
            dist.grads_ready((self.pi.weight.grad, self.pi.bias.grad, self.value.weight.grad, self.value.bias.grad))
            return dh
        if isinstance(ex["pi"], LY.NarrowHead):
            dh = ex["pi"].backward(h_out, dlogits, n)
            return ex["value"].backward(h_out, dvalue.reshape(n, 1), n, dx_add=dh)
        dlog_b = torch.empty(n, A, dtype=torch.bfloat16, device=dev)
        dval_b = torch.empty(n, 1, dtype=torch.bfloat16, device=dev)
        L.call("cb_add_to_bf16", L.ptr(dlogits), None, L.ptr(dlog_b), n * A, st)
        L.call("cb_add_to_bf16", L.ptr(dvalue), None, L.ptr(dval_b), n, st)
        ex["pi"].wgrad(h_out, dlog_b, n)
        ex["value"].wgrad(h_out, dval_b, n)
        dh = ex["pi"].dgrad(dlog_b, n)
        return ex["value"].dgrad(dval_b, n, dx_add=dh)

    def _lstm_forward(self, gates, h_all, c_all, hn):
        """The recurrence over T.  gates fp32 [T,B,4H] holds x W_ih^T + b_ih on entry and the activated gates on
        exit; h_all bf16 / c_all fp32 [T+1,B,H] with row 0 = the initial state; hn fp32 [1,B,H] or None."""
        ex, st = self._exec, L.stream()
        T, B, H = gates.shape[0], gates.shape[1], self.lstm_size
        hh = ex["hh"]
        hh.prepare()
        gh = torch.empty(B, 4 * H, dtype=torch.float32, device=gates.device)
        L.call("cb_lstm_sequence_fwd", T, B, H, L.ptr(gates), L.ptr(hh.w_fwd), L.ptr(hh.bias), L.ptr(h_all), L.ptr(c_all),
               L.ptr(gh), L.ptr(hn), LY.get_engine(), st)

    def _lstm_backward(self, gates, h_all, c_all, dh_out, dgates):
        """Back-propagation through time.  dh_out bf16 [T,B,H]: gradient arriving at every h_t from the heads;
        dgates bf16 [T,B,4H] receives the gradient w.r.t. the gate pre-activations."""
        ex, st = self._exec, L.stream()
        T, B, H = gates.shape[0], gates.shape[1], self.lstm_size
        hh = ex["hh"]
        hh.prepare()
        dc = torch.empty(2, B, H, dtype=torch.float32, device=gates.device)
        dh = torch.empty(2, B, H, dtype=torch.bfloat16, device=gates.device)
        L.call("cb_lstm_sequence_bwd", T, B, H, L.ptr(gates), L.ptr(c_all), L.ptr(dh_out), L.ptr(hh.w_t), L.ptr(dgates),
               L.ptr(dc), L.ptr(dh), LY.get_engine(), st)

    def _run_backward(self, s, dlogits, dvalue):
        """dlogits fp32 [n,A], dvalue fp32 [n]: gradients of the loss w.r.t. the pi logits and the value.
        Leaves the gradient of every policy parameter in its ``.grad``."""
        ex = self._exec
        dev, H, A = ex["dev"], self.lstm_size, self.action_size
        T, B, n = s["T"], s["B"], s["n"]
        st = L.stream()
        h_all, c_all, gates = s["h_all"], s["c_all"], s["gates"]
        h_out = h_all[1:].reshape(n, H)
        dh_out = self._heads_backward(h_out, dlogits, dvalue, n).view(T, B, H)      # bf16 [T,B,H]
        # ---- back through time
        dgates = torch.empty(T, B, 4 * H, dtype=torch.bfloat16, device=dev)
        self._lstm_backward(gates, h_all, c_all, dh_out, dgates)
        dg = dgates.view(n, 4 * H)
        # ---- weight gradients as single GEMMs over all T*B rows
        ex["hh"].wgrad(h_all[:T].reshape(n, H), dg, n)
        ex["ih"].wgrad(s["feat"], dg, n, side=s["side"])
        dfeat = ex["ih"].dgrad(dg, n, x_saved=s["feat"], x_act=ex["chain"].last_act)
        if ex["inv_perm"] is not None:
            dfeat = dfeat.index_select(1, ex["inv_perm"])
        last = ex["chain"].layers[-1]
        dfeat = dfeat.reshape(n * last.out_h * last.out_w, last.out_c).contiguous()
        ex["chain"].backward(s["x"], s["acts"], dfeat, n, in_dtype=s["dtype"], strides=s["strides"])

    def forward(self, image, prev_action, prev_reward, init_rnn_state):
        """atari_lstm_model.py:117-154.  Leading dims [T,B], [B] or []; results keep them.  Autograd flows to the
        policy parameters through ``_PolicyFn``."""
        lead, T, B = _lead_dims(image)
        dev = self.pi.weight.device
        image = image.to(dev).reshape(T, B, *image.shape[lead:])
        prev_action = prev_action.to(dev).reshape(T, B, -1)
        if init_rnn_state is not None:
            init_rnn_state = tuple(init_rnn_state)
        self._fn_params = list(self.policy_parameters())
        need = torch.is_grad_enabled() and any(p.requires_grad for p in self._fn_params)
        pi, v, hn, cn = _PolicyFn.apply(self, (image, prev_action, init_rnn_state, need), *self._fn_params)
        if lead == 1:
            pi, v = pi.reshape(B, -1), v.reshape(B)
        elif lead == 0:
            pi, v = pi.reshape(-1), v.reshape(())
        return pi, v, self.rnn_state_cls(h=hn, c=cn)
