"""ICM and Disagreement on B200 -- drop-ins for rlpyt/models/curiosity/icm.py:45-145 and
disagreement.py:44-167: same constructors, ``state_dict`` keys, ``feature_size`` attribute and
``compute_bonus`` / ``compute_loss`` signatures.  The ``torch.nn`` sub-modules are parameter
holders; forward and backward run in libcuriosity_b200.so.

Reference behaviours kept on purpose: pixels are fed raw 0..255 (the obs_stats
normalisation is dead code, icm.py:107-112); the forward model sees detached features; Disagreement
has four forward models whatever ``ensemble_size`` says, unbiased variance, and an always-on
dropout(p=0.2) on the squared error of its loss (masks can be injected for parity tests).
"""
import ctypes as C

import torch
from torch import nn

from .. import _lib as L
from .. import dist
from ..layers import linear_layer, narrow_head, NarrowHead, pad_side, ACT
from .encoders import make_encoder, frame_input



class ResBlock(nn.Module):
    """icm.py:9-21 (holder)."""

    def __init__(self, feature_size, action_size):
        super().__init__()
        self.lin_1 = nn.Linear(feature_size + action_size, feature_size)
        self.lin_2 = nn.Linear(feature_size + action_size, feature_size)


class ResForward(nn.Module):
    """icm.py:23-43 (holder) + executor over the library's linear layers with the one-hot action as
    a side input of every layer (icm.py:19-21: torch.cat([x, action], 2))."""

    def __init__(self, feature_size, action_size):
        super().__init__()
        self.feature_size, self.action_size = feature_size, action_size
        self.lin_1 = nn.Linear(feature_size + action_size, feature_size)
        self.res_block_1 = ResBlock(feature_size, action_size)
        self.res_block_2 = ResBlock(feature_size, action_size)
        self.res_block_3 = ResBlock(feature_size, action_size)
        self.res_block_4 = ResBlock(feature_size, action_size)
        self.lin_last = nn.Linear(feature_size + action_size, feature_size)
        self._layers = None

    def _build(self):
        if self._layers is None:
            A = self.action_size
            blocks = [self.res_block_1, self.res_block_2, self.res_block_3, self.res_block_4]
            self._layers = dict(
                first=linear_layer(self.lin_1, "lrelu", A),
                a=[linear_layer(b.lin_1, "lrelu", A) for b in blocks],
                b=[linear_layer(b.lin_2, "none", A) for b in blocks],
                last=linear_layer(self.lin_last, "none", A))
        return self._layers

    def run_forward(self, phi1_bf16, action, n):
        """-> (saved activations, predicted phi2 fp32 [n,F])."""
        ly = self._build()
        sp = pad_side(action) if self.action_size <= 64 else None      # one-hot as a bf16 k-block, built once
        x, _ = ly["first"].forward(phi1_bf16, n, side=action, side_pad=sp)
        saved = dict(x=[x], r=[])
        for i in range(4):
            r, _ = ly["a"][i].forward(x, n, side=action, side_pad=sp)
            x_new, _ = ly["b"][i].forward(r, n, side=action, residual=x, side_pad=sp)     # r2 + x   (icm.py:22)
            saved["r"].append(r)
            saved["x"].append(x_new)
            x = x_new
        _, pred = ly["last"].forward(x, n, side=action, out_bf16=False, out_f32=True, side_pad=sp)
        return saved, pred

    def run_backward(self, phi1_bf16, action, saved, dpred_bf16, n, accumulate=False):
        ly = self._build()
        xs, rs = saved["x"], saved["r"]
        ly["last"].wgrad(xs[4], dpred_bf16, n, side=action, accumulate=accumulate)
        g = ly["last"].dgrad(dpred_bf16, n)                                   # x4 = r2 + x3: no activation
        for i in range(3, -1, -1):
            ly["b"][i].wgrad(rs[i], g, n, side=action, accumulate=accumulate)
            dr = ly["b"][i].dgrad(g, n, x_saved=rs[i], x_act="lrelu")
            ly["a"][i].wgrad(xs[i], dr, n, side=action, accumulate=accumulate)
            # skip connection: total gradient at x_i = (dr W_a + g); x_0 alone carries a LeakyReLU
            g = ly["a"][i].dgrad(dr, n, dx_add=g, x_saved=xs[0] if i == 0 else None,
                                 x_act="lrelu" if i == 0 else "none")
        ly["first"].wgrad(phi1_bf16, g, n, side=action, accumulate=accumulate)


class GroupedResForward:
    """E ResForward models over the same batch, every layer position as ONE grouped tcgen05 launch
    (cb_grouped_linear_*): activations packed [n, E*F], weights stacked along the output dimension.
    The reference evaluates forward_model_1..4 one after the other (disagreement.py:127-130).
    Shapes the grouped kernels do not cover (F not a multiple of 128) run model by model into the same
    packed layout."""

    POSITIONS = ["first"] + [f"a{i}" for i in range(4)] + [f"b{i}" for i in range(4)] + ["last"]

    def __init__(self, models):
        self.models = list(models)
        self.E = len(self.models)
        self.F, self.A = self.models[0].feature_size, self.models[0].action_size
        self.grouped = self.F % 128 == 0 and self.A <= 64
        self._key = None
        self._sp = None
        self._stk = None

    def _layer(self, fm, pos):
        ly = fm._build()
        return ly[pos] if pos in ("first", "last") else ly[pos[0]][int(pos[1])]

    def _stacks(self):
        """bf16 weights of all models stacked per layer position (rebuilt when a master weight changed)."""
        for fm in self.models:
            for pos in self.POSITIONS:
                self._layer(fm, pos).prepare()
        key = tuple(self._layer(fm, pos)._version for fm in self.models for pos in self.POSITIONS)
        if key != self._key:
            stk = {}
            for pos in self.POSITIONS:
                ls = [self._layer(fm, pos) for fm in self.models]
                stk[pos] = dict(w=torch.cat([l.w_ext for l in ls]).contiguous(),
                                w_t=torch.cat([l.w_t for l in ls]).contiguous(),
                                bias=torch.cat([l.bias for l in ls]).contiguous(),
                                side_w=torch.cat([l.side_w for l in ls]).contiguous(),
                                act=ls[0].act)
            self._stk, self._key = stk, key
        return self._stk

    def _fwd(self, pos, x, n, action, shared=False, residual=None, f32=False):
        st = self._stk[pos]
        if self._sp is None:
            self._sp = pad_side(action)
        dev = x.device
        EF = self.E * self.F
        yb = None if f32 else torch.empty(n, EF, dtype=torch.bfloat16, device=dev)
        yf = torch.empty(n, EF, dtype=torch.float32, device=dev) if f32 else None
        e = L.Epilogue()
        e.bias, e.residual = L.ptr(st["bias"]), L.ptr(residual)
        e.side, e.side_w, e.side_dim = L.ptr(action), L.ptr(st["side_w"]), self.A
        e.side_bf16 = L.ptr(self._sp)
        e.out_bf16, e.out_f32 = L.ptr(yb), L.ptr(yf)
        L.call("cb_grouped_linear_fwd", n, self.E, self.F, self.F, 1 if shared else 0, L.ptr(x), L.ptr(st["w"]),
               st["act"], C.byref(e), L.stream())
        return yf if f32 else yb

    def _dgrad(self, pos, dy, n, x_saved=None, x_act="none", dx_add=None):
        dx = torch.empty(n, self.E * self.F, dtype=torch.bfloat16, device=dy.device)
        L.call("cb_grouped_linear_dgrad", n, self.E, self.F, self.F, L.ptr(dy), L.ptr(self._stk[pos]["w_t"]),
               L.ptr(x_saved), ACT[x_act], L.ptr(dx_add), L.ptr(dx), L.stream())
        return dx

    def _wgrad(self, pos, x, dy, n, action, shared=False):
        dev, E, F, A = dy.device, self.E, self.F, self.A
        dw = torch.empty(E * F, F, dtype=torch.float32, device=dev)
        db = torch.empty(E * F, dtype=torch.float32, device=dev)
        dsw = torch.empty(E * F, A, dtype=torch.float32, device=dev)
        L.call("cb_grouped_linear_wgrad", n, E, F, F, 1 if shared else 0, L.ptr(x), L.ptr(dy), L.ptr(dw), L.ptr(db),
               0.0, L.stream())
        L.call("cb_side_wgrad", L.ptr(dy), L.ptr(action), L.ptr(dsw), n, E * F, A, 0.0, L.stream())
        for e, fm in enumerate(self.models):
            mod = self._layer(fm, pos).module
            if mod.weight.grad is None:
                mod.weight.grad = torch.zeros_like(mod.weight)
            if mod.bias.grad is None:
                mod.bias.grad = torch.zeros_like(mod.bias)
            mod.weight.grad[:, :F].copy_(dw[e * F:(e + 1) * F])
            mod.weight.grad[:, F:].copy_(dsw[e * F:(e + 1) * F])
            mod.bias.grad.copy_(db[e * F:(e + 1) * F])
            dist.grads_ready((mod.weight.grad, mod.bias.grad))

    def run_forward(self, phi1_bf16, action, n):
        """-> (saved activations, predictions fp32 packed [n, E*F])."""
        if not self.grouped:
            outs = [fm.run_forward(phi1_bf16, action, n) for fm in self.models]
            return [o[0] for o in outs], torch.cat([o[1] for o in outs], 1).contiguous()
        self._stacks()
        self._sp = None
        x = self._fwd("first", phi1_bf16, n, action, shared=True)
        saved = dict(x=[x], r=[])
        for i in range(4):
            r = self._fwd(f"a{i}", x, n, action)
            x = self._fwd(f"b{i}", r, n, action, residual=x)
            saved["r"].append(r)
            saved["x"].append(x)
        return saved, self._fwd("last", x, n, action, f32=True)

    def run_backward(self, phi1_bf16, action, saved, dpred_bf16, n):
        if not self.grouped:
            F = self.F
            for e, fm in enumerate(self.models):
                fm.run_backward(phi1_bf16, action, saved[e], dpred_bf16[:, e * F:(e + 1) * F].contiguous(), n)
            return
        xs, rs = saved["x"], saved["r"]
        self._wgrad("last", xs[4], dpred_bf16, n, action)
        g = self._dgrad("last", dpred_bf16, n)
        for i in range(3, -1, -1):
            self._wgrad(f"b{i}", rs[i], g, n, action)
            dr = self._dgrad(f"b{i}", g, n, x_saved=rs[i], x_act="lrelu")
            self._wgrad(f"a{i}", xs[i], dr, n, action)
            g = self._dgrad(f"a{i}", dr, n, dx_add=g, x_saved=xs[0] if i == 0 else None,
                            x_act="lrelu" if i == 0 else "none")
        self._wgrad("first", phi1_bf16, g, n, action, shared=True)


class _IcmLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, args, *params):
        need = any(p.requires_grad for p in params)
        inv, fwd, g_inv, g_fwd = model._loss_and_grads(*args, need_grad=need)
        ctx.g_inv, ctx.g_fwd = g_inv, g_fwd
        return inv, fwd

    @staticmethod
    def backward(ctx, d_inv, d_fwd):
        out = []
        for gi, gf in zip(ctx.g_inv, ctx.g_fwd):
            if gi is not None:
                out.append(gi * d_inv)
            elif gf is not None:
                out.append(gf * d_fwd)
            else:
                out.append(None)
        return (None, None) + tuple(out)


class _IcmBase(nn.Module):
    n_forward_models = 1
    default_reuse_features = False     # integration.install() switches this on for models built under rlpyt

    def _init_common(self, image_shape, action_size, feature_encoding, batch_norm, prediction_beta, obs_stats,
                     forward_loss_wt):
        self.prediction_beta = prediction_beta
        self.feature_encoding = feature_encoding
        self.obs_stats = obs_stats
        self.action_size = action_size
        if forward_loss_wt == -1.0:                       # icm.py:70-75
            self.forward_loss_wt, self.inverse_loss_wt = 1.0, 1.0
        else:
            self.forward_loss_wt, self.inverse_loss_wt = forward_loss_wt, 1 - forward_loss_wt
        self.encoder, self.feature_size = make_encoder(feature_encoding, image_shape, batch_norm)
        self.inverse_model = nn.Sequential(nn.Linear(self.feature_size * 2, self.feature_size), nn.ReLU(),
                                           nn.Linear(self.feature_size, action_size))
        self._exec = None
        # BatchNorm2d in the encoder (`-batch_norm`): every encoder pass is its own batch in the reference (statistics
        # of that pass, running statistics updated by each, icm.py:114-117), so neither merging the obs / next_obs
        # passes (de-duplication) nor serving the loss pass from the bonus pass's features (reuse) would be exact
        self._batch_stats = bool(getattr(self.encoder, "batch_norm", False))
        # opt-in exact reuse of a batch's forward pass between compute_bonus and compute_loss (see _cached_forward)
        self.reuse_features = self.default_reuse_features
        self._cache = None
        self._wver = 0
        self.reuse_hits = {"full": 0, "miss": 0}

    @property
    def reuse_features(self):
        return self._reuse_features and not self._batch_stats

    @reuse_features.setter
    def reuse_features(self, on):
        self._reuse_features = bool(on)

    # ------------------------------------------------------------------ executors
    def _forward_models(self):
        raise NotImplementedError

    def _build(self):
        dev = self.inverse_model[0].weight.device
        if dev.type != "cuda":
            raise L.CuriosityLibError("curiosity model parameters must be on a CUDA device (no CPU fallback)")
        if self._exec is None or self._exec["dev"] != dev:
            chain, perm = self.encoder.build_chain(True)
            self._exec = dict(dev=dev, chain=chain,
                              perm=None if perm is None else perm.to(dev),
                              inv_perm=None if perm is None else torch.argsort(perm).to(dev),
                              inv0=linear_layer(self.inverse_model[0], "relu"),
                              inv2=narrow_head(self.inverse_model[2]))
        return self._exec

    def _encode(self, frames, n, first_out=None):
        """``first_out``: this encoder's first-conv output for ``frames`` computed elsewhere (the policy network's
        conv1 pass over the same frames runs it as a twin launch, algos/ppo.py)."""
        ex = self._exec
        keep, x, dtype, strides = frame_input(frames)
        acts, phi_f = ex["chain"].forward(x, n, in_dtype=dtype, strides=strides, last_f32=True, last_bf16=True,
                                          first_out=first_out)
        phi_b = acts[-1].reshape(n, -1)
        phi_f = phi_f.reshape(n, -1)
        if ex["perm"] is not None:                         # NHWC -> the reference's CHW feature order
            phi_f = phi_f.index_select(1, ex["perm"])
            phi_b = phi_b.index_select(1, ex["perm"])
        return dict(keep=keep, x=x, dtype=dtype, strides=strides, acts=acts), phi_f, phi_b

    def _encode_backward(self, enc, dphi_bf16, n, accumulate, final=True):
        ex = self._exec
        if ex["inv_perm"] is not None:
            dphi_bf16 = dphi_bf16.index_select(1, ex["inv_perm"])
        last = ex["chain"].layers[-1]
        dphi_bf16 = dphi_bf16.reshape(n * last.out_h * last.out_w, last.out_c)      # back to NHWC rows
        ex["chain"].backward(enc["x"], enc["acts"], dphi_bf16.contiguous(), n, in_dtype=enc["dtype"],
                             strides=enc["strides"], accumulate=accumulate, final=final)

    # ------------------------------------------------------------------ observation de-duplication
    # rlpyt's collectors write next_observation[t] and observation[t+1] from the same array
    # (samplers/parallel/cpu/collectors.py:118,156), so a rollout holds T+1 distinct frame stacks, not 2T.  When the two
    # arguments are PROVABLY that -- views one time step apart into one device buffer of T+1 steps, which is how
    # agents/hooks.BatchStaging.get_shifted lays them out -- the encoder runs once over the T+1 steps and phi(obs) /
    # phi(next_obs) are row ranges of its output; the backward pass adds the two feature gradients and runs once.
    # Exact (same kernels, same operands per frame); roughly 40 % fewer FLOPs per ICM step.
    dedup_shifted_observations = True

    @staticmethod
    def shifted(observations, next_observations):
        """True iff next_observations[t] aliases observations[t+1] for every t (structural, no content check)."""
        a, b = observations, next_observations
        if not (a.is_cuda and b.is_cuda and a.dtype == b.dtype == torch.uint8 and a.dim() == 5 and a.shape == b.shape
                and a.is_contiguous() and b.is_contiguous() and a.shape[0] >= 1):
            return False
        step = a[0].numel()
        if b.data_ptr() != a.data_ptr() + step:
            return False
        st = a.untyped_storage()
        return st.data_ptr() == b.untyped_storage().data_ptr() and st.nbytes() >= a.storage_offset() + (a.shape[0] + 1) * step

    def _features(self, observations, next_observations, actions, first_out_obs=None):
        """Shared by bonus and loss: (T, B, n, enc1, enc2, phi1_b, phi2_f, phi2_b, action[n,A])."""
        self._build()
        if observations.dim() != 5:
            raise ValueError("expected [T,B,C,H,W] observations")
        T, B = observations.shape[:2]
        n = T * B
        shp = observations.shape[2:]
        if self.dedup_shifted_observations and not self._batch_stats and self.shifted(observations, next_observations):
            n1 = (T + 1) * B
            allf = torch.as_strided(observations, (T + 1, B) + tuple(shp), observations.stride(),
                                    observations.storage_offset())
            fo = first_out_obs[0] if first_out_obs is not None and first_out_obs[1] == n1 else None
            enc, phi_f, phi_b = self._encode(allf.reshape(n1, *shp), n1, first_out=fo)
            action = actions.to(self._exec["dev"], torch.float32).reshape(n, -1).contiguous()
            self.dedup_hits = getattr(self, "dedup_hits", 0) + 1
            return T, B, n, dict(shared=enc, n1=n1), None, phi_b[:n], phi_f[B:], phi_b[B:], action
        fo = first_out_obs[0] if first_out_obs is not None and first_out_obs[1] == n else None
        enc1, _, phi1_b = self._encode(observations.reshape(n, *shp), n, first_out=fo)
        enc2, phi2_f, phi2_b = self._encode(next_observations.reshape(n, *shp), n)
        action = actions.to(self._exec["dev"], torch.float32).reshape(n, -1).contiguous()
        return T, B, n, enc1, enc2, phi1_b, phi2_f, phi2_b, action

    # ------------------------------------------------------------------ forward pass shared by bonus and loss
    # The reference hands the same (observation, next_observation, action) to curiosity_step and curiosity_loss
    # (algos/pg/base.py:64-66, ppo.py:93-98, 226-229).  Until the first optimiser step the loss's forward pass --
    # both encoder passes and the forward models -- is the bonus's: it is reused after a device-side equality
    # check of the three inputs (exact: identical kernels on identical operands).
    def _weights_version(self):
        return (self._wver,) + tuple(p._version for p in self.parameters())

    def _same_batch(self, c, observations, next_observations, actions):
        if c["shapes"] != (tuple(observations.shape), tuple(next_observations.shape), tuple(actions.shape)):
            return False
        dev = self._exec["dev"]
        flag = torch.zeros(1, dtype=torch.int32, device=dev)
        for a, b, ref in ((observations, c["obs"], c["ref_obs"]), (next_observations, c["next"], c["ref_next"])):
            if a.dtype != torch.uint8 or not a.is_cuda:
                return False
            if (a is b and getattr(a, "_b200_staged", None) is not None
                    and (getattr(a, "_b200_staged"), int(a._version)) == c["versions"][id(b)]):
                continue              # the very BatchStaging tensor the bonus saw (never refilled), untouched since
            a, b = a.contiguous(), ref.contiguous()      # tensors we do not own: compare with the private copy
            per = a[0, 0].numel()
            if per % 16 or (a.data_ptr() | b.data_ptr()) & 15:
                return False
            L.call("cb_frames_differ", a.data_ptr(), per, b.data_ptr(), per, per, a.shape[0] * a.shape[1], L.ptr(flag), L.stream())
        same_act = torch.equal(actions.to(dev, torch.float32).reshape(c["act"].shape), c["act"])     # tiny; host sync
        return same_act and int(flag.item()) == 0

    def _cached_forward(self, observations, next_observations, actions, remember):
        """(T, B, n, enc1, enc2, phi1_b, phi2_f, phi2_b, action, saved, pred): features + forward-model outputs."""
        self._build()
        c = self._cache
        # first-conv output of ``observations`` handed over by the policy pass (consumed once, only for that very tensor)
        handed = self.__dict__.pop("_first_out_obs", None)
        first_out_obs = None
        if handed is not None and handed[1] is observations:
            first_out_obs = (handed[0], handed[2] if len(handed) > 2 else observations.shape[0] * observations.shape[1])
        if (self.reuse_features and not remember and c is not None and c["wver"] == self._weights_version()
                and self._same_batch(c, observations, next_observations, actions)):
            self.reuse_hits["full"] += 1
            return c["out"]
        feats = self._features(observations, next_observations, actions, first_out_obs)
        T, B, n, enc1, enc2, phi1_b, phi2_f, phi2_b, action = feats
        saved, pred = self._forward_exec().run_forward(phi1_b, action, n)            # pred fp32, packed [n, E*F]
        out = feats + (saved, pred)
        if self.reuse_features and remember:
            own = lambda t: t if getattr(t, "_b200_staged", None) is not None else t.clone()
            self._cache = dict(obs=observations, next=next_observations, ref_obs=own(observations),
                               ref_next=own(next_observations), act=action.clone(), wver=self._weights_version(),
                               versions={id(t): (getattr(t, "_b200_staged", None), int(t._version))
                                         for t in (observations, next_observations)},
                               shapes=(tuple(observations.shape), tuple(next_observations.shape), tuple(actions.shape)),
                               out=out)
        elif not remember:
            self.reuse_hits["miss"] += 1
        return out

    # ------------------------------------------------------------------ loss
    def _loss_and_grads(self, observations, next_observations, actions, valid, masks, need_grad, inplace=False):
        if actions.dim() == 2:
            actions = actions.unsqueeze(1)                  # icm.py:138
        T, B, n, enc1, enc2, phi1_b, phi2_f, phi2_b, action, saved, pred = self._cached_forward(
            observations, next_observations, actions, remember=False)
        ex, st, dev, F = self._exec, L.stream(), self._exec["dev"], self.feature_size
        valid = valid.to(dev, torch.float32).reshape(n).contiguous()
        # Normaliser convention.  Fused path (``inplace``: FlatAdam SUMS the gradients over ranks): every rank divides
        # by the GLOBAL sum(valid), so the summed gradient is the single-process one.  Autograd path (rlpyt's own
        # optimiser; under SyncRl torch DDP AVERAGES gradients and the reference uses the per-rank valid_mean,
        # utils/tensor.py:39-46): local sum(valid), returned losses are per-rank means like the reference's.
        denom = dist.all_reduce_sum_(valid.sum().reshape(1)) if (inplace and dist.world_size() > 1) else None
        # inverse model (icm.py:88-92, 125)
        cat = torch.cat([phi1_b, phi2_b], 1)
        h, _ = ex["inv0"].forward(cat, n)
        inv2 = self.inverse_model[2]
        # Linear(F -> A) is A <= 8 columns wide: HBM-bound on reading h, so it runs as one fused CUDA-core pass
        fused_head = bool(L.load().cb_policy_heads_supported(F, self.action_size))
        if fused_head:
            logits = torch.empty(n, self.action_size, dtype=torch.float32, device=dev)
            L.call("cb_policy_heads_fwd", L.ptr(h), L.ptr(inv2.weight), L.ptr(inv2.bias), None, None, L.ptr(logits),
                   None, None, n, F, self.action_size, st)
        else:
            if isinstance(ex["inv2"], NarrowHead):
                logits = ex["inv2"].forward(h, n)
            else:
                _, logits = ex["inv2"].forward(h, n, out_bf16=False, out_f32=True)
        ce = torch.empty(n, dtype=torch.float32, device=dev)
        coef_inv = torch.empty(n, dtype=torch.float32, device=dev)
        L.call("cb_loss_coef", L.ptr(valid), None, float(self.inverse_loss_wt), L.ptr(denom), L.ptr(coef_inv), n, st)
        dlogits = torch.empty_like(logits) if need_grad else None
        L.call("cb_cross_entropy", L.ptr(logits), L.ptr(action), L.ptr(coef_inv), L.ptr(ce), L.ptr(dlogits), n,
               self.action_size, st)
        losses = torch.empty(2, dtype=torch.float32, device=dev)
        L.call("cb_dot", L.ptr(ce), L.ptr(coef_inv), L.ptr(losses[0:1]), n, st)
        # forward model(s) on detached features (icm.py:126)
        coef_fwd = torch.empty(n, dtype=torch.float32, device=dev)
        L.call("cb_loss_coef", L.ptr(valid), None, float(self.forward_loss_wt), L.ptr(denom), L.ptr(coef_fwd), n, st)
        E = self.n_forward_models
        fwd_terms = torch.empty(E, dtype=torch.float32, device=dev)
        fmx = self._forward_exec()
        mask = None
        if masks is not None:
            mask = torch.stack([m.to(dev, torch.float32).reshape(n, F) for m in masks]).contiguous()
        err = torch.empty(E, n, dtype=torch.float32, device=dev)
        L.call("cb_ensemble_sqerr", L.ptr(pred), L.ptr(phi2_f), L.ptr(mask), self._mask_scale, 1.0 / F, L.ptr(err),
               n, F, E, st)
        for e in range(E):
            L.call("cb_dot", L.ptr(err[e]), L.ptr(coef_fwd), L.ptr(fwd_terms[e:e + 1]), n, st)
        losses[1] = fwd_terms.sum()
        params = list(self.parameters())
        if not need_grad:
            return losses[0], losses[1], [None] * len(params), [None] * len(params)
        saved_grads = [p.grad for p in params]
        if not inplace:
            for p in params:
                p.grad = None
        # ---- backward: forward models
        dpred = torch.empty(n, E * F, dtype=torch.bfloat16, device=dev)
        L.call("cb_ensemble_sqerr_grad", L.ptr(pred), L.ptr(phi2_f), L.ptr(coef_fwd), 1.0 / F, L.ptr(mask),
               self._mask_scale if mask is not None else 1.0, L.ptr(dpred), n, F, E, st)
        fmx.run_backward(phi1_b, action, saved, dpred, n)
        # ---- backward: inverse model -> both encoder passes
        if fused_head:
            for prm in (inv2.weight, inv2.bias):
                if prm.grad is None:
                    prm.grad = torch.zeros_like(prm)
            dh = torch.empty(n, F, dtype=torch.bfloat16, device=dev)
            L.call("cb_policy_heads_bwd", L.ptr(h), L.ptr(dlogits), None, L.ptr(inv2.weight), None, L.ptr(dh),
                   L.ptr(inv2.weight.grad), L.ptr(inv2.bias.grad), None, None, n, F, self.action_size, L.ACT_RELU, st)
            dist.grads_ready((inv2.weight.grad, inv2.bias.grad))
        else:
            if isinstance(ex["inv2"], NarrowHead):
                dh = ex["inv2"].backward(h, dlogits, n, x_saved=h, x_act="relu")
            else:
                dlog_b = dlogits.to(torch.bfloat16)
                ex["inv2"].wgrad(h, dlog_b, n)
                dh = ex["inv2"].dgrad(dlog_b, n, x_saved=h, x_act="relu")
        ex["inv0"].wgrad(cat, dh, n)
        dcat = ex["inv0"].dgrad(dh, n, x_saved=cat, x_act=ex["chain"].last_act)
        if enc2 is None:            # one encoder pass over the T+1 distinct steps: d phi[t] = d phi1[t] + d phi2[t-1]
            n1 = enc1["n1"]
            dall = torch.zeros(n1, F, dtype=torch.bfloat16, device=dev)
            dall[:n] = dcat[:, :F]
            dall[n1 - n:] += dcat[:, F:]
            self._encode_backward(enc1["shared"], dall, n1, accumulate=False)
        else:
            self._encode_backward(enc1, dcat[:, :F].contiguous(), n, accumulate=False, final=False)
            self._encode_backward(enc2, dcat[:, F:].contiguous(), n, accumulate=True)
        if inplace:
            return losses[0], losses[1], None, None
        g_inv, g_fwd = [], []
        fwd_ids = {id(p) for fm in self._forward_models() for p in fm.parameters()}
        for p, old in zip(params, saved_grads):
            g = p.grad
            p.grad = old
            g_inv.append(None if id(p) in fwd_ids else g)
            g_fwd.append(g if id(p) in fwd_ids else None)
        return losses[0], losses[1], g_inv, g_fwd

    _mask_scale = 1.0

    def _default_masks(self, T, B):
        return None

    def twin_first_conv(self):
        """The encoder's first layer when it is the tcgen05 uint8 first convolution (Burda head): a network reading
        the same frames can run it as the twin of its own first conv (shared pixel conversion, N = 64 MMAs) and
        hand the result back through ``_first_out_obs``.  None when the forward pass would be served from the cache."""
        self._build()
        c = self._cache
        if self.reuse_features and c is not None and c["wver"] == self._weights_version():
            return None
        layer = self._exec["chain"].layers[0]
        return layer if (layer.k, layer.stride, layer.pad, layer.out_c) == (8, 4, 0, 32) else None

    def invalidate_weights(self):
        """Call after the fp32 master weights were updated outside torch's version tracking (FlatAdam)."""
        self._wver += 1
        if self._exec is None:
            return
        layers = list(self._exec["chain"].layers) + [self._exec["inv0"], self._exec["inv2"]]
        for fm in self._forward_models():
            ly = fm._build()
            layers += [ly["first"], ly["last"]] + ly["a"] + ly["b"]
        for layer in layers:
            layer._version = None
        if getattr(self, "_grouped", None) is not None:
            self._grouped._key = None          # the stacked ensemble weights must be rebuilt as well

    def compute_loss(self, observations, next_observations, actions, valid, dropout_masks="default"):
        """icm.py:136-145 / disagreement.py:142-167 -> (inverse_loss_wt*inv, forward_loss_wt*fwd)."""
        masks = self._default_masks(*observations.shape[:2]) if isinstance(dropout_masks, str) else dropout_masks
        return _IcmLoss.apply(self, (observations, next_observations, actions, valid, masks), *self.parameters())

    def loss_and_grads_(self, observations, next_observations, actions, valid, dropout_masks="default"):
        """Fused training entry: losses + gradients left in ``.grad`` (no autograd graph)."""
        masks = self._default_masks(*observations.shape[:2]) if isinstance(dropout_masks, str) else dropout_masks
        inv, fwd, _, _ = self._loss_and_grads(observations, next_observations, actions, valid, masks, True, inplace=True)
        return inv, fwd


class ICM(_IcmBase):
    """icm.py:45-145."""

    def __init__(self, image_shape, action_size, feature_encoding="idf", batch_norm=False, prediction_beta=1.0,
                 obs_stats=None, forward_loss_wt=0.2):
        super().__init__()
        self._init_common(image_shape, action_size, feature_encoding, batch_norm, prediction_beta, obs_stats,
                          forward_loss_wt)
        self.forward_model = ResForward(feature_size=self.feature_size, action_size=action_size)

    def _forward_models(self):
        return [self.forward_model]

    def _forward_exec(self):
        return self.forward_model

    @torch.no_grad()
    def compute_bonus(self, observations, next_observations, actions):
        """icm.py:131-134: beta * sum_f (pred_phi2 - phi2)^2 / F -> fp32 [T,B]."""
        T, B, n, _, _, phi1_b, phi2_f, _, action, _, pred = self._cached_forward(observations, next_observations,
                                                                                 actions, remember=True)
        out = torch.empty(T, B, dtype=torch.float32, device=self._exec["dev"])
        L.call("cb_rowwise_sqerr", L.ptr(pred), L.ptr(phi2_f), float(self.prediction_beta) / self.feature_size,
               L.ptr(out), n, self.feature_size, L.stream())
        return out


class Disagreement(_IcmBase):
    """disagreement.py:44-167."""

    n_forward_models = 4
    _mask_scale = 1.0 / 0.8          # F.dropout(p=0.2) rescales the kept elements

    def __init__(self, image_shape, action_size, ensemble_size=5, feature_encoding="idf", batch_norm=False,
                 prediction_beta=1.0, obs_stats=None, device="cpu", forward_loss_wt=0.2):
        super().__init__()
        self.ensemble_size = ensemble_size      # accepted and ignored, like the reference (disagreement.py:97-100)
        self._init_common(image_shape, action_size, feature_encoding, batch_norm, prediction_beta, obs_stats,
                          forward_loss_wt)
        for e in range(1, 5):
            setattr(self, f"forward_model_{e}", ResForward(feature_size=self.feature_size, action_size=action_size))

    def _forward_models(self):
        return [getattr(self, f"forward_model_{e}") for e in range(1, 5)]

    def _forward_exec(self):
        if getattr(self, "_grouped", None) is None:
            self._grouped = GroupedResForward(self._forward_models())
        return self._grouped

    def _default_masks(self, T, B):
        """disagreement.py:155-164: dropout is always in training mode."""
        dev = self.inverse_model[0].weight.device
        return [(torch.rand(T * B, self.feature_size, device=dev) > 0.2).float() for _ in range(4)]

    @torch.no_grad()
    def compute_bonus(self, observations, next_observations, actions):
        """disagreement.py:136-140: beta * mean_f var_e(pred_e) with the unbiased variance."""
        T, B, n, _, _, phi1_b, _, _, action, _, pred = self._cached_forward(observations, next_observations, actions,
                                                                             remember=True)         # pred packed [n, 4*F]
        out = torch.empty(T, B, dtype=torch.float32, device=self._exec["dev"])
        L.call("cb_ensemble_variance_packed", L.ptr(pred), self.n_forward_models, float(self.prediction_beta),
               L.ptr(out), n, self.feature_size, L.stream())
        return out
