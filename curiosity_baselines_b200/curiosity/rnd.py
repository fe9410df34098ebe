"""Random Network Distillation on B200 -- drop-in for rlpyt/models/curiosity/rnd.py:15-212.

Same constructor, ``state_dict`` keys (forward_model.{0,2,4,7,9,11}.*, target_model.{0,2,4,7}.*),
initialisation (orthogonal, gain sqrt 2, zero bias) and method signatures as the reference.
The ``torch.nn`` modules below are parameter holders only; every forward / backward runs in
libcuriosity_b200.so.  Differences from the reference, by design:
  * running statistics stay on the GPU in float64 (no .cpu().numpy() round trips);
  * ``compute_loss`` takes an optional ``mask`` replaying rnd.py:208-209 for parity tests.
"""
import numpy as np
import torch
from torch import nn

from .. import _lib as L
from .. import dist
from ..layers import Chain, GemmLayer, conv_layer, linear_layer
from ..utils.averages import RunningMeanStd, RewardForwardFilter


def _require_cuda(t, what):
    if not t.is_cuda:
        raise L.CuriosityLibError(f"{what} must be on a CUDA device (no CPU fallback)")


class _RndLoss(torch.autograd.Function):
    """loss = valid_mean(err * mask); gradients of the predictor are produced by the library
    during forward (upstream 1) and scaled by the incoming gradient in backward."""

    @staticmethod
    def forward(ctx, model, observations, valid, mask, *params):
        loss, grads = model._loss_and_grads(observations, valid, mask,
                                            need_grad=any(p.requires_grad for p in params))
        ctx.grads = grads
        return loss

    @staticmethod
    def backward(ctx, g):
        grads = ctx.grads
        out = tuple(None if gr is None else gr * g for gr in grads)
        return (None, None, None, None) + out


class RND(nn.Module):
    """rnd.py:15-124."""

    default_reuse_features = False     # integration.install() switches this on for models built under rlpyt

    def __init__(self, image_shape, prediction_beta=1.0, drop_probability=1.0, gamma=0.99,
                 device="cpu"):
        super().__init__()
        self.prediction_beta = prediction_beta
        self.drop_probability = drop_probability
        self.gamma = gamma
        c, h, w = 1, image_shape[1], image_shape[2]
        self.image_hw = (h, w)
        self.feature_size = 512
        self.conv_feature_size = 7 * 7 * 64

        def trunk():
            return [nn.Conv2d(1, 32, kernel_size=8, stride=4), nn.LeakyReLU(),
                    nn.Conv2d(32, 64, kernel_size=4, stride=2), nn.LeakyReLU(),
                    nn.Conv2d(64, 64, kernel_size=3, stride=1), nn.LeakyReLU(),
                    nn.Flatten(), nn.Linear(self.conv_feature_size, self.feature_size)]

        self.forward_model = nn.Sequential(*trunk(), nn.ReLU(),
                                           nn.Linear(self.feature_size, self.feature_size), nn.ReLU(),
                                           nn.Linear(self.feature_size, self.feature_size))
        self.target_model = nn.Sequential(*trunk())
        for net in (self.forward_model, self.target_model):
            for m in net:
                if isinstance(m, (nn.Conv2d, nn.Linear)):
                    nn.init.orthogonal_(m.weight, np.sqrt(2))
                    m.bias.data.zero_()
        for p in self.target_model.parameters():
            p.requires_grad = False

        # running state (rnd.py:34-36), allocated on first use on the parameters' device
        self.obs_rms = RunningMeanStd(shape=(1, c, h, w))
        self.rew_rms = RunningMeanStd()
        self.rew_rff = RewardForwardFilter(gamma)
        self._chains = None
        self._norm = None
        # Opt-in exact reuse of a batch's features between compute_bonus and compute_loss (see _reusable).
        self.reuse_features = self.default_reuse_features
        self._cache = None
        self._wver = 0            # bumped whenever the predictor's master weights change outside torch's tracking
        self.reuse_hits = {"full": 0, "target": 0, "miss": 0}
        # first two convolutions of both networks without the HBM round trip between them: frames normalised once per pass
        # (cb_rnd_normalize_s2d), then one launch per pass for target + predictor (cb_rnd_conv12_twin_fwd; the predictor
        # alone, on the feature-reuse path, through cb_rnd_conv12_fwd).  RND step 23.4 -> 22.0-22.7 ms at T=128, B=1024.
        self.fuse_conv12 = True
        self._fuse_ok = None

    # ------------------------------------------------------------------ plumbing
    def _build(self):
        dev = self.forward_model[0].weight.device
        if dev.type != "cuda":
            raise L.CuriosityLibError("RND parameters must be on a CUDA device (no CPU fallback)")
        if self._chains is not None and self._dev == dev:
            return
        h, w = self.image_hw
        fm, tm = self.forward_model, self.target_model

        def trunk(net, trainable, fc_act):
            l1 = conv_layer(net[0], (h, w, 1), "lrelu", trainable)
            l2 = conv_layer(net[2], (l1.out_h, l1.out_w, 32), "lrelu", trainable)
            l3 = conv_layer(net[4], (l2.out_h, l2.out_w, 64), "lrelu", trainable)
            assert l3.out_h == l3.out_w and l3.out_h * l3.out_w * 64 == self.conv_feature_size
            fc = linear_layer(net[7], fc_act, spatial=(l3.out_h, 64), trainable=trainable)
            return [l1, l2, l3, fc]

        pred = trunk(fm, True, "relu") + [linear_layer(fm[9], "relu"), linear_layer(fm[11], "none")]
        self._chains = (Chain(pred), Chain(trunk(tm, False, "none")))
        self._dev = dev
        self.obs_rms.to(dev); self.rew_rms.to(dev)
        px = h * w
        # per-pixel integer sums and the frame count share one buffer: ONE all-reduce per bonus under data parallelism
        self._stat = torch.zeros(2 * px + 1, dtype=torch.int64, device=dev)
        self._sum = self._stat[:2 * px].view(2, px)
        self._n_total = self._stat[2 * px:]
        self._mean_len = torch.zeros(1, dtype=torch.float64, device=dev)
        self._norm = (torch.zeros(px, dtype=torch.float32, device=dev),
                      torch.ones(px, dtype=torch.float32, device=dev))
        self._norm_ver = 0
        self._refresh_norm()

    def _refresh_norm(self):
        """fp32 mean and 1/sqrt(var) of the current obs_rms (rnd.py:162-169)."""
        self._norm_ver = getattr(self, "_norm_ver", 0) + 1
        self._norm[0].copy_(self.obs_rms._mean)
        # same arithmetic as cb_rms_merge_from_sums (1 / sqrtf((float)var)) so a resumed run normalises bit-identically
        self._norm[1].copy_(torch.sqrt(self.obs_rms._var.float()).reciprocal())

    # ------------------------------------------------------------------ checkpointing of the running statistics
    # The reference keeps obs_rms / rew_rms / rew_rff as NumPy attributes outside ``state_dict`` (rnd.py:34-36), so a
    # resumed run restarts them from scratch.  These two methods carry them explicitly; the
    # ``state_dict`` keys stay the reference's.
    def stats_state_dict(self):
        """CPU copies of the running statistics (float64 moments, fp32 per-env filter state)."""
        self._build()
        rff = self.rew_rff
        return {"obs_rms": {k: v.cpu() for k, v in self.obs_rms.state_dict().items()},
                "rew_rms": {k: v.cpu() for k, v in self.rew_rms.state_dict().items()},
                "rew_rff": None if rff._rewems is None else {"rewems": rff._rewems.cpu(), "init": rff._init.cpu()}}

    def load_stats_state_dict(self, sd):
        self._build()
        self.obs_rms.load_state_dict(sd["obs_rms"])
        self.rew_rms.load_state_dict(sd["rew_rms"])
        if sd.get("rew_rff") is not None:
            r = sd["rew_rff"]
            self.rew_rff._rewems = r["rewems"].to(self._dev, torch.float32).clone()
            self.rew_rff._init = r["init"].to(self._dev, torch.int32).clone()
        self._refresh_norm()
        self._cache = None

    def _frame_view(self, obs):
        """Address the newest frame of each stack in place (rnd.py:130-131): base pointer and
        element strides (n, h, w, c) of the uint8 [T,B,C,H,W] tensor."""
        _require_cuda(obs, "observations")
        if obs.dtype != torch.uint8:
            raise L.CuriosityLibError("RND expects uint8 frame stacks")
        if obs.dim() != 5:
            raise ValueError("RND expects [T,B,C,H,W] observations")
        obs = obs.contiguous()
        T, B, Cs, H, W = obs.shape
        base = obs.data_ptr() + (Cs - 1) * H * W
        return obs, T, B, base, (Cs * H * W, W, 1, H * W)

    S2D_BYTES = 21 * 21 * 32      # one normalised 84x84 frame in 4x4-pixel-block order (include/curiosity_b200.h)

    def _can_fuse12(self, base, strides):
        """The conv1+conv2 kernel (cb_rnd_conv12_fwd) covers the reference's 84x84 geometry."""
        if not self.fuse_conv12 or self.image_hw != (84, 84) or base % 4 or strides[0] % 4:
            return False
        l1, l2 = self._chains[0].layers[:2]
        if type(l1) is not GemmLayer or type(l2) is not GemmLayer:
            return False
        if self._fuse_ok is None:
            self._fuse_ok = bool(L.load().cb_rnd_conv12_supported())
        return self._fuse_ok

    def _normalized(self, base, n, stride_n):
        """Normalised, clamped bf16 copy of the newest frames (rnd.py:169-170) in the block order the first
        convolution reads: computed once, consumed by target and predictor."""
        img = torch.empty(n * self.S2D_BYTES, dtype=torch.uint8, device=self._dev)
        L.call("cb_rnd_normalize_s2d", base, stride_n, n, L.ptr(self._norm[0]), L.ptr(self._norm[1]), L.ptr(img),
               L.stream())
        return img

    def _conv12(self, chain, img, n, keep_a1):
        """First two convolutions of one network in one launch; the 20x20x32 activation between them is written
        out only when the backward pass will need it."""
        l1, l2 = chain.layers[0], chain.layers[1]
        l1.prepare(); l2.prepare()
        dev = l1.w_fwd.device
        a1 = torch.empty(n * 400, 32, dtype=torch.bfloat16, device=dev) if keep_a1 else None
        a2 = torch.empty(n * 81, 64, dtype=torch.bfloat16, device=dev)
        L.call("cb_rnd_conv12_fwd", L.ptr(img), n, L.ptr(l1.w_native), L.ptr(l1.bias), L.ptr(l2.w_fwd), L.ptr(l2.bias),
               L.ptr(a1), L.ptr(a2), L.stream())
        return [a1, a2]

    def _conv12_twin(self, img, n, keep_a1):
        """First two convolutions of target AND predictor in one launch (the block image is read once, the first
        layer is a single N = 64 problem); the predictor's 20x20x32 activation is written out only when the backward
        pass will need it."""
        pred, targ = self._chains
        (t1, t2), (p1, p2) = targ.layers[:2], pred.layers[:2]
        for layer in (t1, t2, p1, p2):
            layer.prepare()
        dev = t1.w_fwd.device
        a1_p = torch.empty(n * 400, 32, dtype=torch.bfloat16, device=dev) if keep_a1 else None
        a2_t = torch.empty(n * 81, 64, dtype=torch.bfloat16, device=dev)
        a2_p = torch.empty(n * 81, 64, dtype=torch.bfloat16, device=dev)
        L.call("cb_rnd_conv12_twin_fwd", L.ptr(img), n, L.ptr(t1.w_native), L.ptr(t1.bias), L.ptr(t2.w_fwd), L.ptr(t2.bias),
               L.ptr(p1.w_native), L.ptr(p1.bias), L.ptr(p2.w_fwd), L.ptr(p2.bias), L.ptr(a2_t), L.ptr(a1_p), L.ptr(a2_p),
               L.stream())
        return a2_t, a1_p, a2_p

    def _forward_nets(self, obs, keep_acts):
        obs, T, B, base, strides = self._frame_view(obs)
        n = T * B
        pred, targ = self._chains
        if self._can_fuse12(base, strides):
            img = self._normalized(base, n, strides[0])
            a2_t, a1_p, a2_p = self._conv12_twin(img, n, keep_acts)
            _, phi = targ.forward(base, n, first_out=[None, a2_t])
            acts, phi_hat = pred.forward(base, n, first_out=[a1_p, a2_p])
            return obs, T, B, base, strides, phi, phi_hat, (acts if keep_acts else None)
        # target and predictor read the same normalised frame: one fused first-conv pass for both
        a1_p, _, a1_t = pred.layers[0].forward(base, n, in_dtype=L.U8, strides=strides, norm=self._norm,
                                               twin=targ.layers[0])
        _, phi = targ.forward(base, n, first_out=a1_t)
        acts, phi_hat = pred.forward(base, n, first_out=a1_p)
        return obs, T, B, base, strides, phi, phi_hat, (acts if keep_acts else None)

    def _forward_predictor(self, obs):
        """Predictor alone (target features reused): single-network first conv."""
        obs, T, B, base, strides = self._frame_view(obs)
        if self._can_fuse12(base, strides):
            pred = self._chains[0]
            img = self._normalized(base, T * B, strides[0])
            return pred.forward(base, T * B, first_out=self._conv12(pred, img, T * B, True))
        acts, phi_hat = self._chains[0].forward(base, T * B, in_dtype=L.U8, strides=strides, norm=self._norm)
        return acts, phi_hat

    # ------------------------------------------------------------------ feature reuse
    # The reference hands the SAME frames to both calls of a training iteration: process_returns passes
    # samples.env.next_observation to compute_bonus (algos/pg/base.py:72-73) and optimize_agent passes it again
    # to compute_loss (ppo.py:107-110, 235-237); compute_loss does not touch the observation statistics.  The
    # target network is frozen, so its features of that batch are valid for every epoch / minibatch over it; the
    # predictor's forward is valid until its weights change, i.e. for the first update after the bonus.  Reuse is
    # exact (identical kernels on identical operands) and guarded by a device-side comparison of the frames.
    def _weights_version(self):
        return (self._wver,) + tuple(p._version for p in self.forward_model.parameters())

    def _remember(self, obs, T, B, phi, phi_hat, acts):
        staged = getattr(obs, "_b200_staged", None)
        # a tensor we do not own may be refilled behind torch's back (raw-pointer uploads keep address and _version):
        # the guard then compares against a PRIVATE copy of the newest frames (0.93 GB at T=128, B=1024; ~0.3 ms)
        ref = obs if staged is not None else obs[:, :, -1:].clone()
        self._cache = dict(obs=obs, ref=ref, obs_version=int(obs._version), staged=staged, shape=(T, B), phi=phi,
                           phi_hat=phi_hat, acts=acts, norm_ver=self._norm_ver, wver=self._weights_version())

    def _reusable(self, obs, T, B, base, strides):
        """None, or the cached entry whose frames equal `obs` under the current observation statistics."""
        c = self._cache
        if not self.reuse_features or c is None or c["shape"] != (T, B) or c["norm_ver"] != self._norm_ver:
            return None
        co = c["obs"]
        # Provably the same frames only when it is the very tensor object the bonus saw, it was produced by the hooks'
        # BatchStaging (filled once at creation, never refilled: tagged ``_b200_staged``) and nobody wrote to it through
        # torch since.  Anything else -- another tensor, or a persistent buffer that a raw-pointer upload may have
        # refilled without bumping ``_version`` -- is compared on the device against the reference kept by _remember.
        if (co is not obs or c["staged"] is None or getattr(obs, "_b200_staged", None) != c["staged"]
                or c["obs_version"] != int(obs._version)):
            cobs, _, _, cbase, cstrides = self._frame_view(c["ref"])
            if not ((base | cbase | strides[0] | cstrides[0]) & 15 == 0):
                return None
            flag = torch.zeros(1, dtype=torch.int32, device=self._dev)
            L.call("cb_frames_differ", base, strides[0], cbase, cstrides[0], self.image_hw[0] * self.image_hw[1], T * B,
                   L.ptr(flag), L.stream())
            if int(flag.item()) != 0:           # host sync: the price of the guard (one scalar)
                return None
        return c

    # ------------------------------------------------------------------ bonus
    @torch.no_grad()
    def compute_bonus(self, next_observation, done):
        """rnd.py:181-203 -> fp32 [T,B] on the device."""
        self._build()
        obs, T, B, base, strides = self._frame_view(next_observation)
        done = done.to(self._dev, torch.float32).reshape(T, B).contiguous()
        st = L.stream()
        px = self.image_hw[0] * self.image_hw[1]
        # observation statistics over the first n_valid[b] frames of each env (rnd.py:150-160)
        n_valid = torch.empty(B, dtype=torch.int32, device=self._dev)
        L.call("cb_valid_lengths", L.ptr(done), L.ptr(n_valid), L.ptr(self._n_total), T, B, st)
        self._sum.zero_()
        L.call("cb_obs_moments_u8", base, strides[0], px, L.ptr(n_valid), T, B,
               L.ptr(self._sum[0]), L.ptr(self._sum[1]), st)
        dist.all_reduce_sum_(self._stat)         # exact integer sums + count: sharded == single process
        rms = self.obs_rms
        L.call("cb_rms_merge_from_sums", L.ptr(self._sum[0]), L.ptr(self._sum[1]),
               L.ptr(self._n_total), L.ptr(rms._mean), L.ptr(rms._var), L.ptr(rms._count),
               L.ptr(self._norm[0]), L.ptr(self._norm[1]), px, st)
        self._norm_ver += 1                       # the statistics the forward passes normalise with have changed
        # target / predictor features and per-sample error (rnd.py:174-183)
        _, _, _, _, _, phi, phi_hat, acts = self._forward_nets(obs, self.reuse_features)
        if self.reuse_features:
            self._remember(obs, T, B, phi, phi_hat, acts)
        err = torch.empty(T, B, dtype=torch.float32, device=self._dev)
        L.call("cb_rowwise_sqerr", L.ptr(phi_hat), L.ptr(phi), 1.0 / self.feature_size, L.ptr(err),
               T * B, self.feature_size, st)
        # reward normalisation (rnd.py:184-203)
        total = self.rew_rff.update_sequence(err, done)
        # sum |done-1| over the GLOBAL batch (numerator of the mean valid length, rnd.py:190): for 0/1 dones it is the
        # frame count that was just all-reduced with the pixel sums; anything else is summed and reduced separately
        if bool(getattr(self, "binary_done", True)):
            self._mean_len.copy_(self._n_total)
        else:
            L.call("cb_valid_count", L.ptr(done), L.ptr(self._mean_len), T * B, st)
            dist.all_reduce_sum_(self._mean_len)
        self.rew_rms.update_from_values(total, self._mean_len, float(B * dist.world_size()))
        out = torch.empty_like(err)
        L.call("cb_rnd_scale_bonus", L.ptr(err), L.ptr(done), L.ptr(self.rew_rms._var),
               float(self.prediction_beta), L.ptr(out), T * B, st)
        return out

    # ------------------------------------------------------------------ loss
    def _loss_and_grads(self, observations, valid, mask, need_grad, inplace=False):
        self._build()
        obs, T, B, base, strides = self._frame_view(observations)
        hit = self._reusable(obs, T, B, base, strides)
        if hit is not None and hit["wver"] == self._weights_version() and (hit["acts"] is not None or not need_grad):
            phi, phi_hat, acts = hit["phi"], hit["phi_hat"], hit["acts"]          # same frames, statistics and weights
            self.reuse_hits["full"] += 1
        elif hit is not None:
            phi = hit["phi"]                                                     # frozen target: still valid
            acts, phi_hat = self._forward_predictor(obs)
            self.reuse_hits["target"] += 1
        else:
            obs, T, B, base, strides, phi, phi_hat, acts = self._forward_nets(observations, need_grad)
            self.reuse_hits["miss"] += 1
        n, F = T * B, self.feature_size
        st = L.stream()
        valid = valid.to(self._dev, torch.float32).reshape(n).contiguous()
        if mask is None:   # rnd.py:208-209
            mask = (torch.rand(T, B, device=self._dev) > self.drop_probability).float()
        mask = mask.to(self._dev, torch.float32).reshape(n).contiguous()
        err = torch.empty(n, dtype=torch.float32, device=self._dev)
        L.call("cb_rowwise_sqerr", L.ptr(phi_hat), L.ptr(phi), 1.0 / F, L.ptr(err), n, F, st)
        coef = torch.empty(n, dtype=torch.float32, device=self._dev)
        denom = None
        # fused path (FlatAdam sums gradients over ranks): global sum(valid).  Autograd path (rlpyt's optimiser; DDP
        # averages gradients, the reference takes the per-rank valid_mean): local sum(valid).  See icm.py.
        if inplace and dist.world_size() > 1:
            denom = dist.all_reduce_sum_(valid.sum().reshape(1))
        L.call("cb_loss_coef", L.ptr(valid), L.ptr(mask), 1.0, L.ptr(denom), L.ptr(coef), n, st)
        loss = torch.empty(1, dtype=torch.float32, device=self._dev)
        L.call("cb_dot", L.ptr(err), L.ptr(coef), L.ptr(loss), n, st)
        params = [p for p in self.forward_model.parameters()]
        if not need_grad:
            return loss[0], [None] * len(params)
        dphi = torch.empty(n, F, dtype=torch.bfloat16, device=self._dev)
        L.call("cb_sqerr_grad", L.ptr(phi_hat), L.ptr(phi), L.ptr(coef), 1.0 / F, None, 1.0,
               L.ptr(dphi), n, F, st)
        if inplace:     # overwrite the existing .grad buffers (views of a flat buffer)
            self._chains[0].backward(base, acts, dphi, n, in_dtype=L.U8, strides=strides,
                                     norm=self._norm)
            return loss[0], None
        saved = [p.grad for p in params]
        for p in params:
            p.grad = None
        self._chains[0].backward(base, acts, dphi, n, in_dtype=L.U8, strides=strides, norm=self._norm)
        grads = [p.grad for p in params]
        for p, g in zip(params, saved):
            p.grad = g
        return loss[0], grads

    def compute_loss(self, observations, valid, mask=None):
        """rnd.py:205-212 -> scalar (autograd-connected to the predictor's parameters)."""
        params = list(self.forward_model.parameters())
        return _RndLoss.apply(self, observations, valid, mask, *params)

    def invalidate_weights(self):
        """Call after the fp32 master weights were updated outside torch's version tracking."""
        self._wver += 1
        if self._chains is not None:
            for layer in self._chains[0].layers:
                layer._version = None

    def loss_and_grads_(self, observations, valid, mask=None):
        """Fused training entry used by the B200 step driver: computes the loss and leaves the
        predictor gradients directly in ``.grad`` (no autograd graph, no extra copies)."""
        loss, _ = self._loss_and_grads(observations, valid, mask, need_grad=True, inplace=True)
        return loss
