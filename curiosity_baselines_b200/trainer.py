"""The unit of work BASELINE.json's metric is quoted on: for one sampled [T,B] rollout batch,
  compute_bonus (forward, no grad) -> reward += r_int -> GAE / returns / valid ->
  one epoch of compute_loss forward+backward over all T*B samples -> gradient all-reduce ->
  global-norm clip -> Adam step                      (ppo.py:72-190, base.py:53-108)
driven entirely through libcuriosity_b200.so.  ``FlatAdam`` keeps the trainable parameters,
their gradients and the Adam moments in flat fp32 buffers so that clip + Adam is two kernel
launches and the data-parallel gradient exchange is a single all-reduce.
"""
import torch

from . import _lib as L
from . import dist
from .algos.pg_base import returns_on_device


class FlatAdam:
    """torch.optim.Adam (defaults) + torch.nn.utils.clip_grad_norm_ over one flat buffer
    (ppo.py:147-150).  Parameters are re-pointed to views of ``self.flat``; ``.grad`` to views of
    ``self.grad`` so the wgrad kernels write in place."""

    def __init__(self, params, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, max_norm=1.0):
        self.params = [p for p in params if p.requires_grad]
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        self.sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
        off = 0
        for p in self.params:
            k = p.numel()
            self.flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + k].view(p.shape)
            p.grad = self.grad[off:off + k].view(p.shape)
            off += k
        self.lr, self.betas, self.eps, self.max_norm = lr, betas, eps, max_norm
        self.step_count = 0

    def step(self):
        st = L.stream()
        n = self.flat.numel()
        if dist.world_size() > 1:
            dist.all_reduce_sum_(self.grad)       # coef already carries 1/global sum(valid)
        self.sumsq.zero_()
        L.call("cb_sumsq", L.ptr(self.grad), n, L.ptr(self.sumsq), st)
        self.step_count += 1
        L.call("cb_adam_clip_step", L.ptr(self.flat), L.ptr(self.m), L.ptr(self.v), L.ptr(self.grad), n,
               L.ptr(self.sumsq), float(self.max_norm), float(self.lr), float(self.betas[0]),
               float(self.betas[1]), float(self.eps), self.step_count, st)

    def grad_norm(self):
        return float(self.sumsq.sqrt().item())


class RndStep:
    """bonus + returns + update for the RND model (BASELINE configs[1])."""

    def __init__(self, model, lr=1e-4, discount=0.99, gae_lambda=0.95, clip_grad_norm=1.0):
        self.model = model
        self.opt = FlatAdam(model.forward_model.parameters(), lr=lr, max_norm=clip_grad_norm)
        self.discount, self.gae_lambda = discount, gae_lambda

    def __call__(self, next_obs, done, reward, value, bootstrap, mask=None):
        """All tensors on the device: next_obs uint8 [T,B,C,H,W]; done/reward/value fp32 [T,B];
        bootstrap [B].  Returns (loss, r_int, advantage, return_) as device tensors."""
        m = self.model
        r_int = m.compute_bonus(next_obs, done)
        reward.add_(r_int)          # in place like base.py:66 (one elementwise add on [T,B])
        ret, adv, valid, _ = returns_on_device(reward, done, value, bootstrap, self.discount,
                                               self.gae_lambda)
        loss = m.loss_and_grads_(next_obs, valid, mask)
        self.opt.step()
        m.invalidate_weights()
        return loss, r_int, adv, ret


class IcmStep:
    """bonus + returns + update for ICM / Disagreement (BASELINE configs[0], [2]): the same unit of work as
    ``RndStep`` with the (observation, next_observation, one-hot action) inputs of icm.py:131-145."""

    def __init__(self, model, lr=1e-4, discount=0.99, gae_lambda=0.95, clip_grad_norm=1.0):
        self.model = model
        self.opt = FlatAdam(model.parameters(), lr=lr, max_norm=clip_grad_norm)
        self.discount, self.gae_lambda = discount, gae_lambda

    def __call__(self, obs, next_obs, action, done, reward, value, bootstrap, dropout_masks="default"):
        m = self.model
        r_int = m.compute_bonus(obs, next_obs, action)
        reward.add_(r_int)
        ret, adv, valid, _ = returns_on_device(reward, done, value, bootstrap, self.discount, self.gae_lambda)
        inv, fwd = m.loss_and_grads_(obs, next_obs, action, valid, dropout_masks)
        self.opt.step()
        m.invalidate_weights()
        return inv + fwd, r_int, adv, ret
