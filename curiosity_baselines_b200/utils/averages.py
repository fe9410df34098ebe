"""Device-resident running statistics: rlpyt/utils/averages.py:41-89.

State lives on the GPU as float64 so a bonus computation never bounces to the host
(the reference does .cpu().numpy() three times per call, rnd.py:151-152,184-185).  The
``mean`` / ``var`` / ``count`` attributes read back as numpy for API compatibility.
"""
import numpy as np
import torch

from .. import _lib as L


class RunningMeanStd:
    """averages.py:41-68: float64, mean 0, var 1, count epsilon."""

    def __init__(self, epsilon=1e-4, shape=(), device=None):
        self.shape = tuple(shape)
        self.n = int(np.prod(self.shape)) if self.shape else 1
        self._epsilon = epsilon
        self._device = device
        self._mean = self._var = self._count = None
        if device is not None:
            self._alloc(device)

    def _alloc(self, device):
        self._device = torch.device(device)
        self._mean = torch.zeros(self.n, dtype=torch.float64, device=device)
        self._var = torch.ones(self.n, dtype=torch.float64, device=device)
        self._count = torch.full((1,), self._epsilon, dtype=torch.float64, device=device)
        self._ws = torch.empty(2048, dtype=torch.float64, device=device)

    def to(self, device):
        device = torch.device(device)
        if self._mean is None:
            self._alloc(device)
        elif self._mean.device != device:
            self._mean, self._var, self._count, self._ws = (
                t.to(device) for t in (self._mean, self._var, self._count, self._ws))
            self._device = device
        return self

    # numpy views for readers written against the reference
    @property
    def mean(self):
        m = self._mean.cpu().numpy().reshape(self.shape)
        return m if self.shape else m[()]

    @property
    def var(self):
        v = self._var.cpu().numpy().reshape(self.shape)
        return v if self.shape else v[()]

    @property
    def count(self):
        return float(self._count.item())

    def update(self, x):
        """averages.py:48-52: merge the batch ``x`` [n, *shape] (any float dtype, on the device) -- per-element mean
        and population variance over axis 0 in float64, two passes like numpy, then the Chan merge."""
        if self._mean is None:
            self._alloc(x.device)
        x = x.to(self._mean.device, torch.float64).reshape(x.shape[0], -1).contiguous()
        if x.shape[1] != self.n:
            raise ValueError(f"RunningMeanStd.update: expected trailing shape {self.shape}")
        L.call("cb_rms_update_columns", L.ptr(x), x.shape[0], self.n, L.ptr(self._mean), L.ptr(self._var),
               L.ptr(self._count), L.stream())

    def update_from_values(self, x, count_num=None, count_den=1.0):
        """Merge the moments of all elements of the fp32 device tensor ``x`` into a scalar
        state (averages.py:54-68).  The batch count is ``count_num / count_den`` when
        ``count_num`` (f64 device [1]) is given -- rnd.py:191-193 passes the mean valid length --
        else the number of elements (algos/pg/base.py:81).  Under data parallelism the moment
        sums (and count_num) are all-reduced first, so every rank keeps the global state."""
        from .. import dist
        assert self.n == 1
        x = x.contiguous()
        sums = torch.empty(3, dtype=torch.float64, device=x.device)
        L.call("cb_moments_partial", L.ptr(x), x.numel(), L.ptr(sums), L.ptr(self._ws), L.stream())
        dist.all_reduce_sum_(sums)
        L.call("cb_rms_merge_moments", L.ptr(sums), L.ptr(count_num), float(count_den),
               L.ptr(self._mean), L.ptr(self._var), L.ptr(self._count), L.stream())

    def state_dict(self):
        return {"mean": self._mean.clone(), "var": self._var.clone(), "count": self._count.clone()}

    def load_state_dict(self, sd):
        self._mean.copy_(sd["mean"]); self._var.copy_(sd["var"]); self._count.copy_(sd["count"])


class RewardForwardFilter:
    """averages.py:70-89, state on the device: per-env discounted running sum carried across
    batches; the mask passed by RND is the *not-done* mask and the first call seeds all envs."""

    def __init__(self, gamma, device=None):
        self.gamma = gamma
        self._rewems = None
        self._init = None

    def _ensure(self, B, device):
        if self._rewems is None:
            self._rewems = torch.zeros(B, dtype=torch.float32, device=device)
            self._init = torch.zeros(1, dtype=torch.int32, device=device)
        elif self._rewems.numel() != B:
            raise ValueError("RewardForwardFilter: number of envs changed between calls")

    @property
    def rewems(self):
        if self._rewems is None or int(self._init.item()) == 0:
            return None
        return self._rewems.cpu().numpy()

    def update(self, rews, mask=None):
        """averages.py:75-89 for one step: ``rews`` fp32 device [B]; ``mask`` is the NOT-done mask RND passes
        (rnd.py:185-188), None = no masking.  Returns the copy of the running sums [B]."""
        rews = rews.reshape(1, -1).to(torch.float32).contiguous()
        done = torch.zeros_like(rews) if mask is None else (1.0 - mask.reshape(1, -1).to(rews.device, torch.float32))
        return self.update_sequence(rews, done.contiguous())[0]

    def update_sequence(self, rews, done):
        """rews, done: fp32 device [T,B].  Applies rnd.py:186-189 for t = 0..T-1 and returns
        the stacked per-step copies [T,B]."""
        T, B = rews.shape
        self._ensure(B, rews.device)
        total = torch.empty_like(rews)
        L.call("cb_reward_filter", L.ptr(rews), L.ptr(done), float(self.gamma), L.ptr(self._rewems),
               L.ptr(self._init), L.ptr(total), T, B, L.stream())
        return total
