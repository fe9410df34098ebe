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

    LOG_SLOTS = 8           # fp32 scalars riding at the tail of the gradient buffer (summed across ranks for free)

    def __init__(self, params, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, max_norm=1.0, bucket_bytes=2 << 20):
        self.params = [p for p in params if p.requires_grad]
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.n = n
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        self._grad_all = torch.zeros(n + self.LOG_SLOTS, dtype=torch.float32, device=dev)
        self.grad = self._grad_all[:n]
        self.logged = self._grad_all[n:]          # per-rank partial sums of logged scalars (loss, ...)
        self.reducer = dist.BucketReducer(self._grad_all, bucket_bytes)
        dist.set_grad_reducer(self.reducer)
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
        # learning rate and step counter live on the device (cb_adam_clip_step_dev): nothing that changes from step to
        # step is a launch parameter, so the whole training step can be captured in a CUDA graph (GraphedStep)
        self._lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=dev)
        self._step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self._lr = float(lr)
        self.betas, self.eps, self.max_norm = betas, eps, max_norm

    @property
    def lr(self):
        return self._lr

    @lr.setter
    def lr(self, value):
        self._lr = float(value)
        self._lr_dev.fill_(self._lr)

    @property
    def step_count(self):
        return int(self._step_dev.item())

    @step_count.setter
    def step_count(self, value):
        self._step_dev.fill_(int(value))

    def step(self):
        st = L.stream()
        n = self.n
        # gradients were all-reduced bucket by bucket while the backward pass ran (coef already carries
        # 1 / global sum(valid), so the reduction is a plain SUM); this reduces the rest and waits
        self.reducer.finish()
        self.sumsq.zero_()
        L.call("cb_sumsq", L.ptr(self.grad), n, L.ptr(self.sumsq), st)
        L.call("cb_adam_clip_step_dev", L.ptr(self.flat), L.ptr(self.m), L.ptr(self.v), L.ptr(self.grad), n,
               L.ptr(self.sumsq), float(self.max_norm), float(self.betas[0]), float(self.betas[1]), float(self.eps),
               L.ptr(self._lr_dev), L.ptr(self._step_dev), st)

    def grad_norm(self):
        return float(self.sumsq.sqrt().item())

    def log_scalar(self, slot, value):
        """Stage a per-rank partial sum (a loss normalised by the GLOBAL sum(valid)) in the tail of the gradient
        buffer: the final bucket of ``step`` sums it across ranks together with the gradients.  Returns the view
        that holds the global value once ``step`` has run (single process: the value itself)."""
        self.logged[slot].copy_(value.detach().reshape(()))
        return self.logged[slot]

    # ------------------------------------------------------------------ checkpointing (runners/minibatch_rl.py:148)
    def state_dict(self):
        """``torch.optim.Adam.state_dict()`` layout -- {'state': {i: {step, exp_avg, exp_avg_sq}}, 'param_groups'} with
        the parameters in construction order -- so that a snapshot's ``optimizer_state_dict`` interchanges with the
        reference's (agent.parameters() order, algos/pg/base.py:44)."""
        state, off, steps_done = {}, 0, self.step_count
        for i, p in enumerate(self.params):
            k = p.numel()
            if steps_done > 0:
                state[i] = {"step": torch.tensor(float(steps_done)),
                            "exp_avg": self.m[off:off + k].view(p.shape).clone(),
                            "exp_avg_sq": self.v[off:off + k].view(p.shape).clone()}
            off += k
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": 0, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        groups = sd["param_groups"]
        ids = [i for g in groups for i in g["params"]]
        if len(ids) != len(self.params):
            raise ValueError(f"optimizer state has {len(ids)} parameters, this optimiser {len(self.params)}")
        g0 = groups[0]
        self.lr, self.betas, self.eps = g0["lr"], tuple(g0["betas"]), g0["eps"]
        steps, off = set(), 0
        self.m.zero_(); self.v.zero_()
        for pos, p in zip(ids, self.params):
            k = p.numel()
            st = sd["state"].get(pos)
            if st is not None:
                if tuple(st["exp_avg"].shape) != tuple(p.shape):
                    raise ValueError(f"optimizer state {pos}: shape {tuple(st['exp_avg'].shape)} != {tuple(p.shape)}")
                self.m[off:off + k].copy_(st["exp_avg"].reshape(-1))
                self.v[off:off + k].copy_(st["exp_avg_sq"].reshape(-1))
                steps.add(int(float(st["step"])))
            off += k
        if len(steps) > 1:
            raise ValueError("FlatAdam keeps one step counter; the loaded state has several")
        self.step_count = steps.pop() if steps else 0


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
        loss = self.opt.log_scalar(0, loss)
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
        loss = self.opt.log_scalar(0, inv + fwd)
        self.opt.step()
        m.invalidate_weights()
        return loss, r_int, adv, ret


class GraphedStep:
    """CUDA-graph replay of a fused step driver (``RndStep`` / ``IcmStep``): the ~85 (RND) to ~300 (Disagreement) kernel
    launches of one bonus + returns + update step are captured once and replayed with a single launch, which removes
    the per-launch host cost that dominates when the per-GPU batch is small (strong scaling: B/N env columns per GPU).

    Everything that changes from step to step is device-resident (Adam's step counter and learning rate,
    ``cb_adam_clip_step_dev``; the running statistics; the reward filter), so a replay is a full training step.  The
    inputs are STATIC tensors: pass the same tensor objects on every call (e.g. the staging buffers the uploads land
    in) or let ``__call__`` copy into them.  Data-parallel runs capture the NCCL all-reduces with the rest."""

    def __init__(self, step, example_inputs, warmup=3):
        self.step = step
        self.static_in = list(example_inputs)
        dev = next(t for t in self.static_in if torch.is_tensor(t)).device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                   # first-use allocations and attribute calls happen here
            for _ in range(warmup):
                step(*self.static_in)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = step(*self.static_in)

    def __call__(self, *inputs):
        for dst, src in zip(self.static_in, inputs):
            if torch.is_tensor(dst) and dst is not src:
                dst.copy_(src)
        self.graph.replay()
        return self.static_out
