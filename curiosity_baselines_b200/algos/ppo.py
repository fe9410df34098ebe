"""PPO on B200: the reference's ``PPO.optimize_agent`` / ``PPO.loss``
(rlpyt/algos/pg/ppo.py:72-242) for the recurrent categorical agent, with the policy network, the curiosity
model, the loss and the optimiser all running in libcuriosity_b200.so.

Per iteration (``optimize_agent``):  curiosity bonus -> reward += r_int -> GAE / returns / valid
(``algos.pg_base.returns_on_device``) -> ``epochs`` x ``minibatches`` gradient steps, each over all T and a
shuffled slice of B (recurrent minibatching, ppo.py:133-140):  policy forward -> ``cb_ppo_loss`` (clipped
surrogate + value + entropy and their gradient) -> policy backward -> curiosity ``loss_and_grads_`` ->
ONE global-norm clip + Adam over policy U curiosity parameters (``FlatAdam``; ppo.py:147-150) -> linear
learning-rate and ratio-clip decay (:184-186).  Data-parallel: every normaliser ``sum(valid)`` is all-reduced
so the sharded loss equals the single-process one; gradients are summed by one all-reduce in ``FlatAdam``.
"""
from collections import namedtuple

import numpy as np
import torch

from .. import _lib as L
from .. import dist
from ..trainer import FlatAdam
from .pg_base import returns_on_device

PpoBatch = namedtuple("PpoBatch", ["observation", "next_observation", "prev_action", "action", "old_prob", "reward",
                                   "done", "value", "bootstrap_value", "init_rnn_state"])
PpoBatch.__doc__ = """One sampled rollout batch on the device (the fields of rlpyt's ``Samples`` that ppo.py reads):
observation / next_observation uint8 [T,B,C,H,W]; prev_action / action int64 [T,B]; old_prob fp32 [T,B,A]
(``samples.agent.agent_info.dist_info.prob``); reward / done / value fp32 [T,B]; bootstrap_value [B];
init_rnn_state (h, c) each [B,1,H] (``prev_rnn_state[0]``, ppo.py:124-126) or None."""

OptInfo = namedtuple("OptInfo", ["loss", "pi_loss", "value_loss", "entropy_loss", "inv_loss", "forward_loss",
                                 "entropy", "perplexity", "gradNorm"])


def _onehot(idx, A):
    return torch.zeros(idx.shape + (A,), dtype=torch.float32, device=idx.device).scatter_(-1, idx.unsqueeze(-1), 1)


class PPO:
    """ppo.py:20-70 constructor arguments (same names and defaults)."""

    twin_conv1 = True       # run the policy's and the curiosity encoder's first convolution over `observation` as one launch

    def __init__(self, discount=0.99, learning_rate=0.001, value_loss_coeff=1., entropy_loss_coeff=0.01,
                 clip_grad_norm=1., gae_lambda=1, minibatches=4, epochs=4, ratio_clip=0.1, linear_lr_schedule=True,
                 normalize_advantage=False, normalize_reward=False, curiosity_type="none",
                 initial_optim_state_dict=None):
        self.discount, self.learning_rate = discount, learning_rate
        self.value_loss_coeff, self.entropy_loss_coeff = value_loss_coeff, entropy_loss_coeff
        self.clip_grad_norm, self.gae_lambda = clip_grad_norm, gae_lambda
        self.minibatches, self.epochs = minibatches, epochs
        self.ratio_clip = self._ratio_clip = ratio_clip
        self.linear_lr_schedule = linear_lr_schedule
        self.normalize_advantage, self.normalize_reward = normalize_advantage, normalize_reward
        self.curiosity_type = curiosity_type
        self.initial_optim_state_dict = initial_optim_state_dict      # algos/pg/base.py:45-46
        self.intrinsic_rewards = None
        self.update_counter = 0
        self._returns_state = None

    def initialize(self, model, n_itr, rng=None):
        """``model``: a ``pg.atari_lstm_model.AtariLstmModel`` on the device.  One optimiser over ALL of its parameters
        (policy + curiosity, algos/pg/base.py:44)."""
        self.model, self.n_itr = model, n_itr
        self.opt = FlatAdam(model.parameters(), lr=self.learning_rate, max_norm=self.clip_grad_norm)
        self.rng = rng or np.random
        if self.initial_optim_state_dict is not None:
            self.opt.load_state_dict(self.initial_optim_state_dict)
        if self.normalize_reward:
            from .pg_base import ReturnsState
            self._returns_state = ReturnsState(self.discount)

    def optim_state_dict(self):
        """algos/base.py: what the runner snapshots as ``optimizer_state_dict`` (runners/minibatch_rl.py:148), in
        torch.optim.Adam's layout so that it interchanges with the reference."""
        return self.opt.state_dict()

    def load_optim_state_dict(self, sd):
        self.opt.load_state_dict(sd)

    # ------------------------------------------------------------------ loss + gradients of one minibatch
    def loss_and_grads_(self, observation, prev_action, action, return_, advantage, valid, old_prob, init_rnn_state,
                        curiosity_inputs=None, dropout_masks="default"):
        """ppo.py:192-242 with gradients left in ``.grad`` of every parameter (no autograd graph).  Tensors on the
        device, time-major; init_rnn_state (h, c) each [B,1,H] as ppo.py slices it (transposed here like :203-205)
        or None.  ``curiosity_inputs``: the positional inputs of the curiosity model's ``loss_and_grads_`` after
        which ``valid`` is appended.  Returns an ``OptInfo`` of device scalars (gradNorm filled by ``step``)."""
        m = self.model
        dev = m.pi.weight.device
        T, B = action.shape
        n, A = T * B, m.action_size
        st = L.stream()
        if init_rnn_state is not None:
            init_rnn_state = tuple(s.transpose(0, 1).contiguous() for s in init_rnn_state)
        # the policy's conv1 and the curiosity encoder's conv1 read the same uint8 frames: one twin launch (shared pixel
        # conversion, N = 64 MMAs) computes both; the curiosity model picks its half up in loss_and_grads_ below
        cm = getattr(m, "curiosity_model", None)
        twin, twin_steps = None, T
        if (self.twin_conv1 and self.curiosity_type in ("icm", "disagreement") and curiosity_inputs is not None
                and curiosity_inputs[0] is observation):
            twin = cm.twin_first_conv()
            if (twin is not None and cm.dedup_shifted_observations and not getattr(cm, "_batch_stats", False)
                    and cm.shifted(curiosity_inputs[0], curiosity_inputs[1])):
                twin_steps = T + 1          # next_obs = obs shifted by one step: the curiosity encoder wants all T+1 steps
        s = m._run_forward(observation, _onehot(prev_action, A), init_rnn_state, save=True, twin=twin, twin_steps=twin_steps)
        if s["twin_out"] is not None:
            cm._first_out_obs = (s["twin_out"], observation, twin_steps * B)
        validf = (torch.ones(n, device=dev) if valid is None else valid.to(dev, torch.float32).reshape(n)).contiguous()
        denom = dist.all_reduce_sum_(validf.sum().reshape(1)) if dist.world_size() > 1 else None
        coef = torch.empty(n, dtype=torch.float32, device=dev)
        L.call("cb_loss_coef", L.ptr(validf), None, 1.0, L.ptr(denom), L.ptr(coef), n, st)
        terms = torch.empty(4, n, dtype=torch.float32, device=dev)
        dlogits = torch.empty(n, A, dtype=torch.float32, device=dev)
        dvalue = torch.empty(n, dtype=torch.float32, device=dev)
        # (named so that the operands outlive the asynchronous launch)
        act_i = action.to(dev, torch.int64).reshape(n).contiguous()
        old_p = old_prob.to(dev, torch.float32).reshape(n, A).contiguous()
        adv_f = advantage.to(dev, torch.float32).reshape(n).contiguous()
        ret_f = return_.to(dev, torch.float32).reshape(n).contiguous()
        L.call("cb_ppo_loss", L.ptr(s["logits"]), L.ptr(s["value"]), L.ptr(act_i), L.ptr(old_p), L.ptr(adv_f),
               L.ptr(ret_f), L.ptr(coef), float(self.ratio_clip), float(self.value_loss_coeff),
               float(self.entropy_loss_coeff), None, L.ptr(terms), L.ptr(dlogits), L.ptr(dvalue), n, A, st)
        means = torch.empty(4, dtype=torch.float32, device=dev)
        for i in range(4):
            L.call("cb_dot", L.ptr(terms[i]), L.ptr(coef), L.ptr(means[i:i + 1]), n, st)
        m._run_backward(s, dlogits, dvalue)
        pi_loss = -means[0]
        value_loss = self.value_loss_coeff * means[1]
        entropy_loss = -self.entropy_loss_coeff * means[2]
        loss = pi_loss + value_loss + entropy_loss
        zero = torch.zeros((), device=dev)
        inv = fwd = zero
        if self.curiosity_type in ("icm", "disagreement"):
            inv, fwd = m.curiosity_model.loss_and_grads_(*curiosity_inputs, valid, dropout_masks)
        elif self.curiosity_type == "rnd":
            fwd = m.curiosity_model.loss_and_grads_(*curiosity_inputs, valid)
        elif self.curiosity_type == "ndigo":
            # ndigo.py:220-274 (``valid`` is ignored there).  NDIGO's stages are tied together by autograd Functions
            # whose backward passes run in the library; the gradients accumulate into the flat buffer's views.
            cm_params = [p for p in m.curiosity_model.parameters() if p.grad is not None]
            for p in cm_params:
                p.grad.zero_()
            fwd = m.curiosity_model.compute_loss(*curiosity_inputs, valid)
            fwd.backward()
            fwd = fwd.detach()
        loss = loss + inv + fwd
        return OptInfo(loss, pi_loss, value_loss, entropy_loss, inv, fwd, means[2], means[3], None)

    def step(self):
        """clip_grad_norm_ over all parameters + Adam.step (ppo.py:147-150); weights changed -> prepared copies stale."""
        self.opt.step()
        self.model.invalidate_weights()
        self.update_counter += 1

    # ------------------------------------------------------------------ one iteration
    def process_returns(self, batch):
        """algos/pg/base.py:53-108 on device tensors: bonus, in-place reward add, returns / advantages / valid."""
        cm = getattr(self.model, "curiosity_model", None)
        A = self.model.action_size
        r_int = None
        with torch.no_grad():
            if self.curiosity_type in ("icm", "disagreement"):
                r_int = cm.compute_bonus(batch.observation, batch.next_observation, _onehot(batch.action, A))
            elif self.curiosity_type == "rnd":
                r_int = cm.compute_bonus(batch.next_observation, batch.done)
            elif self.curiosity_type == "ndigo":
                r_int = cm.compute_bonus(batch.observation, _onehot(batch.prev_action, A), _onehot(batch.action, A))
        if r_int is not None:
            batch.reward.add_(r_int)                       # in place (base.py:66)
            self.intrinsic_rewards = r_int
        ret, adv, valid, _ = returns_on_device(batch.reward, batch.done, batch.value, batch.bootstrap_value,
                                               self.discount, self.gae_lambda, self.normalize_reward,
                                               self.normalize_advantage, True, self._returns_state)
        self.last_return, self.last_advantage = ret, adv
        return ret, adv, valid

    def minibatch_indices(self, B):
        """utils/misc.py:17-26 (iterate_mb_idxs with shuffle) over the B dimension (recurrent agents)."""
        mb = B // self.minibatches
        idx = np.arange(B)
        self.rng.shuffle(idx)
        return [idx[s:s + mb] for s in range(0, B - mb + 1, mb)]

    def optimize_agent(self, itr, batch, dropout_masks="default"):
        """ppo.py:72-190 for one ``PpoBatch``; returns a list of ``OptInfo`` (one per gradient step)."""
        ret, adv, valid = self.process_returns(batch)
        T, B = batch.reward.shape
        infos = []
        for _ in range(self.epochs):
            for idxs in self.minibatch_indices(B):
                whole = len(idxs) == B and self.minibatches == 1
                sl = (lambda t: t) if whole else (lambda t, i=torch.as_tensor(idxs, device=batch.reward.device):
                                                  t.index_select(1, i).contiguous())
                rnn = None if batch.init_rnn_state is None else tuple(
                    s if whole else s.index_select(0, torch.as_tensor(idxs, device=s.device)) for s in batch.init_rnn_state)
                obs, act = sl(batch.observation), sl(batch.action)
                cur = None
                if self.curiosity_type in ("icm", "disagreement"):
                    cur = (obs, sl(batch.next_observation), _onehot(act, self.model.action_size))
                elif self.curiosity_type == "rnd":
                    cur = (sl(batch.next_observation),)
                elif self.curiosity_type == "ndigo":
                    A = self.model.action_size
                    cur = (obs, _onehot(sl(batch.prev_action), A), _onehot(act, A))
                info = self.loss_and_grads_(obs, sl(batch.prev_action), act, sl(ret), sl(adv), sl(valid),
                                            sl(batch.old_prob), rnn, cur, dropout_masks)
                self.step()
                infos.append(info._replace(gradNorm=self.opt.sumsq.sqrt().float()))
        if self.linear_lr_schedule:                        # ppo.py:184-186 (LambdaLR stepped once per iteration)
            frac = (self.n_itr - (itr + 1)) / self.n_itr
            self.opt.lr = self.learning_rate * frac
            self.ratio_clip = self._ratio_clip * (self.n_itr - itr) / self.n_itr
        return infos


# ----------------------------------------------------------------------------------------------------------------------
# rlpyt-facing entry: ``PPO.optimize_agent(itr, samples)`` on the sampler's HOST buffers (ppo.py:72-190)
# ----------------------------------------------------------------------------------------------------------------------
RefOptInfo = namedtuple("OptInfo", ["return_", "intrinsic_rewards", "valpred", "advantage", "loss", "pi_loss",
                                    "value_loss", "entropy_loss", "inv_loss", "forward_loss", "reward_total_std",
                                    "curiosity_loss", "gradNorm", "entropy", "perplexity"])
RefOptInfo.__doc__ = "rlpyt/algos/pg/base.py:9-23 (same fields, same order): what the runner logs."


def stage_samples(agent, curiosity_type, samples, st=None):
    """rlpyt ``Samples`` on the host -> ``PpoBatch`` on the device through the agent's ``BatchStaging``: every tensor
    crosses PCIe once per iteration (the reference uploads ``observation`` twice and clones both frame tensors on the
    host first, ppo.py:81-98), frames asynchronously on the staging stream; a batch staged ahead with
    ``hooks.prefetch_samples`` is picked up without another upload."""
    from ..agents import hooks
    st = st or hooks.staging(agent)
    st.new_iteration()
    dev = st.dev
    env, ag = samples.env, samples.agent
    T, B = env.reward.shape[:2]
    f32 = lambda t: torch.as_tensor(t).to(dev, torch.float32, non_blocking=True)
    if curiosity_type in ("icm", "disagreement") and st.assume_shifted:
        obs, nxt = st.get_shifted(env.observation, env.next_observation)
    else:
        obs = st.get(env.observation)
        if curiosity_type == "rnd":            # RND reads the newest frame of each stack only (rnd.py:130-131)
            nxt = st.get(env.next_observation, newest_only=True)
        else:
            nxt = st.get(env.next_observation) if curiosity_type in ("icm", "disagreement") else obs
    rnn = getattr(ag.agent_info, "prev_rnn_state", None)
    init = None
    if rnn is not None:
        init = tuple(torch.as_tensor(s[0]).to(dev, non_blocking=True) for s in (rnn.h, rnn.c))       # T = 0: [B,N,H]
    return PpoBatch(obs, nxt, torch.as_tensor(ag.prev_action).to(dev, non_blocking=True).long(),
                    torch.as_tensor(ag.action).to(dev, non_blocking=True).long(), f32(ag.agent_info.dist_info.prob),
                    f32(env.reward).reshape(T, B).contiguous(), f32(env.done).reshape(T, B).contiguous(),
                    f32(ag.agent_info.value).reshape(T, B).contiguous(), f32(ag.bootstrap_value).reshape(-1).contiguous(),
                    init)


def optimize_agent_from_samples(native, agent, itr, samples, after_staging=None):
    """One iteration of the fused PPO on rlpyt ``Samples``: stage -> ``PPO.optimize_agent`` -> the reference's
    ``OptInfo`` (lists of python floats, one entry per gradient step) read back in ONE device->host copy.  Side
    effects of the reference are kept: the bonus is added into ``samples.env.reward`` in place and kept as numpy in
    ``intrinsic_rewards`` (algos/pg/base.py:66-67)."""
    batch = stage_samples(agent, native.curiosity_type, samples)
    if after_staging is not None:      # e.g. stage the NEXT batch now: behind this iteration's small uploads in the DMA queue
        after_staging()
    reward_before = batch.reward.clone() if native.curiosity_type != "none" else None
    infos = native.optimize_agent(itr, batch)
    fields = ("loss", "pi_loss", "value_loss", "entropy_loss", "inv_loss", "forward_loss", "gradNorm", "entropy",
              "perplexity")
    packed = torch.stack([torch.stack([torch.as_tensor(getattr(i, f), dtype=torch.float32, device=batch.reward.device)
                                       .reshape(()) for f in fields]) for i in infos])
    extra = torch.stack([native.last_return.mean(), native.last_advantage.mean(), batch.value.mean()])
    host = torch.cat([packed.reshape(-1), extra]).cpu()                 # the iteration's only blocking read-back
    vals = host[:-3].reshape(len(infos), len(fields))
    out = {f: [float(v) for v in vals[:, j]] for j, f in enumerate(fields)}
    r_mean = []
    if native.intrinsic_rewards is not None:
        r_host = native.intrinsic_rewards.cpu()
        samples.env.reward.copy_(batch.reward.reshape(samples.env.reward.shape).cpu())      # in-place add (base.py:66)
        native.intrinsic_rewards_numpy = r_host.numpy()
        r_mean = [float(r_host.mean())] * len(infos)
    ctype = native.curiosity_type
    return RefOptInfo(return_=[float(host[-3])], intrinsic_rewards=r_mean, valpred=[float(host[-1])],
                      advantage=[float(host[-2])], loss=out["loss"], pi_loss=out["pi_loss"], value_loss=out["value_loss"],
                      entropy_loss=out["entropy_loss"], inv_loss=out["inv_loss"] if ctype in ("icm", "disagreement") else [],
                      forward_loss=out["forward_loss"] if ctype != "none" else [], reward_total_std=[], curiosity_loss=[],
                      gradNorm=out["gradNorm"], entropy=out["entropy"], perplexity=out["perplexity"])


def rlpyt_optimize_agent(self, itr, samples):
    """Drop-in for rlpyt ``PPO.optimize_agent`` (bound by ``integration.install(policy=True, algo=True)``): the
    reference algo object keeps its hyper-parameters and its place in the runner; the iteration itself -- staging, bonus,
    returns, epochs x minibatches of loss / backward / clip / Adam -- runs on the fused path.  The optimiser state lives
    in ``FlatAdam`` and is exposed through ``optim_state_dict()`` in torch.optim.Adam's layout."""
    native = getattr(self, "_b200_native", None)
    model = getattr(self.agent.model, "module", self.agent.model)
    if native is None:
        native = PPO(discount=self.discount, learning_rate=self.learning_rate, value_loss_coeff=self.value_loss_coeff,
                     entropy_loss_coeff=self.entropy_loss_coeff, clip_grad_norm=self.clip_grad_norm,
                     gae_lambda=self.gae_lambda, minibatches=self.minibatches, epochs=self.epochs,
                     ratio_clip=self.ratio_clip, linear_lr_schedule=self.linear_lr_schedule,
                     normalize_advantage=self.normalize_advantage, normalize_reward=self.normalize_reward,
                     curiosity_type=self.curiosity_type)
        sd = self.optimizer.state_dict()
        native.initialize(model, n_itr=self.n_itr)
        if sd["state"]:
            native.load_optim_state_dict(sd)
        self._b200_native = native
        self.optim_state_dict = native.optim_state_dict
    info = optimize_agent_from_samples(native, self.agent, itr, samples)
    self.intrinsic_rewards = getattr(native, "intrinsic_rewards_numpy", None)
    self.ratio_clip = native.ratio_clip
    self.update_counter = native.update_counter
    return info, dict()
