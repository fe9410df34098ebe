"""Drop-in for ``PolicyGradientAlgo.process_returns`` (rlpyt/algos/pg/base.py:53-108).

``process_returns(algo, samples)`` has the reference's signature and side effects (the bonus is
added into ``samples.env.reward`` in place, ``algo.intrinsic_rewards`` is kept as numpy) and can
be bound onto the reference class (``PolicyGradientAlgo.process_returns = process_returns``).
The arithmetic -- reward filter, running moments, GAE / discounted return, valid mask,
advantage normalisation -- runs in libcuriosity_b200.so on the GPU.
"""
import numpy as np
import torch

from .. import _lib as L
from .. import dist
from ..utils.averages import RunningMeanStd, RewardForwardFilter
from .utils import _dev


class ReturnsState:
    """The two pieces of state the reference keeps on the algo when normalize_reward is set
    (ppo.py:47-49): a RewardForwardFilter and a scalar RunningMeanStd, device-resident here."""

    def __init__(self, discount):
        self.reward_ff = RewardForwardFilter(discount)
        self.reward_rms = RunningMeanStd()


def returns_on_device(reward, done, value, bootstrap_value, discount, gae_lambda,
                      normalize_reward=False, normalize_advantage=False, use_valid=True,
                      state=None):
    """Device core of process_returns after the bonus was added.  All inputs fp32 CUDA
    [T,B] (bootstrap [B]).  Returns (return_, advantage, valid-or-None, reward_used)."""
    T, B = reward.shape
    st = L.stream()
    if normalize_reward:      # base.py:77-82 -- filter over T without mask, moments over all N
        zeros = torch.zeros_like(done)
        total = state.reward_ff.update_sequence(reward, zeros)
        state.reward_rms.to(reward.device).update_from_values(total)
        scaled = torch.empty_like(reward)
        L.call("cb_rnd_scale_bonus", L.ptr(reward), L.ptr(zeros), L.ptr(state.reward_rms._var), 1.0,
               L.ptr(scaled), T * B, st)
        reward = scaled
    adv, ret = torch.empty_like(reward), torch.empty_like(reward)
    if gae_lambda == 1:       # base.py:84-86
        L.call("cb_discount_return", L.ptr(reward), L.ptr(done), L.ptr(bootstrap_value),
               float(discount), L.ptr(ret), T, B, st)
        adv = ret - value     # one elementwise op; kept in torch as in the reference line
    else:
        L.call("cb_gae", L.ptr(reward), L.ptr(value), L.ptr(done), L.ptr(bootstrap_value),
               float(discount), float(gae_lambda), L.ptr(adv), L.ptr(ret), T, B, st)
    valid = None
    if use_valid:             # base.py:90-93
        valid = torch.empty_like(done)
        L.call("cb_valid_from_done", L.ptr(done), L.ptr(valid), T, B, st)
    if normalize_advantage:   # base.py:95-103
        # masked {sum, sum^2, count} -> all-reduce -> apply: a sharded batch normalises with the global moments
        out = torch.empty_like(adv)
        sums = torch.empty(3, dtype=torch.float64, device=adv.device)
        ws = torch.empty(1536, dtype=torch.float64, device=adv.device)
        L.call("cb_masked_moments", L.ptr(adv), L.ptr(valid), T * B, L.ptr(sums), L.ptr(ws), st)
        dist.all_reduce_sum_(sums)
        L.call("cb_normalize_apply", L.ptr(adv), L.ptr(sums), 1e-6, L.ptr(out), T * B, st)
        adv = out
    return ret, adv, valid, reward


def process_returns(self, samples):
    """base.py:53-108, same observable behaviour."""
    dev = _dev()
    reward_host = samples.env.reward
    reward = reward_host.to(dev, torch.float32)
    T, B = reward.shape[:2]
    done = samples.env.done.to(dev, torch.float32).reshape(T, B).contiguous()
    value = samples.agent.agent_info.value.to(dev, torch.float32).reshape(T, B).contiguous()
    bv = samples.agent.bootstrap_value.to(dev, torch.float32).reshape(-1).contiguous()

    ctype = getattr(self, "curiosity_type", "none")
    r_int = None
    if ctype in ("icm", "disagreement"):
        r_int, _ = self.agent.curiosity_step(ctype, samples.env.observation, samples.env.next_observation,
                                             samples.agent.action)
    elif ctype == "ndigo":
        r_int, _ = self.agent.curiosity_step(ctype, samples.env.observation, samples.agent.prev_action,
                                             samples.agent.action)
    elif ctype == "rnd":
        r_int, _ = self.agent.curiosity_step(ctype, samples.env.next_observation, done)
    if r_int is not None:
        r_dev = r_int.to(dev, torch.float32).reshape(T, B)
        reward = (reward.reshape(T, B) + r_dev).contiguous()
        reward_host.copy_(reward.reshape(reward_host.shape))     # in-place add on the sampler buffer (base.py:66)
        self.intrinsic_rewards = r_int.detach().cpu().numpy()
    reward = reward.reshape(T, B).contiguous()

    if getattr(self, "normalize_reward", False) and not isinstance(getattr(self, "_b200_state", None), ReturnsState):
        self._b200_state = ReturnsState(self.discount)
    use_valid = (not self.mid_batch_reset) or self.agent.recurrent
    ret, adv, valid, _ = returns_on_device(
        reward, done, value, bv, self.discount, self.gae_lambda,
        normalize_reward=getattr(self, "normalize_reward", False),
        normalize_advantage=getattr(self, "normalize_advantage", False),
        use_valid=use_valid, state=getattr(self, "_b200_state", None))
    if getattr(self, "kernel_params", None) is not None:    # base.py:105-106 (rarely used)
        a = adv.cpu().numpy()
        adv = torch.tensor(np.piecewise(a, [abs(a) < self.mu, abs(a) >= self.mu],
                                        [self.kernel_line, self.kernel_gauss]), device=dev)
    host = reward_host.device
    shape = reward_host.shape
    ret_h, adv_h = ret.reshape(shape).to(host), adv.reshape(shape).to(host)
    valid_h = valid.reshape(shape).to(host) if valid is not None else None
    if host.type == "cpu" and valid_h is not None and hasattr(self.agent, "curiosity_loss"):
        # the caller hands ``valid`` straight back to agent.curiosity_loss (ppo.py:93-110): remember its device original
        try:
            from ..agents import hooks
            hooks.staging(self.agent).alias(valid_h, valid.reshape(shape))
        except Exception:
            pass
    return ret_h, adv_h, valid_h
