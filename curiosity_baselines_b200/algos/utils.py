"""Drop-in for rlpyt/algos/utils.py:8-40,104-112 -- same names, argument meaning and results
(bit-exact fp32), computed by one coalesced CUDA kernel parallel over B instead of a Python
loop over T.  Inputs may live on the CPU (as the sampler's buffers do) or on the GPU; the
result comes back on the device of ``reward`` / ``done``.
"""
import torch

from .. import _lib as L


def _dev():
    if not torch.cuda.is_available():
        raise L.CuriosityLibError("no CUDA device: the intrinsic-reward path has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _prep(x, T):
    """[T, ...] tensor (any device / dtype) -> contiguous fp32 CUDA [T, B]."""
    x = torch.as_tensor(x)
    x = x.to(_dev(), torch.float32, non_blocking=True)
    return x.reshape(T, -1).contiguous()


def discount_return(reward, done, bootstrap_value, discount, return_dest=None):
    """algos/utils.py:8-21."""
    reward_t = torch.as_tensor(reward)
    T = reward_t.shape[0]
    r = _prep(reward_t, T)
    B = r.shape[1]
    d = _prep(done, T)
    bv = torch.as_tensor(bootstrap_value).to(_dev(), torch.float32).reshape(-1).contiguous()
    if bv.numel() != B:
        raise ValueError(f"bootstrap_value has {bv.numel()} elements, expected {B}")
    out = torch.empty_like(r)
    L.call("cb_discount_return", L.ptr(r), L.ptr(d), L.ptr(bv), float(discount), L.ptr(out), T, B,
           L.stream())
    out = out.reshape(reward_t.shape).to(reward_t.device)
    if return_dest is not None:
        return_dest[:] = out
        return return_dest
    return out


def generalized_advantage_estimation(reward, value, done, bootstrap_value, discount, gae_lambda,
                                     advantage_dest=None, return_dest=None):
    """algos/utils.py:24-40 -> (advantage, return_)."""
    reward_t = torch.as_tensor(reward)
    T = reward_t.shape[0]
    r = _prep(reward_t, T)
    B = r.shape[1]
    v, d = _prep(value, T), _prep(done, T)
    bv = torch.as_tensor(bootstrap_value).to(_dev(), torch.float32).reshape(-1).contiguous()
    if bv.numel() != B:
        raise ValueError(f"bootstrap_value has {bv.numel()} elements, expected {B}")
    adv, ret = torch.empty_like(r), torch.empty_like(r)
    L.call("cb_gae", L.ptr(r), L.ptr(v), L.ptr(d), L.ptr(bv), float(discount), float(gae_lambda),
           L.ptr(adv), L.ptr(ret), T, B, L.stream())
    adv = adv.reshape(reward_t.shape).to(reward_t.device)
    ret = ret.reshape(reward_t.shape).to(reward_t.device)
    if advantage_dest is not None:
        advantage_dest[:] = adv
        adv = advantage_dest
    if return_dest is not None:
        return_dest[:] = ret
        ret = return_dest
    return adv, ret


def valid_from_done(done):
    """algos/utils.py:104-112."""
    done_t = torch.as_tensor(done)
    T = done_t.shape[0]
    d = _prep(done_t, T)
    valid = torch.empty_like(d)
    L.call("cb_valid_from_done", L.ptr(d), L.ptr(valid), T, d.shape[1], L.stream())
    return valid.reshape(done_t.shape).to(done_t.device)
