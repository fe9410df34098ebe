"""Generate tests/golden/*.npz by running the UNMODIFIED reference (read-only at
/root/reference) on seeded inputs.  Run in the build container only:

    python tests/golden/make_golden.py

The reference imports three modules that are not installed (and not used on the hot
path): mpi4py, gym, matplotlib (+ torchviz).  They are stubbed in sys.modules before
the import; no reference source is modified or copied.  Only seeds and outputs are
written -- inputs and weights are rebuilt from tests/golden/recipes.py.
"""
import os
import sys
import types
from collections import namedtuple

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import recipes as R  # noqa: E402

REF = os.environ.get("CURIOSITY_REFERENCE", "/root/reference")


def import_reference():
    def stub(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m
    stub("mpi4py", MPI=None)
    stub("gym")
    mpl = stub("matplotlib")
    mpl.pyplot = stub("matplotlib.pyplot")
    mpl.colors = stub("matplotlib.colors", ListedColormap=object)
    mpl.lines = stub("matplotlib.lines", Line2D=object)
    stub("torchviz", make_dot=None)
    sys.path.insert(0, REF)


def save(name, **arrays):
    out = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, len(out), "arrays")


def load_into(module, params):
    sd = module.state_dict()
    assert set(sd.keys()) == set(params.keys()), (set(sd) ^ set(params))
    for k in sd:
        assert tuple(sd[k].shape) == tuple(params[k].shape), k
    module.load_state_dict({k: v.clone() for k, v in params.items()})


def grad_summary(module, prefix="g."):
    """Per-parameter gradient: full tensor when small, else norm + a strided sample."""
    out = {}
    for k, p in module.named_parameters():
        if p.grad is None:
            continue
        g = p.grad.detach().reshape(-1)
        out[prefix + k + ".norm"] = g.norm().double()
        step = max(1, g.numel() // 4096)
        out[prefix + k + ".sample"] = g[::step].clone()
    return out


# ------------------------------------------------------------------ cases
def gen_returns():
    from rlpyt.algos.utils import (discount_return, generalized_advantage_estimation,
                                   valid_from_done)
    for tag, T, B, seed, done_fn in [("a", 16, 5, 11, R.make_random_done),
                                     ("b", 128, 33, 12, R.make_suffix_done),
                                     ("c", 1, 4, 13, R.make_random_done)]:
        reward = R.make_normal(seed, T, B)
        value = R.make_normal(seed + 100, T, B)
        boot = R.make_normal(seed + 200, B)
        done = done_fn(seed + 300, T, B)
        adv, ret = generalized_advantage_estimation(reward, value, done, boot, 0.99, 0.95)
        dret = discount_return(reward, done, boot, 0.99)
        valid = valid_from_done(done)
        save(f"returns_{tag}", T=T, B=B, seed=seed, advantage=adv, return_=ret,
             discount_return=dret, valid=valid)


def gen_process_returns():
    """algos/pg/base.py:53-108 with a stub agent that returns a preset bonus."""
    from rlpyt.algos.pg.base import PolicyGradientAlgo
    from rlpyt.utils.averages import RunningMeanStd, RewardForwardFilter
    T, B, seed = 24, 6, 21
    Step = namedtuple("Step", ["r_int", "agent_curiosity_info"])

    class Agent:
        recurrent = True

        def __init__(self):
            self.calls = 0

        def curiosity_step(self, kind, *args):
            self.calls += 1
            return Step(R.make_normal(seed + 7 * self.calls, T, B).abs() * 0.1, None)

    Env = namedtuple("Env", ["reward", "done", "observation", "next_observation"])
    AgentInfo = namedtuple("AgentInfo", ["value"])
    Ag = namedtuple("Ag", ["agent_info", "bootstrap_value", "action", "prev_action"])
    Samples = namedtuple("Samples", ["env", "agent"])
    for tag, norm_r, norm_a, lam in [("plain", False, False, 0.95), ("norm", True, True, 0.95),
                                     ("lam1", False, True, 1.0)]:
        algo = PolicyGradientAlgo.__new__(PolicyGradientAlgo)
        algo.curiosity_type = "icm"
        algo.agent = Agent()
        algo.normalize_reward, algo.normalize_advantage = norm_r, norm_a
        algo.gae_lambda, algo.discount = lam, 0.99
        algo.mid_batch_reset, algo.kernel_params = False, None
        algo.reward_ff, algo.reward_rms = RewardForwardFilter(0.99), RunningMeanStd()
        outs = {}
        for it in range(2):
            reward = R.make_normal(seed + 1000 * it, T, B) * 0.0  # no_extrinsic
            done = R.make_suffix_done(seed + 1000 * it + 1, T, B) > 0.5
            value = R.make_normal(seed + 1000 * it + 2, T, B)
            boot = R.make_normal(seed + 1000 * it + 3, 1, B)
            obs = torch.zeros(T, B, 1)
            act = torch.zeros(T, B, dtype=torch.long)
            s = Samples(Env(reward, done, obs, obs), Ag(AgentInfo(value), boot, act, act))
            ret, adv, valid = algo.process_returns(s)
            outs.update({f"ret{it}": ret, f"adv{it}": adv, f"valid{it}": valid,
                         f"reward_after{it}": reward})
        save(f"process_returns_{tag}", T=T, B=B, seed=seed, **outs)


def gen_rnd():
    from rlpyt.models.curiosity.rnd import RND
    T, B, seed = 6, 5, 31
    torch.manual_seed(0)
    m = RND(image_shape=(4, 84, 84), prediction_beta=1.0, drop_probability=0.25, gamma=0.99)
    params = R.make_params(R.rnd_shapes(), seed, gain=1.4)
    load_into(m, params)
    outs = {}
    for it in range(3):
        frames = R.make_frames(seed + 10 * it, T, B)
        done = R.make_suffix_done(seed + 10 * it + 1, T, B, frac=0.6)
        outs[f"bonus{it}"] = m.compute_bonus(frames, done)
        outs[f"obs_mean{it}"] = m.obs_rms.mean.copy()
        outs[f"obs_var{it}"] = m.obs_rms.var.copy()
        outs[f"obs_count{it}"] = m.obs_rms.count
        outs[f"rew_var{it}"] = m.rew_rms.var
        outs[f"rew_mean{it}"] = m.rew_rms.mean
        outs[f"rew_count{it}"] = m.rew_rms.count
        outs[f"rewems{it}"] = m.rew_rff.rewems.copy()
    # loss + gradients on the last batch, with the drop mask replayed through torch.rand
    frames = R.make_frames(seed + 20, T, B)
    valid = 1 - R.make_suffix_done(seed + 22, T, B, frac=0.5)
    mask = R.make_mask(seed + 23, (T, B), 0.25)
    real_rand = torch.rand
    torch.rand = lambda *a, **k: 0.125 + 0.5 * mask    # (u > 0.25) == mask
    try:
        loss = m.compute_loss(frames, valid)
    finally:
        torch.rand = real_rand
    loss.backward()
    outs["loss"] = loss.double()
    outs.update(grad_summary(m))
    save("rnd", T=T, B=B, seed=seed, **outs)


def gen_icm():
    from rlpyt.models.curiosity.icm import ICM
    for enc, A, T, B, seed in [("idf_burda", 4, 5, 3, 41), ("idf", 7, 4, 3, 42)]:
        m = ICM(image_shape=(4, 84, 84), action_size=A, feature_encoding=enc,
                prediction_beta=0.7, forward_loss_wt=0.2)
        params = R.make_params(R.icm_shapes(A, enc), seed)
        load_into(m, params)
        obs = R.make_frames(seed + 1, T, B)
        nxt = R.make_frames(seed + 2, T, B)
        act = R.onehot(R.make_actions(seed + 3, T, B, A), A)
        valid = 1 - R.make_suffix_done(seed + 4, T, B, frac=0.5)
        with torch.no_grad():
            bonus = m.compute_bonus(obs, nxt, act)
        inv, fwd = m.compute_loss(obs, nxt, act, valid)
        (inv + fwd).backward()
        save(f"icm_{enc}", T=T, B=B, A=A, seed=seed, bonus=bonus, inv_loss=inv.double(),
             fwd_loss=fwd.double(), **grad_summary(m))


def gen_disagreement():
    import torch.nn.functional as F
    from rlpyt.models.curiosity.disagreement import Disagreement
    enc, A, T, B, seed = "idf_burda", 7, 4, 3, 51
    m = Disagreement(image_shape=(4, 84, 84), action_size=A, feature_encoding=enc,
                     prediction_beta=1.0, forward_loss_wt=0.2)
    params = R.make_params(R.disagreement_shapes(A, enc), seed)
    load_into(m, params)
    obs = R.make_frames(seed + 1, T, B)
    nxt = R.make_frames(seed + 2, T, B)
    act = R.onehot(R.make_actions(seed + 3, T, B, A), A)
    valid = 1 - R.make_suffix_done(seed + 4, T, B, frac=0.5)
    with torch.no_grad():
        bonus = m.compute_bonus(obs, nxt, act)
    masks = [R.make_mask(seed + 10 + e, (T, B, 512), 0.2) for e in range(4)]
    it = iter(masks)
    real = F.dropout
    F.dropout = lambda x, p=0.5, **k: x * next(it) / (1 - p)     # replay the masks
    try:
        inv, fwd = m.compute_loss(obs, nxt, act, valid)
    finally:
        F.dropout = real
    (inv + fwd).backward()
    outs = dict(bonus=bonus, inv_loss=inv.double(), fwd_loss=fwd.double(), **grad_summary(m))
    m.zero_grad()
    F.dropout = lambda x, p=0.5, **k: x                          # dropout off
    try:
        inv2, fwd2 = m.compute_loss(obs, nxt, act, valid)
    finally:
        F.dropout = real
    outs.update(inv_loss_nodrop=inv2.double(), fwd_loss_nodrop=fwd2.double())
    save("disagreement", T=T, B=B, A=A, seed=seed, **outs)


def gen_ndigo():
    from rlpyt.models.curiosity.ndigo import NDIGO
    A, T, B, seed = 5, 14, 4, 61
    for H in (1, 3):
        m = NDIGO(image_shape=(3, 5, 5), action_size=A, horizon=H, feature_encoding="idf_maze",
                  num_predictors=10)
        params = R.make_params(R.ndigo_shapes(A), seed)
        load_into(m, params)
        obs = R.make_binary_frames(seed + 1, T, B)
        prev = R.onehot(R.make_actions(seed + 2, T, B, A), A)
        act = R.onehot(R.make_actions(seed + 3, T, B, A), A)
        with torch.no_grad():
            bonus = m.compute_bonus(obs, prev, act)
        loss = m.compute_loss(obs, prev, act, torch.ones(T, B))
        loss.backward()
        save(f"ndigo_h{H}", T=T, B=B, A=A, H=H, seed=seed, bonus=bonus, loss=loss.double(),
             **grad_summary(m))


def gen_stats():
    from rlpyt.utils.averages import RunningMeanStd, RewardForwardFilter
    seed = 71
    rms = RunningMeanStd(shape=(3,))
    outs = {}
    for it in range(4):
        x = R.make_normal(seed + it, 10 + it, 3).numpy().astype(np.float64) * (it + 1)
        rms.update(x)
        outs[f"mean{it}"], outs[f"var{it}"], outs[f"count{it}"] = rms.mean.copy(), rms.var.copy(), rms.count
    rff = RewardForwardFilter(0.9)
    for it in range(5):
        r = R.make_normal(seed + 50 + it, 7).numpy()
        mask = (R.make_normal(seed + 60 + it, 7).numpy() > -0.5).astype(np.float32)
        outs[f"rff{it}"] = rff.update(r, done=mask)
    save("stats", seed=seed, **outs)


def gen_policy():
    """AtariLstmModel.forward + PPO.loss (policy + ICM terms) through the reference's own agent
    (agents/pg/atari.py, agents/pg/categorical.py:36-56,106-130; algos/pg/ppo.py:192-242)."""
    from rlpyt.agents.pg.atari import AtariLstmAgent
    from rlpyt.agents.base import AgentInputs
    from rlpyt.algos.pg.ppo import PPO, IcmAgentCuriosityInputs
    from rlpyt.distributions.categorical import DistInfo
    from rlpyt.envs.base import EnvSpaces
    from rlpyt.models.pg.atari_lstm_model import RnnState
    from rlpyt.spaces.int_box import IntBox
    A, T, B, seed = 4, 6, 3, 81
    agent = AtariLstmAgent(model_kwargs=dict(curiosity_kwargs=dict(
        curiosity_alg="icm", feature_encoding="idf_burda", batch_norm=False, prediction_beta=1.0,
        forward_loss_wt=0.2)), no_extrinsic=True)
    agent.initialize(EnvSpaces(observation=IntBox(0, 255, shape=(4, 84, 84), dtype="uint8"), action=IntBox(0, A)),
                     share_memory=False, global_B=B, obs_stats=None, env_ranks=list(range(B)))
    params = R.make_policy_params(A, seed)
    params.update({"curiosity_model." + k: v for k, v in R.make_params(R.icm_shapes(A, "idf_burda"), seed + 2).items()})
    load_into(agent.model, params)
    obs = R.make_frames(seed + 3, T, B)
    nxt = R.make_frames(seed + 4, T, B)
    prev_action = R.make_actions(seed + 5, T, B, A)
    action = R.make_actions(seed + 6, T, B, A)
    old_prob = R.make_probs(seed + 7, T, B, A)
    ret = R.make_normal(seed + 8, T, B)
    adv = R.make_normal(seed + 9, T, B)
    valid = 1 - R.make_suffix_done(seed + 10, T, B, frac=0.5)
    h0 = R.make_normal(seed + 11, B, 1, 512) * 0.5          # [B,N,H] as ppo.py:124-126 slices it
    c0 = R.make_normal(seed + 12, B, 1, 512) * 0.5
    algo = PPO(curiosity_type="icm")
    algo.agent = agent
    outs = {}
    with torch.no_grad():
        fc = agent.model.conv(obs.float().view(T * B, 4, 84, 84))
        print("policy fc_out std", float(fc.std()), "abs max", float(fc.abs().max()))
        pi, v, (hn, cn) = agent.model(obs, R.onehot(prev_action, A), torch.zeros(T, B),
                                      (h0.transpose(0, 1).contiguous(), c0.transpose(0, 1).contiguous()))
    outs.update(pi=pi, value=v, hn=hn, cn=cn)
    agent_inputs = AgentInputs(observation=obs, prev_action=prev_action, prev_reward=torch.zeros(T, B))
    cur = IcmAgentCuriosityInputs(observation=obs.clone(), next_observation=nxt.clone(), action=action.clone(),
                                  valid=valid)
    loss, pi_loss, value_loss, entropy_loss, entropy, perplexity, (inv, fwd) = algo.loss(
        agent_inputs, cur, action, ret, adv, valid, DistInfo(prob=old_prob), RnnState(h=h0, c=c0))
    loss.backward()
    gn = torch.nn.utils.clip_grad_norm_(agent.model.parameters(), 1e9)
    outs.update(loss=loss.double(), pi_loss=pi_loss.double(), value_loss=value_loss.double(),
                entropy_loss=entropy_loss.double(), entropy=entropy.double(), perplexity=perplexity.double(),
                inv_loss=inv.double(), fwd_loss=fwd.double(), grad_norm=gn.double(),
                ratio_clip=algo.ratio_clip, value_loss_coeff=algo.value_loss_coeff,
                entropy_loss_coeff=algo.entropy_loss_coeff)
    outs.update(grad_summary(agent.model))
    save("policy_ppo_icm", T=T, B=B, A=A, seed=seed, **outs)


if __name__ == "__main__":
    import_reference()
    torch.set_num_threads(8)
    if len(sys.argv) > 1:
        for name in sys.argv[1:]:
            globals()["gen_" + name]()
        sys.exit(0)
    gen_returns()
    gen_stats()
    gen_process_returns()
    gen_rnd()
    gen_icm()
    gen_disagreement()
    gen_ndigo()
