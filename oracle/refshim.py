"""Locate and import the UNMODIFIED reference (Improbable-AI/curiosity_baselines, package ``rlpyt``).
TEST / BENCH INFRASTRUCTURE ONLY -- nothing under curiosity_baselines_b200/ imports this.

Search order: $CURIOSITY_REFERENCE, /root/reference (the read-only checkout of the build container),
``baseline/_ref`` (the git-ignored copy made by ``tools/vendor_ref.py`` so that the reference travels to the GPU
box with the repo snapshot).  The reference imports four packages that are absent from this image and unused on the
hot path (mpi4py, gym, matplotlib, torchviz); empty stand-ins are registered in ``sys.modules`` first
("importable with stubs").  No reference source is modified.
"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VENDORED = os.path.join(ROOT, "baseline", "_ref")

STUBS = [("mpi4py", dict(MPI=None)), ("gym", {}), ("torchviz", dict(make_dot=None)), ("matplotlib", {}),
         ("matplotlib.pyplot", {}), ("matplotlib.colors", dict(ListedColormap=object)),
         ("matplotlib.lines", dict(Line2D=object))]


def locate():
    """Directory that contains the reference's ``rlpyt`` package, or None."""
    for cand in (os.environ.get("CURIOSITY_REFERENCE"), "/root/reference", VENDORED):
        if cand and os.path.isfile(os.path.join(cand, "rlpyt", "algos", "pg", "ppo.py")):
            return cand
    return None


def available():
    return locate() is not None


def kind():
    """'reference' when the code that will be imported is the reference's own (checkout or vendored copy)."""
    return "reference" if available() else None


def import_reference():
    """Put the reference on sys.path (once) with the stand-ins registered; returns its directory."""
    where = locate()
    if where is None:
        raise ImportError("reference not found: neither /root/reference nor baseline/_ref (run tools/vendor_ref.py "
                          "in the build container)")
    for name, attrs in STUBS:
        if name not in sys.modules:
            try:
                __import__(name)
                continue
            except Exception:
                pass
            m = types.ModuleType(name)
            for k, v in attrs.items():
                setattr(m, k, v)
            sys.modules[name] = m
            if "." in name:
                setattr(sys.modules[name.split(".")[0]], name.split(".")[1], m)
    if where not in sys.path:
        sys.path.insert(0, where)
    return where


# ------------------------------------------------------------------------------------------------------------------
# Harness: the reference's own agent + algorithm objects on synthetic ``Samples`` (SURVEY 8c recipe)
# ------------------------------------------------------------------------------------------------------------------
def make_agent_and_algo(curiosity_kwargs, A, T, B, n_itr=10, ppo_kwargs=None, image_shape=(4, 84, 84), cuda_idx=None):
    """rlpyt AtariLstmAgent + PPO initialised the way runners/minibatch_rl.py does it (without a sampler)."""
    import_reference()
    from rlpyt.agents.pg.atari import AtariLstmAgent
    from rlpyt.algos.pg.ppo import PPO
    from rlpyt.envs.base import EnvSpaces
    from rlpyt.spaces.int_box import IntBox
    from rlpyt.utils.quick_args import save__init__args  # noqa: F401  (import check)
    from rlpyt.samplers.collections import BatchSpec
    agent = AtariLstmAgent(model_kwargs=dict(curiosity_kwargs=dict(curiosity_kwargs)), no_extrinsic=True)
    if image_shape[-1] > 16:
        obs_space = IntBox(0, 255, shape=tuple(image_shape), dtype="uint8")
    else:                                   # PyColab mazes: float {0,1} masks (mazeworld_env.py:60-74)
        from rlpyt.spaces.float_box import FloatBox
        obs_space = FloatBox(0.0, 1.0, shape=tuple(image_shape))
    agent.initialize(EnvSpaces(observation=obs_space, action=IntBox(0, A)),
                     share_memory=False, global_B=B, obs_stats=None, env_ranks=list(range(B)))
    if cuda_idx is not None:
        agent.to_device(cuda_idx)
    kw = dict(curiosity_type=curiosity_kwargs["curiosity_alg"])
    kw.update(ppo_kwargs or {})
    algo = PPO(**kw)
    algo.initialize(agent=agent, n_itr=n_itr, batch_spec=BatchSpec(T, B), mid_batch_reset=False)
    return agent, algo


def make_samples(obs, next_obs, action, prev_action, old_prob, value, bootstrap_value, reward, done, h0=None, c0=None):
    """rlpyt ``Samples`` (samplers/collections.py) from time-major CPU tensors; h0 / c0 [T,B,1,H] or None (zeros)."""
    import torch
    import_reference()
    from rlpyt.agents.pg.base import AgentInfoRnn
    from rlpyt.distributions.categorical import DistInfo
    from rlpyt.models.pg.atari_lstm_model import RnnState
    from rlpyt.samplers.collections import Samples, AgentSamplesBsv, EnvSamples
    T, B = action.shape
    if h0 is None:
        h0 = torch.zeros(T, B, 1, 512)
        c0 = torch.zeros(T, B, 1, 512)
    info = AgentInfoRnn(dist_info=DistInfo(prob=old_prob), value=value, prev_rnn_state=RnnState(h=h0, c=c0))
    agent_s = AgentSamplesBsv(action=action, prev_action=prev_action, agent_info=info,
                              bootstrap_value=bootstrap_value.reshape(1, B))
    env_s = EnvSamples(reward=reward, prev_reward=torch.zeros_like(reward), observation=obs, next_observation=next_obs,
                       done=done.bool(), env_info=None)
    return Samples(agent=agent_s, env=env_s)
