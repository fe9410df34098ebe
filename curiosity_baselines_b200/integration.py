"""Drop-in wiring: make an importable rlpyt (the reference checkout) run its intrinsic-reward hot
path on this package without touching launch.py.

    import curiosity_baselines_b200.integration as cbi
    cbi.install()            # before launch.start_experiment(...)

What gets replaced (reference file:line -> ours):
  rlpyt/models/curiosity/{icm,disagreement,rnd,ndigo}.py classes      -> curiosity.{ICM,Disagreement,RND,NDIGO}
  the same names inside rlpyt/models/pg/atari_lstm_model.py:45-78     (it constructs them by name)
  rlpyt/algos/utils.py:8-40,104-112 discount_return / GAE / valid_from_done -> algos.utils
  PolicyGradientAlgo.process_returns  (algos/pg/base.py:53-108)       -> algos.pg_base.process_returns
  RecurrentCategoricalPgAgentBase.curiosity_step / curiosity_loss
                                      (agents/pg/categorical.py:83-130) -> agents.hooks
With ``policy=True`` the policy network is replaced as well:
  rlpyt/models/pg/atari_lstm_model.py:19-154 AtariLstmModel           -> pg.atari_lstm_model.AtariLstmModel
  (also the ``ModelCls`` defaults of AtariLstmAgent / AlternatingAtariLstmAgent, agents/pg/atari.py:26-36)
The reference's ``PPO.loss`` then back-propagates through our kernels via the model's autograd boundary.  It is
opt-in because the library has no CPU path: use it with samplers that evaluate the agent on the GPU
(``GpuSampler``); ``CpuSampler`` workers evaluate a CPU copy of the model and need the reference class.
With ``policy=True, algo=True`` the whole iteration is replaced as well:
  rlpyt/algos/pg/ppo.py:72-190 PPO.optimize_agent                     -> algos.ppo.rlpyt_optimize_agent
  (the sampler's host buffers are staged once per iteration -- the reference clones both frame tensors on the host
  and uploads ``observation`` twice -- then bonus, returns and every epoch x minibatch gradient step run on the fused
  path with one FlatAdam over policy U curiosity parameters; the runner still gets the reference's ``OptInfo``).
"""
import importlib

_originals = []        # (object, attribute, previous value) in patch order, for uninstall()


def uninstall():
    """Undo every ``install()``: put the reference's own classes / functions back (newest patch first)."""
    from .curiosity.icm import ICM, Disagreement
    from .curiosity.rnd import RND
    while _originals:
        target, leaf, old = _originals.pop()
        setattr(target, leaf, old)
    RND.default_reuse_features = ICM.default_reuse_features = Disagreement.default_reuse_features = False
    from .agents import hooks
    hooks.BatchStaging.default_pin_host = False
    hooks.BatchStaging.default_assume_shifted = False


def install(strict=True, reuse_features=True, policy=False, pin_sampler_buffers=False, algo=False,
            shifted_next_observation=False):
    """Patch rlpyt in place; returns the list of patched attributes.  ``strict=False`` skips modules
    that fail to import (e.g. optional dependencies missing).  ``reuse_features``: RND models built afterwards reuse
    a batch's features between compute_bonus and compute_loss when the frames are verified equal (the reference
    passes samples.env.next_observation to both, base.py:72-73 / ppo.py:107-110) -- exact, ~30 % less work."""
    from .curiosity.icm import ICM, Disagreement
    from .curiosity.rnd import RND
    from .curiosity.ndigo import NDIGO
    from .algos import utils as U
    from .algos.pg_base import process_returns
    from .agents import hooks

    patched = []
    # rlpyt allocates its sample buffers once per run as numpy shared memory (samplers/buffer.py): page-locking them in
    # place lets the hooks' uploads run at PCIe speed and asynchronously (agents/hooks.BatchStaging)
    hooks.BatchStaging.default_pin_host = bool(pin_sampler_buffers)
    # rlpyt's collectors fill next_observation[t] and observation[t+1] from the same array (cpu/collectors.py:118,156):
    # upload T+1 steps instead of 2T and run the ICM / Disagreement encoder once over them (opt-in: sampler contract)
    hooks.BatchStaging.default_assume_shifted = bool(shifted_next_observation)
    RND.default_reuse_features = bool(reuse_features)
    ICM.default_reuse_features = Disagreement.default_reuse_features = bool(reuse_features)

    def patch(modname, **attrs):
        try:
            mod = importlib.import_module(modname)
        except Exception:
            if strict:
                raise
            return
        for k, v in attrs.items():
            target = mod
            *path, leaf = k.split(".")
            for part in path:
                target = getattr(target, part)
            _originals.append((target, leaf, getattr(target, leaf)))
            setattr(target, leaf, v)
            patched.append(f"{modname}.{k}")

    patch("rlpyt.models.curiosity.icm", ICM=ICM)
    patch("rlpyt.models.curiosity.disagreement", Disagreement=Disagreement)
    patch("rlpyt.models.curiosity.rnd", RND=RND)
    patch("rlpyt.models.curiosity.ndigo", NDIGO=NDIGO)
    patch("rlpyt.models.pg.atari_lstm_model", ICM=ICM, Disagreement=Disagreement, RND=RND, NDIGO=NDIGO)
    patch("rlpyt.algos.utils", discount_return=U.discount_return,
          generalized_advantage_estimation=U.generalized_advantage_estimation, valid_from_done=U.valid_from_done)
    patch("rlpyt.algos.pg.base", discount_return=U.discount_return,
          generalized_advantage_estimation=U.generalized_advantage_estimation, valid_from_done=U.valid_from_done,
          **{"PolicyGradientAlgo.process_returns": process_returns})
    patch("rlpyt.agents.pg.categorical",
          **{"RecurrentCategoricalPgAgentBase.curiosity_step": hooks.curiosity_step,
             "RecurrentCategoricalPgAgentBase.curiosity_loss": hooks.curiosity_loss})
    if policy:
        from .pg import atari_lstm_model as ours
        try:
            ref = importlib.import_module("rlpyt.models.pg.atari_lstm_model")
            _originals.append((ours.AtariLstmModel, "rnn_state_cls", ours.AtariLstmModel.rnn_state_cls))
            ours.AtariLstmModel.rnn_state_cls = ref.RnnState       # a namedarraytuple: the agents slice / transpose it
        except Exception:
            if strict:
                raise
        patch("rlpyt.models.pg.atari_lstm_model", AtariLstmModel=ours.AtariLstmModel)
        patch("rlpyt.agents.pg.atari", AtariLstmModel=ours.AtariLstmModel)
        try:
            atari = importlib.import_module("rlpyt.agents.pg.atari")
            for cls in (atari.AtariLstmAgent, atari.AlternatingAtariLstmAgent):
                _originals.append((cls.__init__, "__defaults__", cls.__init__.__defaults__))
                cls.__init__.__defaults__ = (ours.AtariLstmModel,)
                patched.append(f"rlpyt.agents.pg.atari.{cls.__name__}.__init__ (ModelCls default)")
        except Exception:
            if strict:
                raise
    if algo:
        if not policy:
            raise ValueError("install(algo=True) runs the fused PPO iteration on our policy network: needs policy=True")
        from .algos.ppo import rlpyt_optimize_agent
        patch("rlpyt.algos.pg.ppo", **{"PPO.optimize_agent": rlpyt_optimize_agent})
    return patched
