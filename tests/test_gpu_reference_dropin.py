"""The drop-in claim, driven by the REFERENCE's own code (round-1 verdict item 7): rlpyt's ``AtariLstmAgent`` +
``PPO.optimize_agent`` (rlpyt/algos/pg/ppo.py:72-190, unmodified, imported from /root/reference or from the vendored
``baseline/_ref``) run one iteration twice on the same synthetic ``Samples``:

  (a) untouched, on the CPU -- the reference path itself;
  (b) after ``integration.install(policy=True)`` with the agent on cuda:0 -- the reference's control flow, loss
      arithmetic and torch.optim.Adam, but every model forward / backward, the bonus, the returns and the hooks
      run in libcuriosity_b200.so.

Losses, gradient norms and the parameters after the update must agree within the north_star tolerances."""
import numpy as np
import pytest
import torch

import recipes as R
from oracle import refshim
from curiosity_baselines_b200 import _lib as L

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not refshim.available(), reason="reference neither mounted nor vendored")]

FIELDS = ("loss", "pi_loss", "value_loss", "entropy_loss", "inv_loss", "forward_loss", "entropy", "perplexity")


_REF_INIT = {}


def _params(kind, A, seed):
    if kind == "icm_bn":       # `-batch_norm`: the reference's own (seeded) initialisation, captured from its model once
        if kind not in _REF_INIT:
            raise RuntimeError("the reference pass must run first")
        return {k: v.clone() for k, v in _REF_INIT[kind].items()}
    if kind in ("icm", "disagreement"):
        pol = R.make_policy_params(A, seed)
        shapes = R.icm_shapes(A, "idf_burda") if kind == "icm" else R.disagreement_shapes(A, "idf_burda")
        pol.update({"curiosity_model." + k: v for k, v in R.make_params(shapes, seed + 2).items()})
        return pol
    if kind == "rnd":          # the policy brings its own 16/32-channel encoder (atari_lstm_model.py:94-101)
        pol = R.make_policy_params(A, seed, feature_encoding="none", conv_gain=0.5)
        pol.update({"curiosity_model." + k: v for k, v in R.make_params(R.rnd_shapes(), seed + 2, gain=1.4).items()})
        return pol
    if kind == "ndigo":
        pol = R.make_policy_params(A, seed, feature_encoding="idf_maze", conv_gain=1.0, c_in=3)
        pol.update({"curiosity_model." + k: v for k, v in R.make_params(R.ndigo_shapes(A), seed + 2).items()})
        return pol
    raise ValueError(kind)


def _samples(A, T, B, seed, kind="icm"):
    frames = (lambda s: R.make_binary_frames(s, T, B)) if kind == "ndigo" else (lambda s: R.make_frames(s, T, B))
    return refshim.make_samples(
        frames(seed + 3), frames(seed + 4), R.make_actions(seed + 6, T, B, A),
        R.make_actions(seed + 5, T, B, A), R.make_probs(seed + 7, T, B, A), R.make_normal(seed + 8, T, B),
        R.make_normal(seed + 9, B), torch.zeros(T, B), R.make_suffix_done(seed + 10, T, B, frac=0.3))


def _one_iteration(kind, ck, A, T, B, seed, cuda_idx):
    agent, algo = refshim.make_agent_and_algo(ck, A, T, B, cuda_idx=cuda_idx,
                                              image_shape=(3, 5, 5) if kind == "ndigo" else (4, 84, 84),
                                              ppo_kwargs=dict(epochs=2, minibatches=1, gae_lambda=0.95, learning_rate=1e-4))
    if kind == "icm_bn" and kind not in _REF_INIT:
        torch.manual_seed(seed)
        for mod in agent.model.modules():
            if hasattr(mod, "reset_parameters"):
                mod.reset_parameters()
        _REF_INIT[kind] = {k: v.detach().cpu().clone() for k, v in agent.model.state_dict().items()}
    agent.model.load_state_dict(_params(kind, A, seed))
    before = {k: v.detach().cpu().clone() for k, v in agent.model.state_dict().items()}
    s = _samples(A, T, B, seed, kind)
    agent.train_mode(0)
    info, _ = algo.optimize_agent(0, s)
    after = {k: v.detach().cpu().clone() for k, v in agent.model.state_dict().items()}
    out = {f: [float(x) for x in getattr(info, f)] for f in FIELDS if len(getattr(info, f))}
    out["gradNorm"] = [float(x) for x in info.gradNorm]
    return out, before, after, s.env.reward.clone(), np.array(algo.intrinsic_rewards), agent


@pytest.mark.parametrize("kind,native_algo", [("icm", False), ("disagreement", False), ("rnd", False), ("ndigo", False),
                                              ("icm", True), ("rnd", True), ("ndigo", True), ("icm_bn", False), ("icm_bn", True)])
def test_reference_optimize_agent_through_install(kind, native_algo):
    """native_algo: ``install(policy=True, algo=True)`` -- the runner-level call ``algo.optimize_agent(itr, samples)`` on
    host ``Samples`` lands on the fused iteration (algos.ppo.rlpyt_optimize_agent) instead of the reference's loop."""
    import curiosity_baselines_b200.integration as cbi
    refshim.import_reference()
    A, T, B, seed = (5 if kind == "ndigo" else 4), 16, 8, 55
    ck = dict(curiosity_alg=kind.split("_")[0], feature_encoding="idf_burda", batch_norm=kind.endswith("_bn"),
              prediction_beta=1.0, forward_loss_wt=0.2, device="cpu")
    if kind == "rnd":          # launch.py's kwargs for RND (drop_probability 0: the mask rand > 0 keeps every sample)
        ck = dict(curiosity_alg="rnd", feature_encoding="none", prediction_beta=1.0, drop_probability=0.0, gamma=0.99,
                  device="cpu")
    if kind == "ndigo":
        ck = dict(curiosity_alg="ndigo", feature_encoding="idf_maze", pred_horizon=3, batch_norm=False,
                  num_predictors=10, device="cpu")
    if kind == "disagreement":
        # the reference's dropout draws from torch's global generator; keep it out of the comparison
        import rlpyt.models.curiosity.disagreement as ref_dis
        orig_dropout = ref_dis.nn.functional.dropout
    torch.manual_seed(0)
    np.random.seed(0)
    if kind == "disagreement":
        ref_dis.nn.functional.dropout = lambda x, p=0.5, training=True, inplace=False: x
    try:
        ref, before, ref_after, ref_reward, ref_rint, ref_agent = _one_iteration(kind, ck, A, T, B, seed, None)
    finally:
        if kind == "disagreement":
            ref_dis.nn.functional.dropout = orig_dropout
    from rlpyt.models.pg.atari_lstm_model import AtariLstmModel as RefModel
    assert isinstance(ref_agent.model, RefModel)
    calls0 = L.launch_count()
    cbi.install(policy=True, algo=native_algo)
    try:
        from curiosity_baselines_b200.pg.atari_lstm_model import AtariLstmModel as Ours
        from curiosity_baselines_b200.curiosity.icm import Disagreement
        if kind == "disagreement":
            Disagreement._default_masks = lambda self, T, B: None          # same: no dropout on this side either
        torch.manual_seed(0)
        np.random.seed(0)
        ours, _, our_after, our_reward, our_rint, agent = _one_iteration(kind, ck, A, T, B, seed, 0)
        assert isinstance(agent.model, Ours) and next(agent.model.parameters()).is_cuda
    finally:
        if kind == "disagreement":
            del Disagreement._default_masks
        cbi.uninstall()
    assert L.launch_count() - calls0 > 100                       # the iteration ran in the library
    if native_algo:
        assert hasattr(agent, "_b200_staging") and agent._b200_staging.bytes_uploaded > 0
    import rlpyt.models.pg.atari_lstm_model as alm
    assert alm.AtariLstmModel is RefModel                        # uninstall() restored the reference
    # bonus added into the sampler's reward buffer in place (base.py:66), kept as numpy on the algo (:67)
    assert float((our_reward - ref_reward).abs().max()) <= 1e-2 * float(ref_reward.abs().max())
    assert our_rint.shape == ref_rint.shape == (T, B)
    for f in [f for f in FIELDS + ("gradNorm",) if f in ref]:
        assert len(ours[f]) == len(ref[f])
        for step, (a, b) in enumerate(zip(ours[f], ref[f])):
            tol = 1e-2 if step == 0 else 2e-2                    # step 1 sees weights updated by step 0's bf16 gradients
            assert abs(a - b) <= tol * max(abs(b), 1e-3), (f, step, a, b)
    # the update itself: direction and size of the parameter change over the two Adam steps
    num = den_a = den_b = 0.0
    for k in before:
        da, db = (our_after[k] - before[k]).double().flatten(), (ref_after[k] - before[k]).double().flatten()
        num += float(da @ db); den_a += float(da @ da); den_b += float(db @ db)
    assert num / (den_a * den_b) ** 0.5 > 0.98
    assert abs((den_a / den_b) ** 0.5 - 1) < 2e-2
