"""Host-side logic of algos/ppo.py that needs no GPU: constructor defaults and the recurrent minibatch split
(rlpyt/algos/pg/ppo.py:20-70,133-140; rlpyt/utils/misc.py:14-24)."""
import inspect
import os
import sys
import types

import numpy as np
import pytest

from curiosity_baselines_b200.algos.ppo import PPO, PpoBatch

REF = os.environ.get("CURIOSITY_REFERENCE", "/root/reference")


def reference_mb_idxs(data_length, minibatch_size, rng):
    """utils/misc.py:14-24 restated (shuffle=True) on an explicit RandomState."""
    indexes = np.arange(data_length)
    rng.shuffle(indexes)
    return [indexes[s:s + minibatch_size] for s in range(0, data_length - minibatch_size + 1, minibatch_size)]


@pytest.mark.parametrize("B,mbs", [(8, 2), (8, 4), (10, 4), (7, 1), (5, 5)])
def test_minibatch_split_matches_iterate_mb_idxs(B, mbs):
    algo = PPO(minibatches=mbs)
    algo.rng = np.random.RandomState(42)
    ours = algo.minibatch_indices(B)
    ref = reference_mb_idxs(B, B // mbs, np.random.RandomState(42))
    assert len(ours) == len(ref) and all(np.array_equal(a, b) for a, b in zip(ours, ref))
    # ragged B: the tail that does not fill a minibatch is dropped, like the reference
    assert sum(len(a) for a in ours) == (B // (B // mbs)) * (B // mbs)


def test_batch_fields():
    assert PpoBatch._fields == ("observation", "next_observation", "prev_action", "action", "old_prob", "reward", "done",
                                "value", "bootstrap_value", "init_rnn_state")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "rlpyt")), reason="reference checkout not mounted")
def test_constructor_defaults_match_reference():
    for name, attrs in [("mpi4py", dict(MPI=None)), ("gym", {}), ("torchviz", dict(make_dot=None)),
                        ("matplotlib", {}), ("matplotlib.pyplot", {}),
                        ("matplotlib.colors", dict(ListedColormap=object)), ("matplotlib.lines", dict(Line2D=object))]:
        if name not in sys.modules:
            m = types.ModuleType(name)
            for k, v in attrs.items():
                setattr(m, k, v)
            sys.modules[name] = m
    sys.path.insert(0, REF)
    try:
        from rlpyt.algos.pg.ppo import PPO as RefPPO
        ref = inspect.signature(RefPPO.__init__).parameters
        ours = inspect.signature(PPO.__init__).parameters
        for k, prm in ours.items():
            if k == "self":
                continue
            assert k in ref and ref[k].default == prm.default, k
    finally:
        sys.path.remove(REF)


def test_policy_own_encoder_builds_with_the_reference_keys():
    """feature_encoding='none' (Conv2dHeadModel, the RND policy encoder, atari_lstm_model.py:94-112) and the
    curiosity-free policy: the holder carries the reference's parameter names and shapes; unsupported options are
    refused loudly instead of building something that would run elsewhere."""
    import recipes as R
    from curiosity_baselines_b200.pg.atari_lstm_model import AtariLstmModel
    m = AtariLstmModel((4, 84, 84), 4, curiosity_kwargs=dict(curiosity_alg="rnd", feature_encoding="none",
                                                             prediction_beta=1.0, drop_probability=0.25, gamma=0.99,
                                                             device="cpu"))
    keys = {k: tuple(v.shape) for k, v in m.state_dict().items() if not k.startswith("curiosity_model.")}
    assert keys == dict(R.policy_shapes(4, "none"))
    assert {k[len("curiosity_model."):]: tuple(v.shape) for k, v in m.state_dict().items()
            if k.startswith("curiosity_model.")} == dict(R.rnd_shapes())
    m = AtariLstmModel((4, 84, 84), 6)            # curiosity_alg='none' -> Conv2dHeadModel as well
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == dict(R.policy_shapes(6, "none"))
    with pytest.raises(NotImplementedError):
        AtariLstmModel((4, 84, 84), 4, use_maxpool=True)
