"""install() rewires an importable rlpyt (only available in the build container, where the reference
is mounted at /root/reference) onto this package.  CPU only: checks the wiring, not the arithmetic."""
import os
import sys
import types

import pytest

REF = os.environ.get("CURIOSITY_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "rlpyt")), reason="reference checkout not mounted")
def test_install_patches_reference_modules():
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
        import curiosity_baselines_b200.integration as cbi
        patched = cbi.install()
        import rlpyt.models.pg.atari_lstm_model as alm
        import rlpyt.algos.pg.base as pgb
        import rlpyt.agents.pg.categorical as cat
        from curiosity_baselines_b200.curiosity.icm import ICM
        from curiosity_baselines_b200.curiosity.rnd import RND
        from curiosity_baselines_b200.algos.pg_base import process_returns
        from curiosity_baselines_b200.agents import hooks
        assert alm.ICM is ICM and alm.RND is RND
        assert pgb.PolicyGradientAlgo.process_returns is process_returns
        assert cat.RecurrentCategoricalPgAgentBase.curiosity_step is hooks.curiosity_step
        assert len(patched) >= 15
        # the reference's policy model now builds OUR curiosity module, with the reference's state_dict keys
        model = alm.AtariLstmModel(image_shape=(4, 84, 84), output_size=4,
                                   curiosity_kwargs=dict(curiosity_alg="icm", feature_encoding="idf_burda",
                                                         batch_norm=False, prediction_beta=1.0, forward_loss_wt=0.2))
        assert isinstance(model.curiosity_model, ICM)
        assert "curiosity_model.encoder.model.0.weight" in model.state_dict()
        # opt-in: the policy network itself
        patched = cbi.install(policy=True)
        import rlpyt.agents.pg.atari as atari
        from curiosity_baselines_b200.pg.atari_lstm_model import AtariLstmModel as Ours
        assert alm.AtariLstmModel is Ours and atari.AtariLstmModel is Ours
        assert atari.AtariLstmAgent.__init__.__defaults__ == (Ours,)
        assert Ours.rnn_state_cls is alm.RnnState
        agent = atari.AtariLstmAgent(model_kwargs=dict(curiosity_kwargs=dict(
            curiosity_alg="icm", feature_encoding="idf_burda", batch_norm=False, prediction_beta=1.0,
            forward_loss_wt=0.2)))
        assert agent.ModelCls is Ours
    finally:
        sys.path.remove(REF)
