"""Host-side structure of the encoder holders (no GPU): the nn.Sequential layouts -- hence the ``state_dict`` keys a
``params.pkl`` checkpoint carries -- must be the reference's (rlpyt/models/curiosity/encoders.py:7-119), with and without
`-batch_norm`; the oracle's BatchNorm restatement must be torch's."""
import pytest
import torch

from oracle import curiosity_oracle as O
from oracle import refshim
from curiosity_baselines_b200.curiosity.encoders import BurdaHead, MazeHead, UniverseHead
from curiosity_baselines_b200.curiosity.icm import ICM, Disagreement


def _keys(m):
    return [(k, tuple(v.shape)) for k, v in m.state_dict().items()]


def test_batch_norm_layouts():
    b = BurdaHead((4, 84, 84), batch_norm=True)
    assert [type(x).__name__ for x in b.model] == ["Conv2d", "LeakyReLU", "BatchNorm2d"] * 3 + ["Flatten", "Linear"]
    assert b.model[2].num_features == 32 and b.model[5].num_features == 64 and b.model[8].num_features == 64
    u = UniverseHead((4, 84, 84), batch_norm=True)
    assert [type(x).__name__ for x in u.model] == ["Conv2d", "ELU", "BatchNorm2d"] * 5
    m = MazeHead((3, 5, 5), batch_norm=True)            # the reference's MazeHead ignores the flag
    assert not any(isinstance(x, torch.nn.BatchNorm2d) for x in m.model)
    assert "model.10.weight" in dict(_keys(b)) and "model.7.weight" in dict(_keys(BurdaHead((4, 84, 84))))


def test_batch_norm_turns_reuse_and_dedup_off():
    for cls in (ICM, Disagreement):
        m = cls(image_shape=(4, 84, 84), action_size=4, feature_encoding="idf_burda", batch_norm=True)
        m.reuse_features = True
        assert m._batch_stats and not m.reuse_features
        m = cls(image_shape=(4, 84, 84), action_size=4, feature_encoding="idf_burda", batch_norm=False)
        m.reuse_features = True
        assert not m._batch_stats and m.reuse_features
        m = cls(image_shape=(3, 5, 5), action_size=5, feature_encoding="idf_maze", batch_norm=True)
        assert not m._batch_stats


@pytest.mark.parametrize("cls,fn", [(BurdaHead, O.burda_head), (UniverseHead, O.universe_head)])
def test_oracle_batch_norm_is_torchs(cls, fn):
    torch.manual_seed(0)
    m = cls((4, 84, 84), batch_norm=True)
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.weight.uniform_(0.5, 1.5); mod.bias.uniform_(-0.5, 0.5)
                mod.running_mean.uniform_(-1, 1); mod.running_var.uniform_(0.5, 2)
    p = {k: v.detach().clone() for k, v in m.state_dict().items()}
    x = torch.rand(5, 4, 84, 84) * 255
    for training in (True, False):
        m.train(training)
        y = m.model(x).reshape(5, -1)
        assert torch.allclose(fn(x, p, bn_training=training), y.detach(), rtol=1e-5, atol=1e-5)
        if training:                                     # (the module's forward moved its running statistics)
            m.load_state_dict(p)


@pytest.mark.skipif(not refshim.available(), reason="reference neither mounted nor vendored")
@pytest.mark.parametrize("bn", [False, True])
def test_state_dict_keys_are_the_references(bn):
    refshim.import_reference()
    from rlpyt.models.curiosity.encoders import BurdaHead as RB, UniverseHead as RU
    from rlpyt.models.curiosity.icm import ICM as RICM
    assert _keys(RB((4, 84, 84), batch_norm=bn)) == _keys(BurdaHead((4, 84, 84), batch_norm=bn))
    assert _keys(RU((4, 84, 84), batch_norm=bn)) == _keys(UniverseHead((4, 84, 84), batch_norm=bn))
    r = RICM(image_shape=(4, 84, 84), action_size=4, feature_encoding="idf_burda", batch_norm=bn)
    o = ICM(image_shape=(4, 84, 84), action_size=4, feature_encoding="idf_burda", batch_norm=bn)
    assert _keys(r) == _keys(o)
