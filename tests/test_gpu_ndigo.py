"""NDIGO on the GPU vs the reference's golden vectors (bonus, loss, gradients)."""
import pytest
import torch

import recipes as R
from helpers import golden, check_grads
from curiosity_baselines_b200.curiosity.ndigo import NDIGO

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_state_dict_keys_match_reference():
    m = NDIGO(image_shape=(3, 5, 5), action_size=5, horizon=1)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == dict(R.ndigo_shapes(5))


@pytest.mark.parametrize("H", [1, 3])
def test_golden(H):
    g = golden(f"ndigo_h{H}")
    T, B, A, seed = int(g["T"]), int(g["B"]), int(g["A"]), int(g["seed"])
    m = NDIGO(image_shape=(3, 5, 5), action_size=A, horizon=H, feature_encoding="idf_maze", num_predictors=10).to(DEV)
    m.load_state_dict(R.make_params(R.ndigo_shapes(A), seed))
    obs = R.make_binary_frames(seed + 1, T, B).to(DEV)
    prev = R.onehot(R.make_actions(seed + 2, T, B, A), A).to(DEV)
    act = R.onehot(R.make_actions(seed + 3, T, B, A), A).to(DEV)
    bonus = m.compute_bonus(obs, prev, act).cpu()
    ref = torch.from_numpy(g["bonus"])
    # the bonus is a difference of two BCE means of order 0.7: absolute tolerance 1e-2 of that scale
    assert float((bonus - ref).abs().max()) < 7e-3
    assert torch.equal(bonus == 0, ref == 0)
    loss = m.compute_loss(obs, prev, act, torch.ones(T, B, device=DEV))
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) / float(g["loss"]) < 1e-2
    grads = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    assert len(grads) == len(list(m.parameters()))
    bad = check_grads(grads, g, cos_min=0.98, norm_tol=8e-2)
    assert not bad, bad


def test_horizon_ten_fails_like_the_reference():
    m = NDIGO(image_shape=(3, 5, 5), action_size=5, horizon=10).to(DEV)
    obs = R.make_binary_frames(1, 14, 2).to(DEV)
    a = R.onehot(R.make_actions(2, 14, 2, 5), 5).to(DEV)
    with pytest.raises(AttributeError):
        m.compute_bonus(obs, a, a)
