"""No hot shape selects the SIMT engine (round-1 verdict item 8): with the engine forced to 'umma' -- the library then
REFUSES any GEMM-shaped call without a tcgen05 specialisation instead of falling back -- every model family of the path
runs a bonus, a loss and a backward pass.  (The SIMT engine stays in the library as the on-device cross-check used by
tests/test_gpu_layers.py.)"""
import pytest
import torch

import recipes as R
from curiosity_baselines_b200 import layers as LY
from curiosity_baselines_b200.algos.ppo import PPO
from curiosity_baselines_b200.curiosity.icm import ICM, Disagreement
from curiosity_baselines_b200.curiosity.ndigo import NDIGO
from curiosity_baselines_b200.curiosity.rnd import RND
from curiosity_baselines_b200.pg.atari_lstm_model import AtariLstmModel

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture
def umma_only():
    LY.set_engine("umma")
    yield
    LY.set_engine("auto")


@pytest.mark.parametrize("enc,A,shape", [("idf", 7, (4, 84, 84)), ("idf_burda", 12, (4, 84, 84)), ("idf_maze", 5, (3, 5, 5))])
@pytest.mark.parametrize("cls_name", ["icm", "disagreement"])
def test_icm_like_models(umma_only, cls_name, enc, A, shape):
    T, B, seed = 4, 3, 5
    cls = ICM if cls_name == "icm" else Disagreement
    m = cls(image_shape=shape, action_size=A, feature_encoding=enc).to(DEV)
    frames = (lambda s: R.make_binary_frames(s, T, B)) if enc == "idf_maze" else (lambda s: R.make_frames(s, T, B))
    obs, nxt = frames(seed).to(DEV), frames(seed + 1).to(DEV)
    act = R.onehot(R.make_actions(seed + 2, T, B, A), A).to(DEV)
    bonus = m.compute_bonus(obs, nxt, act)
    inv, fwd = m.compute_loss(obs, nxt, act, torch.ones(T, B, device=DEV))
    (inv + fwd).backward()
    assert torch.isfinite(bonus).all() and all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())


def test_rnd(umma_only):
    T, B = 4, 3
    m = RND(image_shape=(4, 84, 84), drop_probability=0.0).to(DEV)
    frames = R.make_frames(1, T, B).to(DEV)
    bonus = m.compute_bonus(frames, torch.zeros(T, B, device=DEV))
    m.compute_loss(frames, torch.ones(T, B, device=DEV)).backward()
    assert torch.isfinite(bonus).all() and all(torch.isfinite(p.grad).all() for p in m.forward_model.parameters())


def test_ndigo(umma_only):
    T, B, A = 14, 4, 5
    m = NDIGO(image_shape=(3, 5, 5), action_size=A, horizon=3).to(DEV)
    obs = R.make_binary_frames(1, T, B).to(DEV)
    a = R.onehot(R.make_actions(2, T, B, A), A).to(DEV)
    bonus = m.compute_bonus(obs, a, a)
    m.compute_loss(obs, a, a, torch.ones(T, B, device=DEV)).backward()
    assert torch.isfinite(bonus).all() and all(torch.isfinite(p.grad).all() for p in m.parameters())


@pytest.mark.parametrize("ck,A", [
    (dict(curiosity_alg="rnd", feature_encoding="none", prediction_beta=1.0, drop_probability=0.0, gamma=0.99), 4),
    (dict(curiosity_alg="icm", feature_encoding="idf", batch_norm=False, prediction_beta=1.0, forward_loss_wt=0.2), 12),
    (dict(curiosity_alg="none"), 18)])
def test_policy_ppo_step(umma_only, ck, A):
    T, B, seed = 4, 3, 9
    m = AtariLstmModel((4, 84, 84), A, curiosity_kwargs=ck).to(DEV)
    algo = PPO(curiosity_type=ck["curiosity_alg"], clip_grad_norm=1.0)
    algo.initialize(m, n_itr=10)
    obs, nxt = R.make_frames(seed, T, B).to(DEV), R.make_frames(seed + 1, T, B).to(DEV)
    prev, act = R.make_actions(seed + 2, T, B, A).to(DEV), R.make_actions(seed + 3, T, B, A).to(DEV)
    cur = {"rnd": (nxt,), "icm": (obs, nxt, R.onehot(act.cpu(), A).to(DEV)), "none": None}[ck["curiosity_alg"]]
    info = algo.loss_and_grads_(obs, prev, act, R.make_normal(seed + 4, T, B).to(DEV), R.make_normal(seed + 5, T, B).to(DEV),
                                torch.ones(T, B, device=DEV), R.make_probs(seed + 6, T, B, A).to(DEV), None, cur, None)
    algo.step()
    assert torch.isfinite(info.loss)
