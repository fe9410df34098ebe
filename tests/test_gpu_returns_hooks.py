"""process_returns (algos/pg/base.py:53-108) and the agent hooks (agents/pg/categorical.py:83-130)
against the reference's golden vectors, driven through stub algo / agent objects the way the
reference classes would call them."""
from collections import namedtuple
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import recipes as R
from helpers import golden
from oracle import curiosity_oracle as O
from curiosity_baselines_b200.algos.pg_base import process_returns
from curiosity_baselines_b200.agents import hooks
from curiosity_baselines_b200.curiosity.rnd import RND
from curiosity_baselines_b200.curiosity.icm import ICM

pytestmark = pytest.mark.gpu
DEV = "cuda"
Step = namedtuple("Step", ["r_int", "agent_curiosity_info"])
Env = namedtuple("Env", ["reward", "done", "observation", "next_observation"])
AgentInfo = namedtuple("AgentInfo", ["value"])
Ag = namedtuple("Ag", ["agent_info", "bootstrap_value", "action", "prev_action"])
Samples = namedtuple("Samples", ["env", "agent"])


@pytest.mark.parametrize("tag,nr,na,lam", [("plain", False, False, 0.95), ("norm", True, True, 0.95),
                                           ("lam1", False, True, 1.0)])
def test_process_returns_golden(tag, nr, na, lam):
    g = golden(f"process_returns_{tag}")
    T, B, seed = int(g["T"]), int(g["B"]), int(g["seed"])

    class Agent:
        recurrent = True
        calls = 0

        def curiosity_step(self, kind, *args):
            self.calls += 1
            return Step(R.make_normal(seed + 7 * self.calls, T, B).abs() * 0.1, None)

    algo = SimpleNamespace(curiosity_type="icm", agent=Agent(), normalize_reward=nr, normalize_advantage=na,
                           gae_lambda=lam, discount=0.99, mid_batch_reset=False, kernel_params=None)
    for it in range(2):
        reward = R.make_normal(seed + 1000 * it, T, B) * 0.0
        done = R.make_suffix_done(seed + 1000 * it + 1, T, B) > 0.5
        value = R.make_normal(seed + 1000 * it + 2, T, B)
        boot = R.make_normal(seed + 1000 * it + 3, 1, B)
        obs = torch.zeros(T, B, 1)
        act = torch.zeros(T, B, dtype=torch.long)
        s = Samples(Env(reward, done, obs, obs), Ag(AgentInfo(value), boot, act, act))
        ret, adv, valid = process_returns(algo, s)
        assert ret.device.type == "cpu"
        np.testing.assert_allclose(ret.numpy(), g[f"ret{it}"], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(adv.numpy(), g[f"adv{it}"], rtol=2e-5, atol=2e-5)
        assert np.array_equal(valid.numpy(), g[f"valid{it}"])
        np.testing.assert_allclose(reward.numpy(), g[f"reward_after{it}"], rtol=1e-6)     # in-place add (base.py:66)
        assert algo.intrinsic_rewards.shape == (T, B)


def _agent(model, A):
    return SimpleNamespace(model=SimpleNamespace(curiosity_model=model), device=torch.device(DEV),
                           distribution=SimpleNamespace(dim=A))


def test_hooks_rnd_host_buffers():
    """Host (pinned) uint8 stacks in, CPU tensors out -- the reference's calling convention."""
    T, B, seed = 6, 4, 5
    params = R.make_params(R.rnd_shapes(), seed, gain=1.4)
    m = RND(image_shape=(4, 84, 84), drop_probability=0.0).to(DEV)
    m.load_state_dict(params)
    agent = _agent(m, 4)
    frames = R.make_frames(seed, T, B).pin_memory()
    done = R.make_suffix_done(seed + 1, T, B, frac=0.5)
    step = hooks.curiosity_step(agent, "rnd", frames, done)
    assert step.r_int.device.type == "cpu" and step.r_int.shape == (T, B)
    o = O.RndOracle(drop_probability=0.0)
    ref = o.compute_bonus(params, frames, done)
    nz = ref.abs() > 0
    assert float(((step.r_int - ref).abs()[nz] / ref.abs()[nz]).max()) < 1e-2
    loss = hooks.curiosity_loss(agent, "rnd", frames, torch.ones(T, B))
    assert loss.device.type == "cpu" and loss.requires_grad
    loss.backward()
    assert m.forward_model[0].weight.grad is not None


def test_hooks_icm_index_actions():
    T, B, A, seed = 4, 3, 4, 9
    m = ICM(image_shape=(4, 84, 84), action_size=A, feature_encoding="idf_burda").to(DEV)
    params = R.make_params(R.icm_shapes(A, "idf_burda"), seed)
    m.load_state_dict(params)
    agent = _agent(m, A)
    obs, nxt = R.make_frames(seed + 1, T, B), R.make_frames(seed + 2, T, B)
    idx = R.make_actions(seed + 3, T, B, A)
    step = hooks.curiosity_step(agent, "icm", obs, nxt, idx)
    ref = O.icm_bonus(params, obs, nxt, R.onehot(idx, A), "idf_burda", 1.0)
    assert float(((step.r_int - ref).abs() / ref.abs()).max()) < 1e-2
    inv, fwd = hooks.curiosity_loss(agent, "icm", obs, nxt, idx, torch.ones(T, B))
    (inv + fwd).backward()
    assert inv.device.type == "cpu" and m.encoder.model[0].weight.grad is not None


def test_batch_staging_uploads_once_per_iteration_and_detects_rewrites():
    """agents/hooks.BatchStaging on the real call path: curiosity_step and curiosity_loss of one iteration share ONE
    upload of the frames (the loss then reuses the bonus's features without a device-side comparison or host sync);
    a sampler buffer rewritten in place for the next iteration -- same address, no ``_version`` bump -- is uploaded
    again; a batch staged ahead with ``prefetch_samples`` is picked up without a second upload."""
    T, B, seed = 6, 4, 21
    params = R.make_params(R.rnd_shapes(), seed, gain=1.4)
    m = RND(image_shape=(4, 84, 84), drop_probability=0.0).to(DEV)
    m.load_state_dict(params)
    m.reuse_features = True
    agent = _agent(m, 4)
    frames = R.make_frames(seed, T, B)                       # pageable, like rlpyt's numpy shared memory
    done = R.make_suffix_done(seed + 1, T, B, frac=0.5)
    st = hooks.staging(agent)
    st.pin_host, st.PIN_MIN_BYTES = True, 1 << 16            # page-lock the "sampler buffer" in place (it outlives st)
    per_batch = T * B * 84 * 84
    o = O.RndOracle(drop_probability=0.0)
    for it in range(3):
        if it:                                               # the sampler refills its buffer in place
            frames.numpy()[...] = R.make_frames(seed + 10 * it, T, B).numpy()
        before = st.bytes_uploaded
        step = hooks.curiosity_step(agent, "rnd", frames, done)
        loss = hooks.curiosity_loss(agent, "rnd", frames, torch.ones(T, B))
        loss2 = hooks.curiosity_loss(agent, "rnd", frames, torch.ones(T, B))       # second epoch
        assert per_batch <= st.bytes_uploaded - before < per_batch + 4096, it    # newest frames once (+ the tiny valid masks)
        ref = o.compute_bonus(params, frames, done)
        nz = ref.abs() > 0
        assert float(((step.r_int - ref).abs()[nz] / ref.abs()[nz]).max()) < 1e-2, it
        assert float(loss) == float(loss2)
    assert m.reuse_hits["full"] == 6 and m.reuse_hits["miss"] == 0
    # double-buffered runner: next batch staged during this iteration
    nxt = R.make_frames(seed + 99, T, B)
    Env2 = namedtuple("Env2", ["next_observation", "observation"])
    hooks.prefetch_samples(agent, "rnd", SimpleNamespace(env=Env2(nxt, nxt)))
    before = st.bytes_uploaded
    step = hooks.curiosity_step(agent, "rnd", nxt, done)
    assert st.bytes_uploaded - before < 4096 and st.hits >= 1
    ref = o.compute_bonus(params, nxt, done)
    nz = ref.abs() > 0
    assert float(((step.r_int - ref).abs()[nz] / ref.abs()[nz]).max()) < 1e-2
    st.release_host_registrations()


def test_icm_hooks_share_one_upload_per_tensor():
    T, B, A, seed = 4, 3, 4, 9
    m = ICM(image_shape=(4, 84, 84), action_size=A, feature_encoding="idf_burda").to(DEV)
    m.load_state_dict(R.make_params(R.icm_shapes(A, "idf_burda"), seed))
    m.reuse_features = True
    agent = _agent(m, A)
    obs, nxt = R.make_frames(seed + 1, T, B), R.make_frames(seed + 2, T, B)
    idx = R.make_actions(seed + 3, T, B, A)
    st = hooks.staging(agent)
    hooks.curiosity_step(agent, "icm", obs, nxt, idx)
    inv, fwd = hooks.curiosity_loss(agent, "icm", obs, nxt, idx, torch.ones(T, B))
    assert 2 * obs.numel() <= st.bytes_uploaded < 2 * obs.numel() + 4096 and st.hits >= 3      # frames x2, actions
    assert m.reuse_hits == {"full": 1, "miss": 0}


def test_refilled_persistent_buffer_is_not_served_stale():
    """a device buffer refilled through a raw pointer keeps its address AND its ``_version``; the feature
    cache must not trust either -- only BatchStaging tensors (never refilled) take the comparison-free path."""
    from curiosity_baselines_b200 import _lib as L
    T, B, seed = 4, 3, 33
    m = RND(image_shape=(4, 84, 84), drop_probability=0.0).to(DEV)
    m.load_state_dict(R.make_params(R.rnd_shapes(), seed, gain=1.4))
    m.reuse_features = True
    a, b = R.make_frames(seed, T, B).pin_memory(), R.make_frames(seed + 1, T, B).pin_memory()
    buf = a.to(DEV)
    done, ones = torch.zeros(T, B, device=DEV), torch.ones(T, B, device=DEV)
    m.compute_bonus(buf, done)
    version = buf._version
    L.call("cb_h2d_2d", buf.data_ptr(), buf[0, 0].numel(), b.data_ptr(), buf[0, 0].numel(), buf[0, 0].numel(), T * B, L.stream())
    assert buf._version == version                       # invisible to torch
    stale_or_not = m.compute_loss(buf, ones, ones)
    m2 = RND(image_shape=(4, 84, 84), drop_probability=0.0).to(DEV)
    m2.load_state_dict(m.state_dict())
    m2.compute_bonus(a.to(DEV), done)                   # same statistics as m
    fresh = m2.compute_loss(b.to(DEV), ones, ones)
    assert float(stale_or_not) == float(fresh) and m.reuse_hits["miss"] == 1


def _shifted_rollout(seed, T, B):
    frames = R.make_frames(seed, T + 1, B)
    return frames, frames[:T].clone(), frames[1:].clone()       # what a collector writes: next_obs[t] == obs[t+1]


@pytest.mark.parametrize("cls_name", ["icm", "disagreement"])
def test_shifted_views_run_the_encoder_once_and_match_two_passes(cls_name):
    """Observation de-duplication: when next_observations is provably observations shifted by one
    step (two views of one T+1 buffer) the encoder runs once over T+1 steps.  Bonus and losses equal the two-pass
    result bit for bit, gradients up to summation order and one bf16 add."""
    from curiosity_baselines_b200.curiosity.icm import Disagreement
    T, B, A, seed = 6, 5, 4, 61
    cls = ICM if cls_name == "icm" else Disagreement
    shapes = R.icm_shapes(A, "idf_burda") if cls_name == "icm" else R.disagreement_shapes(A, "idf_burda")
    params = R.make_params(shapes, seed)
    frames, _, _ = _shifted_rollout(seed + 1, T, B)
    buf = frames.to(DEV)
    obs, nxt = buf[:T], buf[1:]
    act = R.onehot(R.make_actions(seed + 3, T, B, A), A).to(DEV)
    valid = (1 - R.make_suffix_done(seed + 4, T, B, frac=0.3)).to(DEV)

    def run(dedup):
        m = cls(image_shape=(4, 84, 84), action_size=A, feature_encoding="idf_burda").to(DEV)
        m.load_state_dict(params)
        m.dedup_shifted_observations = dedup
        assert m.shifted(obs, nxt) and not m.shifted(obs, nxt.clone())
        bonus = m.compute_bonus(obs, nxt, act)
        inv, fwd = m.compute_loss(obs, nxt, act, valid, **({} if cls_name == "icm" else dict(dropout_masks=None)))
        (inv + fwd).backward()
        return m, bonus, inv.detach(), fwd.detach(), {k: p.grad.clone() for k, p in m.named_parameters()}

    m0, b0, i0, f0, g0 = run(False)
    m1, b1, i1, f1, g1 = run(True)
    assert getattr(m0, "dedup_hits", 0) == 0 and m1.dedup_hits == 2
    assert torch.equal(b0, b1) and torch.equal(i0, i1) and torch.equal(f0, f1)
    for k in g0:
        a, b = g0[k].double().flatten(), g1[k].double().flatten()
        assert float(a @ b / (a.norm() * b.norm())) > 0.9999 and abs(float(b.norm() / a.norm()) - 1) < 5e-3, k


def test_hooks_upload_t_plus_one_steps_when_the_sampler_contract_is_assumed():
    T, B, A, seed = 5, 3, 4, 71
    m = ICM(image_shape=(4, 84, 84), action_size=A, feature_encoding="idf_burda").to(DEV)
    params = R.make_params(R.icm_shapes(A, "idf_burda"), seed)
    m.load_state_dict(params)
    agent = _agent(m, A)
    _, obs, nxt = _shifted_rollout(seed + 1, T, B)
    idx = R.make_actions(seed + 3, T, B, A)
    st = hooks.staging(agent)
    st.assume_shifted = True
    step = hooks.curiosity_step(agent, "icm", obs, nxt, idx)
    inv, fwd = hooks.curiosity_loss(agent, "icm", obs, nxt, idx, torch.ones(T, B))
    per_step = B * 4 * 84 * 84
    assert (T + 1) * per_step <= st.bytes_uploaded < (T + 1) * per_step + 4096          # not 2T steps
    assert m.dedup_hits == 2
    ref = O.icm_bonus(params, obs, nxt, R.onehot(idx, A), "idf_burda", 1.0)
    assert float(((step.r_int - ref).abs() / ref.abs()).max()) < 1e-2
    inv_r, fwd_r = O.icm_loss(params, obs, nxt, R.onehot(idx, A), torch.ones(T, B), "idf_burda", 0.2)
    assert abs(float(inv) - float(inv_r)) / float(inv_r) < 1e-2 and abs(float(fwd) - float(fwd_r)) / float(fwd_r) < 1e-2
    # unrelated tensors: the host-side spot check refuses, two plain uploads, two encoder passes
    other = R.make_frames(seed + 9, T, B)
    before, hits = st.bytes_uploaded, m.dedup_hits
    hooks.curiosity_step(agent, "icm", obs, other, idx)
    assert st.bytes_uploaded - before >= 2 * T * per_step and m.dedup_hits == hits
