"""Pins oracle/curiosity_oracle.py against vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import torch

import recipes as R
from helpers import golden, rel_err, check_grads
from oracle import curiosity_oracle as O

TOL = 2e-5


def _leaf(params):
    return {k: v.clone().requires_grad_(True) for k, v in params.items()}


def test_returns():
    for tag, done_fn in [("a", R.make_random_done), ("b", R.make_suffix_done), ("c", R.make_random_done)]:
        g = golden(f"returns_{tag}")
        T, B, seed = int(g["T"]), int(g["B"]), int(g["seed"])
        reward, value = R.make_normal(seed, T, B), R.make_normal(seed + 100, T, B)
        boot, done = R.make_normal(seed + 200, B), done_fn(seed + 300, T, B)
        adv, ret = O.generalized_advantage_estimation(reward, value, done, boot, 0.99, 0.95)
        assert rel_err(adv, g["advantage"]) < 1e-6 and rel_err(ret, g["return_"]) < 1e-6
        assert rel_err(O.discount_return(reward, done, boot, 0.99), g["discount_return"]) < 1e-6
        assert np.array_equal(O.valid_from_done(done).numpy(), g["valid"])


def test_stats():
    g = golden("stats")
    seed = int(g["seed"])
    rms = O.RunningMeanStd(shape=(3,))
    for it in range(4):
        rms.update(R.make_normal(seed + it, 10 + it, 3).numpy().astype(np.float64) * (it + 1))
        assert np.allclose(rms.mean, g[f"mean{it}"], rtol=1e-12)
        assert np.allclose(rms.var, g[f"var{it}"], rtol=1e-12)
        assert np.isclose(rms.count, g[f"count{it}"])
    rff = O.RewardForwardFilter(0.9)
    for it in range(5):
        r = R.make_normal(seed + 50 + it, 7).numpy()
        mask = (R.make_normal(seed + 60 + it, 7).numpy() > -0.5).astype(np.float32)
        assert np.allclose(rff.update(r, mask), g[f"rff{it}"], rtol=1e-6)


def test_process_returns():
    for tag, nr, na, lam in [("plain", False, False, 0.95), ("norm", True, True, 0.95), ("lam1", False, True, 1.0)]:
        g = golden(f"process_returns_{tag}")
        T, B, seed = int(g["T"]), int(g["B"]), int(g["seed"])
        pr = O.ReturnsOracle(0.99, lam, nr, na)
        for it in range(2):
            reward = R.make_normal(seed + 1000 * it, T, B) * 0.0
            done = R.make_suffix_done(seed + 1000 * it + 1, T, B) > 0.5
            value = R.make_normal(seed + 1000 * it + 2, T, B)
            boot = R.make_normal(seed + 1000 * it + 3, 1, B)
            r_int = R.make_normal(seed + 7 * (it + 1), T, B).abs() * 0.1
            ret, adv, valid = pr(reward, done, value, boot, r_int)
            assert rel_err(ret, g[f"ret{it}"]) < TOL, tag
            assert rel_err(adv, g[f"adv{it}"]) < TOL, tag
            assert np.array_equal(valid.numpy(), g[f"valid{it}"])
            assert rel_err(reward, g[f"reward_after{it}"]) < 1e-6


def test_rnd():
    g = golden("rnd")
    T, B, seed = int(g["T"]), int(g["B"]), int(g["seed"])
    params = R.make_params(R.rnd_shapes(), seed, gain=1.4)
    m = O.RndOracle(prediction_beta=1.0, drop_probability=0.25, gamma=0.99)
    for it in range(3):
        frames = R.make_frames(seed + 10 * it, T, B)
        done = R.make_suffix_done(seed + 10 * it + 1, T, B, frac=0.6)
        bonus = m.compute_bonus(params, frames, done)
        assert rel_err(bonus, g[f"bonus{it}"]) < TOL
        assert np.allclose(m.obs_rms.mean, g[f"obs_mean{it}"], rtol=1e-10)
        assert np.allclose(m.obs_rms.var, g[f"obs_var{it}"], rtol=1e-10)
        assert np.isclose(m.rew_rms.var, g[f"rew_var{it}"], rtol=1e-5)
        assert np.isclose(m.rew_rms.count, g[f"rew_count{it}"])
        assert np.allclose(m.rew_rff.rewems, g[f"rewems{it}"], rtol=1e-5)
    p = _leaf(params)
    frames = R.make_frames(seed + 20, T, B)
    valid = 1 - R.make_suffix_done(seed + 22, T, B, frac=0.5)
    mask = R.make_mask(seed + 23, (T, B), 0.25)
    loss = m.compute_loss(p, frames, valid, mask)
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) / float(g["loss"]) < TOL
    grads = {k: v.grad for k, v in p.items() if v.grad is not None}
    assert not check_grads(grads, g, 0.9999, 1e-3)
    assert all(v.grad is None or float(v.grad.abs().sum()) == 0 for k, v in p.items() if k.startswith("target"))


def _icm_case(enc, A):
    g = golden(f"icm_{enc}")
    T, B, seed = int(g["T"]), int(g["B"]), int(g["seed"])
    params = R.make_params(R.icm_shapes(A, enc), seed)
    obs, nxt = R.make_frames(seed + 1, T, B), R.make_frames(seed + 2, T, B)
    act = R.onehot(R.make_actions(seed + 3, T, B, A), A)
    valid = 1 - R.make_suffix_done(seed + 4, T, B, frac=0.5)
    assert rel_err(O.icm_bonus(params, obs, nxt, act, enc, 0.7), g["bonus"]) < TOL
    p = _leaf(params)
    inv, fwd = O.icm_loss(p, obs, nxt, act, valid, enc, 0.2)
    (inv + fwd).backward()
    assert abs(float(inv) - float(g["inv_loss"])) / float(g["inv_loss"]) < TOL
    assert abs(float(fwd) - float(g["fwd_loss"])) / float(g["fwd_loss"]) < TOL
    assert not check_grads({k: v.grad for k, v in p.items()}, g, 0.9999, 1e-3)


def test_icm_burda():
    _icm_case("idf_burda", 4)


def test_icm_universe():
    _icm_case("idf", 7)


def test_disagreement():
    g = golden("disagreement")
    T, B, A, seed = int(g["T"]), int(g["B"]), int(g["A"]), int(g["seed"])
    enc = "idf_burda"
    params = R.make_params(R.disagreement_shapes(A, enc), seed)
    obs, nxt = R.make_frames(seed + 1, T, B), R.make_frames(seed + 2, T, B)
    act = R.onehot(R.make_actions(seed + 3, T, B, A), A)
    valid = 1 - R.make_suffix_done(seed + 4, T, B, frac=0.5)
    assert rel_err(O.disagreement_bonus(params, obs, nxt, act, enc, 1.0), g["bonus"]) < 1e-4
    masks = [R.make_mask(seed + 10 + e, (T, B, 512), 0.2) for e in range(4)]
    p = _leaf(params)
    inv, fwd = O.disagreement_loss(p, obs, nxt, act, valid, enc, 0.2, masks)
    (inv + fwd).backward()
    assert abs(float(inv) - float(g["inv_loss"])) / float(g["inv_loss"]) < TOL
    assert abs(float(fwd) - float(g["fwd_loss"])) / float(g["fwd_loss"]) < TOL
    assert not check_grads({k: v.grad for k, v in p.items()}, g, 0.9999, 1e-3)
    inv2, fwd2 = O.disagreement_loss(params, obs, nxt, act, valid, enc, 0.2, None)
    assert abs(float(fwd2) - float(g["fwd_loss_nodrop"])) / float(g["fwd_loss_nodrop"]) < TOL


def test_ndigo():
    for H in (1, 3):
        g = golden(f"ndigo_h{H}")
        T, B, A, seed = int(g["T"]), int(g["B"]), int(g["A"]), int(g["seed"])
        params = R.make_params(R.ndigo_shapes(A), seed)
        obs = R.make_binary_frames(seed + 1, T, B)
        prev = R.onehot(R.make_actions(seed + 2, T, B, A), A)
        act = R.onehot(R.make_actions(seed + 3, T, B, A), A)
        bonus = O.ndigo_bonus(params, obs, prev, act, H)
        assert float((bonus - torch.from_numpy(g["bonus"])).abs().max()) < 1e-5
        p = _leaf(params)
        loss = O.ndigo_loss(p, obs, prev, act)
        loss.backward()
        assert abs(float(loss) - float(g["loss"])) / float(g["loss"]) < TOL
        assert not check_grads({k: v.grad for k, v in p.items() if v.grad is not None}, g, 0.9999, 1e-3)


def policy_case(g):
    """Inputs of tests/golden/make_golden.py::gen_policy rebuilt from the seed."""
    T, B, A, seed = int(g["T"]), int(g["B"]), int(g["A"]), int(g["seed"])
    return dict(T=T, B=B, A=A, params=R.make_policy_params(A, seed),
                icm_params=R.make_params(R.icm_shapes(A, "idf_burda"), seed + 2),
                obs=R.make_frames(seed + 3, T, B), nxt=R.make_frames(seed + 4, T, B),
                prev_action=R.make_actions(seed + 5, T, B, A), action=R.make_actions(seed + 6, T, B, A),
                old_prob=R.make_probs(seed + 7, T, B, A), ret=R.make_normal(seed + 8, T, B),
                adv=R.make_normal(seed + 9, T, B), valid=1 - R.make_suffix_done(seed + 10, T, B, frac=0.5),
                h0=(R.make_normal(seed + 11, B, 1, 512) * 0.5).transpose(0, 1).contiguous(),
                c0=(R.make_normal(seed + 12, B, 1, 512) * 0.5).transpose(0, 1).contiguous())


def test_policy_ppo_icm():
    """AtariLstmModel.forward + PPO.loss with the ICM terms, through the reference's own agent."""
    g = golden("policy_ppo_icm")
    c = policy_case(g)
    A = c["A"]
    with torch.no_grad():
        pi, v, (hn, cn) = O.policy_forward(c["params"], c["obs"], R.onehot(c["prev_action"], A), (c["h0"], c["c0"]))
    assert rel_err(pi, g["pi"]) < TOL and rel_err(v, g["value"]) < TOL
    assert rel_err(hn, g["hn"]) < TOL and rel_err(cn, g["cn"]) < TOL
    p = _leaf(c["params"])
    q = _leaf(c["icm_params"])
    pi, v, _ = O.policy_forward(p, c["obs"], R.onehot(c["prev_action"], A), (c["h0"], c["c0"]))
    loss, pi_loss, value_loss, entropy_loss, entropy, perplexity = O.ppo_loss(
        pi, v, c["action"], c["old_prob"], c["ret"], c["adv"], c["valid"], float(g["ratio_clip"]),
        float(g["value_loss_coeff"]), float(g["entropy_loss_coeff"]))
    inv, fwd = O.icm_loss(q, c["obs"], c["nxt"], R.onehot(c["action"], A), c["valid"], "idf_burda", 0.2)
    total = loss + inv + fwd
    for name, val in [("pi_loss", pi_loss), ("value_loss", value_loss), ("entropy_loss", entropy_loss),
                      ("entropy", entropy), ("perplexity", perplexity), ("inv_loss", inv), ("fwd_loss", fwd),
                      ("loss", total)]:
        assert abs(float(val) - float(g[name])) <= TOL * max(abs(float(g[name])), 1e-3), name
    total.backward()
    grads = {k: t.grad for k, t in p.items()}
    grads.update({"curiosity_model." + k: t.grad for k, t in q.items()})
    assert not check_grads(grads, g, 0.9999, 1e-3)
    gn = torch.sqrt(sum((t.grad.double() ** 2).sum() for t in list(p.values()) + list(q.values())))
    assert abs(float(gn) - float(g["grad_norm"])) / float(g["grad_norm"]) < 1e-3
