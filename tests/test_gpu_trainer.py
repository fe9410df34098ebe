"""The fused step driver (bonus -> returns -> loss/grad -> clip -> Adam) vs the same sequence
on the CPU oracle with torch.optim.Adam + clip_grad_norm_ (ppo.py:147-150)."""
import pytest
import torch

import recipes as R
from oracle import curiosity_oracle as O
from curiosity_baselines_b200.curiosity.rnd import RND
from curiosity_baselines_b200.trainer import RndStep, FlatAdam
from curiosity_baselines_b200 import _lib as L

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_flat_adam_matches_torch():
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(s, device=DEV)) for s in [(5, 7), (33,), (4, 3, 2, 2)]]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    opt_ref = torch.optim.Adam(ref, lr=1e-2)
    opt = FlatAdam(ps, lr=1e-2, max_norm=0.5)
    for it in range(4):
        for p, q in zip(ps, ref):
            g = torch.randn_like(q) * (it + 1)
            p.grad.copy_(g)
            q.grad = g.clone()
        torch.nn.utils.clip_grad_norm_(ref, 0.5)
        opt_ref.step()
        opt.step()
        for p, q in zip(ps, ref):
            assert torch.allclose(p, q, rtol=1e-5, atol=1e-6), it


def test_flat_adam_state_dict_interchanges_with_torch_adam():
    """the optimiser state must survive a checkpoint.  torch.optim.Adam's state_dict loads into FlatAdam
    (and back) and both continue identically -- Adam moments, step count (bias correction) and learning rate."""
    torch.manual_seed(1)
    shapes = [(6, 5), (17,), (3, 2, 2, 2)]
    ref = [torch.nn.Parameter(torch.randn(s, device=DEV)) for s in shapes]
    opt_ref = torch.optim.Adam(ref, lr=3e-3)
    grads = [[torch.randn(s, device=DEV) for s in shapes] for _ in range(5)]
    for it in range(2):
        for q, g in zip(ref, grads[it]):
            q.grad = g.clone()
        opt_ref.step()
    ps = [torch.nn.Parameter(q.detach().clone()) for q in ref]
    opt = FlatAdam(ps, lr=1.0, max_norm=1e9)
    opt.load_state_dict(opt_ref.state_dict())
    assert opt.step_count == 2 and opt.lr == 3e-3
    for it in range(2, 4):
        for p, q, g in zip(ps, ref, grads[it]):
            p.grad.copy_(g)
            q.grad = g.clone()
        opt_ref.step()
        opt.step()
    for p, q in zip(ps, ref):
        assert torch.allclose(p, q, rtol=1e-5, atol=1e-6)
    # ... and back: a fresh torch Adam resumes from FlatAdam's state
    ref2 = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    opt2 = torch.optim.Adam(ref2, lr=3e-3)
    opt2.load_state_dict(opt.state_dict())
    for p, q, g in zip(ps, ref2, grads[4]):
        p.grad.copy_(g)
        q.grad = g.clone()
    opt.step()
    opt2.step()
    for p, q in zip(ps, ref2):
        assert torch.allclose(p, q, rtol=1e-5, atol=1e-6)
    sd = opt.state_dict()
    assert set(sd) == {"state", "param_groups"} and sd["param_groups"][0]["params"] == [0, 1, 2]
    assert float(sd["state"][0]["step"]) == 5.0


def test_rnd_step_against_oracle_two_iterations():
    T, B, seed = 12, 6, 77
    params = R.make_params(R.rnd_shapes(), seed, gain=1.4)
    model = RND(image_shape=(4, 84, 84), drop_probability=0.0).to(DEV)
    model.load_state_dict(params)
    step = RndStep(model, lr=1e-3)
    o = O.RndOracle(drop_probability=0.0)
    p = {k: v.clone().requires_grad_(k.startswith("forward")) for k, v in params.items()}
    train = [v for k, v in p.items() if k.startswith("forward")]
    opt = torch.optim.Adam(train, lr=1e-3)
    for it in range(2):
        frames = R.make_frames(seed + it, T, B)
        done = R.make_suffix_done(seed + 10 + it, T, B, frac=0.4)
        value, boot = R.make_normal(seed + 20 + it, T, B), R.make_normal(seed + 30 + it, B)
        loss, r_int, adv, ret = step(frames.to(DEV), done.to(DEV), torch.zeros(T, B, device=DEV),
                                     value.to(DEV), boot.to(DEV), torch.ones(T, B, device=DEV))
        r_ref = o.compute_bonus(p, frames, done)
        a_ref, ret_ref = O.generalized_advantage_estimation(r_ref.clone(), value, done, boot, 0.99, 0.95)
        valid = O.valid_from_done(done)
        opt.zero_grad()
        l_ref = o.compute_loss(p, frames, valid, torch.ones(T, B))
        l_ref.backward()
        gn_ref = float(torch.nn.utils.clip_grad_norm_(train, 1.0))
        opt.step()
        assert abs(float(loss) - float(l_ref)) / float(l_ref) < 1e-2, it
        assert abs(step.opt.grad_norm() - gn_ref) / gn_ref < 1e-2, it
        assert float((adv.cpu() - a_ref).abs().max() / a_ref.abs().max()) < 1e-2
        assert float((ret.cpu() - ret_ref).abs().max() / ret_ref.abs().max()) < 1e-2
    # parameter updates point the same way
    for k, prm in model.named_parameters():
        if not k.startswith("forward"):
            assert torch.equal(prm.cpu(), params[k])          # target stays frozen
            continue
        d = (prm.detach().cpu() - params[k]).flatten()
        d_ref = (p[k].detach() - params[k]).flatten()
        cos = float(d @ d_ref / (d.norm() * d_ref.norm()))
        assert cos > 0.97, (k, cos)


def test_host_staging_of_newest_frame():
    from curiosity_baselines_b200.agents.hooks import BatchStaging
    frames = R.make_frames(3, 5, 4).pin_memory()
    out = BatchStaging(DEV).get(frames, newest_only=True)
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), frames[:, :, -1:])


def test_disagreement_step_against_oracle_two_iterations():
    """IcmStep (bonus -> returns -> loss/grad -> clip -> Adam) on Disagreement, grouped ensemble GEMMs included,
    against the CPU oracle trained with torch.optim.Adam for two iterations: the second iteration only matches if
    the stacked / prepared bf16 weights were refreshed after the first update."""
    from curiosity_baselines_b200.curiosity.icm import Disagreement
    from curiosity_baselines_b200.trainer import IcmStep
    T, B, A, seed = 10, 6, 7, 91
    params = R.make_params(R.disagreement_shapes(A, "idf_burda"), seed)
    model = Disagreement(image_shape=(4, 84, 84), action_size=A, feature_encoding="idf_burda").to(DEV)
    model.load_state_dict(params)
    step = IcmStep(model, lr=1e-3)
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    opt = torch.optim.Adam(list(p.values()), lr=1e-3)
    losses = []
    for it in range(2):
        obs, nxt = R.make_frames(seed + it, T, B), R.make_frames(seed + 5 + it, T, B)
        act = R.onehot(R.make_actions(seed + 10 + it, T, B, A), A)
        done = R.make_suffix_done(seed + 15 + it, T, B, frac=0.4)
        value, boot = R.make_normal(seed + 20 + it, T, B), R.make_normal(seed + 30 + it, B)
        loss, r_int, adv, ret = step(obs.to(DEV), nxt.to(DEV), act.to(DEV), done.to(DEV), torch.zeros(T, B, device=DEV),
                                     value.to(DEV), boot.to(DEV), dropout_masks=None)
        r_ref = O.disagreement_bonus(p, obs, nxt, act, "idf_burda", 1.0)
        valid = O.valid_from_done(done)
        opt.zero_grad()
        inv_r, fwd_r = O.disagreement_loss(p, obs, nxt, act, valid, "idf_burda", 0.2, None)
        (inv_r + fwd_r).backward()
        gn_ref = float(torch.nn.utils.clip_grad_norm_(list(p.values()), 1.0))
        opt.step()
        l_ref = float(inv_r + fwd_r)
        losses.append((float(loss), l_ref))
        assert abs(float(loss) - l_ref) / l_ref < 2e-2, (it, losses)
        assert abs(step.opt.grad_norm() - gn_ref) / gn_ref < 5e-2, (it, step.opt.grad_norm(), gn_ref)
    # the loss moved between the iterations (different batch + updated weights) and both sides agree on it
    assert abs(losses[1][0] - losses[0][0]) > 0
    for k, prm in model.named_parameters():
        d = (prm.detach().cpu() - params[k]).flatten()
        d_ref = (p[k].detach() - params[k]).flatten()
        if float(d_ref.norm()) > 0:
            cos = float(d @ d_ref / (d.norm() * d_ref.norm()).clamp_min(1e-30))
            assert cos > 0.9, (k, cos)


def test_graphed_step_replays_a_full_training_step():
    """trainer.GraphedStep: the RND step captured in a CUDA graph and replayed gives the same sequence of losses,
    bonuses and parameters as eager launches -- Adam's step counter / learning rate and all running statistics are
    device-resident, so nothing is frozen at capture time."""
    from curiosity_baselines_b200.trainer import GraphedStep
    T, B, seed = 8, 6, 91
    params = R.make_params(R.rnd_shapes(), seed, gain=1.4)
    frames = [R.make_frames(seed + i, T, B).to(DEV) for i in range(4)]
    done = R.make_suffix_done(seed + 10, T, B, frac=0.4).to(DEV)
    value, boot, ones = R.make_normal(seed + 20, T, B).to(DEV), R.make_normal(seed + 30, B).to(DEV), torch.ones(T, B, device=DEV)

    def build():
        m = RND(image_shape=(4, 84, 84), drop_probability=0.0).to(DEV)
        m.load_state_dict(params)
        return m, RndStep(m, lr=1e-3)

    m0, eager = build()
    ref = []
    for i in range(3 + 3):                                  # GraphedStep runs 3 warm-up steps, then captures (no execution)
        f = frames[0] if i < 3 else frames[i - 2]
        loss, r_int, adv, ret = eager(f, done, torch.zeros(T, B, device=DEV), value, boot, ones)
        ref.append((float(loss), r_int.clone()))
    m1, drv = build()
    static_frames, rew = frames[0].clone(), torch.zeros(T, B, device=DEV)
    g = GraphedStep(drv, (static_frames, done, rew, value, boot, ones))
    assert drv.opt.step_count == 3                         # the warm-up steps really trained; capture itself executes nothing
    for i in range(1, 4):
        rew.zero_()
        loss, r_int, adv, ret = g(frames[i], done, rew, value, boot, ones)
        assert abs(float(loss) - ref[2 + i][0]) <= 1e-4 * abs(ref[2 + i][0]), i
        assert torch.allclose(r_int, ref[2 + i][1], rtol=1e-3, atol=1e-6), i
    assert drv.opt.step_count == 6
    diff = (drv.opt.flat - eager.opt.flat).abs()
    assert float(diff.max()) <= 2e-3 and float(diff.mean()) <= 2e-6         # same training trajectory (atomics reorder sums)
