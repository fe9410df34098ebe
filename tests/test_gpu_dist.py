"""Data-parallel parity ON THE GPU KERNELS: two ranks (processes) shard the env columns of one batch and run the
fused step drivers; parameters, running statistics and the logged loss after two steps must equal the single-process
run on the whole batch.  The ranks share the one GPU of the test box, so the process group is gloo over CUDA tensors
(NCCL refuses two ranks per device); the code under test -- global loss normaliser, integer statistics all-reduce,
bucketed gradient all-reduce overlapped with the backward pass, FlatAdam -- is the code the NCCL runs use."""
import os
import socket

import pytest
import torch
import torch.distributed as tdist
import torch.multiprocessing as mp

import recipes as R

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _batch(kind, seed, T, B, A):
    d = dict(done=R.make_suffix_done(seed + 10, T, B, frac=0.4), value=R.make_normal(seed + 20, T, B),
             boot=R.make_normal(seed + 30, B))
    if kind == "rnd":
        d["frames"] = R.make_frames(seed, T, B)
    else:
        d["obs"], d["nxt"] = R.make_frames(seed, T, B), R.make_frames(seed + 1, T, B)
        d["act"] = R.onehot(R.make_actions(seed + 2, T, B, A), A)
    return d


def _run(kind, T, B, A, lo, hi, iters=2):
    """Build the model from seeded weights and run ``iters`` fused steps on env columns [lo, hi)."""
    from curiosity_baselines_b200.curiosity.icm import ICM
    from curiosity_baselines_b200.curiosity.rnd import RND
    from curiosity_baselines_b200.trainer import RndStep, IcmStep
    seed = 31
    d = lambda t: t[:, lo:hi].contiguous().to(DEV)
    if kind == "rnd":
        m = RND(image_shape=(4, 84, 84), drop_probability=0.0).to(DEV)
        m.load_state_dict(R.make_params(R.rnd_shapes(), seed, gain=1.4))
        step = RndStep(m, lr=1e-3)
    else:
        m = ICM(image_shape=(4, 84, 84), action_size=A, feature_encoding="idf_burda").to(DEV)
        m.load_state_dict(R.make_params(R.icm_shapes(A, "idf_burda"), seed))
        step = IcmStep(m, lr=1e-3)
    step.opt.reducer.bucket = 64 * 1024          # small buckets: several overlapped reductions even for this model
    # fp32 atomics sum the weight gradients in a different order from run to run; with Adam's default eps = 1e-8 an
    # element whose gradient is ~0 turns that rounding into a full +-lr step.  eps = 1e-4 keeps the comparison
    # well conditioned (the gradients that matter are 1e-2 .. 1e2 with raw 0..255 pixels).
    step.opt.eps = 1e-4
    out = {}
    for it in range(iters):
        b = _batch(kind, seed + 100 * it, T, B, A)
        rew = torch.zeros(T, hi - lo, device=DEV)
        if kind == "rnd":
            loss, r_int, adv, ret = step(d(b["frames"]), d(b["done"]), rew, d(b["value"]), b["boot"][lo:hi].to(DEV),
                                         torch.ones(T, hi - lo, device=DEV))
        else:
            loss, r_int, adv, ret = step(d(b["obs"]), d(b["nxt"]), d(b["act"]), d(b["done"]), rew, d(b["value"]),
                                         b["boot"][lo:hi].to(DEV))
        out[f"loss{it}"] = float(loss)
        out[f"r_int{it}"] = r_int.cpu()
        out[f"gnorm{it}"] = step.opt.grad_norm()
    out["flat"] = step.opt.flat.cpu()
    out["launched"] = step.opt.reducer.launched
    if kind == "rnd":
        out.update(obs_mean=torch.as_tensor(m.obs_rms.mean), obs_var=torch.as_tensor(m.obs_rms.var),
                   obs_count=m.obs_rms.count, rew_var=float(m.rew_rms.var), rew_count=m.rew_rms.count)
    return out


def _worker(rank, world, port, kind, T, B, A, path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    from curiosity_baselines_b200 import dist as D
    lo, hi = D.shard_columns(B)
    out = _run(kind, T, B, A, lo, hi)
    torch.save(out, f"{path}.{rank}")
    tdist.destroy_process_group()


@pytest.mark.parametrize("kind", ["rnd", "icm"])
def test_two_ranks_equal_one_rank(kind, tmp_path):
    T, B, A, world = 8, 8, 4, 2
    path = str(tmp_path / "rank")
    mp.spawn(_worker, args=(world, _free_port(), kind, T, B, A, path), nprocs=world, join=True)
    r0, r1 = torch.load(path + ".0"), torch.load(path + ".1")
    one = _run(kind, T, B, A, 0, B)
    assert r0["launched"] >= 3                                   # the gradient was reduced in several buckets
    for it in range(2):
        assert abs(r0[f"loss{it}"] - one[f"loss{it}"]) <= 1e-5 * abs(one[f"loss{it}"])       # all-reduced logged loss
        assert r0[f"loss{it}"] == r1[f"loss{it}"]
        assert abs(r0[f"gnorm{it}"] - one[f"gnorm{it}"]) <= 1e-4 * one[f"gnorm{it}"]
        both = torch.cat([r0[f"r_int{it}"], r1[f"r_int{it}"]], 1)
        assert torch.allclose(both, one[f"r_int{it}"], rtol=2e-5, atol=1e-7)
    assert torch.equal(r0["flat"], r1["flat"])                  # the replicas stay bit-identical
    diff = (r0["flat"] - one["flat"]).abs()
    assert float(diff.max()) <= 1e-5, (float(diff.max()), float(diff.mean()))
    if kind == "rnd":
        assert torch.equal(r0["obs_mean"], one["obs_mean"]) and torch.equal(r0["obs_var"], one["obs_var"])   # integer sums
        assert r0["obs_count"] == one["obs_count"] and r0["rew_count"] == one["rew_count"]
        assert abs(r0["rew_var"] - one["rew_var"]) <= 1e-6 * one["rew_var"]
