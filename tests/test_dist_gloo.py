"""Host-side data-parallel logic on CPU with the gloo backend, world_size 2: column sharding and
the exchange steps of SURVEY 8(e) -- exact integer observation sums, {sum, sum^2, n} moment triples
and the global loss normaliser -- reproduce the single-process quantities on every rank."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import recipes as R


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, T, B, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from curiosity_baselines_b200 import dist as D
    assert D.world_size() == world and D.rank() == rank
    lo, hi = D.shard_columns(B)
    frames = R.make_frames(3, T, B)[:, lo:hi, -1].reshape(-1, 84 * 84).to(torch.int64)
    done = R.make_suffix_done(4, T, B, frac=0.5)[:, lo:hi]
    valid_rows = (torch.arange(T).unsqueeze(1) < (1 - done).sum(0).unsqueeze(0)).reshape(-1)
    # exact integer sums of the valid frames (what cb_obs_moments_u8 produces per rank)
    sums = torch.stack([(frames * valid_rows.unsqueeze(1)).sum(0), (frames * frames * valid_rows.unsqueeze(1)).sum(0)])
    n_total = valid_rows.sum().reshape(1)
    D.all_reduce_sum_(sums)
    D.all_reduce_sum_(n_total)
    # reward moments triple and loss normaliser
    x = R.make_normal(5, T, B)[:, lo:hi].double()
    tri = torch.stack([x.sum(), (x * x).sum(), torch.tensor(float(x.numel()), dtype=torch.float64)])
    D.all_reduce_sum_(tri)
    valid_sum = D.all_reduce_sum_((1 - done).sum().reshape(1))
    if rank == 0:
        torch.save(dict(sums=sums, n_total=n_total, tri=tri, valid_sum=valid_sum), out)
    dist.destroy_process_group()


def test_sharded_statistics_equal_single_process(tmp_path):
    T, B, world = 6, 8, 2
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(world, _free_port(), T, B, out), nprocs=world, join=True)
    got = torch.load(out)
    frames = R.make_frames(3, T, B)[:, :, -1].reshape(-1, 84 * 84).to(torch.int64)
    done = R.make_suffix_done(4, T, B, frac=0.5)
    valid_rows = (torch.arange(T).unsqueeze(1) < (1 - done).sum(0).unsqueeze(0)).reshape(-1)
    ref = torch.stack([(frames * valid_rows.unsqueeze(1)).sum(0), (frames * frames * valid_rows.unsqueeze(1)).sum(0)])
    assert torch.equal(got["sums"], ref)                         # integers: bit-exact across shardings
    assert int(got["n_total"]) == int(valid_rows.sum())
    x = R.make_normal(5, T, B).double()
    assert np.isclose(float(got["tri"][0]), float(x.sum())) and np.isclose(float(got["tri"][1]), float((x * x).sum()))
    assert float(got["tri"][2]) == x.numel()
    assert float(got["valid_sum"]) == float((1 - done).sum())
    # the merged mean / variance equal numpy's over the unsharded batch
    n = float(got["tri"][2])
    mean, var = got["tri"][0] / n, got["tri"][1] / n - (got["tri"][0] / n) ** 2
    assert np.isclose(float(mean), x.mean().item()) and np.isclose(float(var), x.var(unbiased=False).item())


def test_shard_columns():
    from curiosity_baselines_b200.dist import shard_columns
    assert [shard_columns(1024, r, 8) for r in (0, 7)] == [(0, 128), (896, 1024)]
    try:
        shard_columns(10, 0, 4)
        raise AssertionError("expected ValueError")
    except ValueError:
        pass


def _reducer_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from curiosity_baselines_b200 import dist as D
    flat = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    red = D.BucketReducer(flat, bucket_bytes=4 * 100)
    views = [flat[900:1000], flat[700:900], flat[650:700], flat[0:30]]       # backward order: last layer first
    red.ready(views[:1])
    assert red.launched == 1                                                   # 100 elements reach the bucket size
    red.ready(views[2:3])
    assert red.launched == 1 and red.pending == [(650, 700)]                 # 50 elements wait for neighbours
    red.ready(views[1:2])
    assert red.launched == 2 and red.reduced == [(650, 1000)]                # coalesced with the adjacent range
    red.ready(views[3:])
    try:
        red.ready([flat[10:20]])
        raise AssertionError("a range reported twice must raise")
    except RuntimeError:
        pass
    red.finish()                                                               # the rest, reported or not
    assert red.pending == [] and red.reduced == []
    if rank == 0:
        torch.save(flat, out)
    dist.destroy_process_group()


def test_bucket_reducer_sums_every_element_once(tmp_path):
    out = str(tmp_path / "flat.pt")
    mp.spawn(_reducer_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert torch.equal(torch.load(out), torch.arange(1000, dtype=torch.float32) * 3)
