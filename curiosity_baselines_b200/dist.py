"""Data parallelism over the B (environment) dimension: one process per GPU.

The path shards naturally over B: every env column is independent in the models,
the reward filter and GAE.  The exchange steps are few and small, all through
``torch.distributed`` (NCCL over NVLink on GPUs, gloo in the CPU tests):
  * predictor / encoder gradients: one all-reduce(SUM) of the flat gradient buffer per step;
  * loss normaliser sum(valid): one scalar all-reduce so every rank divides by the global count;
  * RND observation statistics: exact integer per-pixel sums (+ frame count) all-reduced, so
    each rank holds the single-process running state bit-for-bit;
  * RND reward statistics / reward and advantage normalisation: {sum, sum^2, n} triples.
"""
import torch
import torch.distributed as dist


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def all_reduce_sum_(t):
    """In-place SUM all-reduce (no-op for a single process).  int64 / fp64 / fp32 tensors."""
    if world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def shard_columns(B, r=None, w=None):
    """Contiguous block of env columns owned by rank r: [lo, hi)."""
    r = rank() if r is None else r
    w = world_size() if w is None else w
    per = B // w
    if per * w != B:
        raise ValueError(f"B={B} is not divisible by world size {w}")
    return r * per, (r + 1) * per
