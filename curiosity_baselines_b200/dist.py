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


def bind_to_gpu_numa_node(device_index):
    """Pin this process (and therefore the pinned host buffers it allocates afterwards: first touch) to the CPUs of
    the NUMA node the GPU hangs off.  With one process per GPU the newest-frame uploads of 8 ranks (8 x 0.93 GB per
    step) otherwise cross the inter-socket link for half the ranks.  Returns the node, or None when the platform
    does not expose the topology (then nothing is changed)."""
    import os
    try:
        prop = torch.cuda.get_device_properties(device_index)
        bus = f"{prop.pci_domain_id:04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except (OSError, AttributeError, ValueError):
        return None
