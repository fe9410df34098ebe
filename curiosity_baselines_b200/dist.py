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


class BucketReducer:
    """Gradient all-reduce overlapped with the backward pass (ppo.py:147-150 under data
    parallelism).  ``flat`` is the optimiser's flat fp32 gradient buffer.  As weight-gradient kernels finish being
    enqueued the model reports their tensors with ``ready``; adjacent reported ranges coalesce and every range that
    reaches ``bucket_bytes`` is all-reduced asynchronously (NCCL runs it on its own stream behind the work already
    queued on the compute stream, so it overlaps the remaining dgrad / wgrad launches).  ``finish`` reduces whatever
    has not been reduced yet -- including parameters nobody reported -- and makes the compute stream wait.  Every
    element is reduced exactly once per step; a range reported twice raises."""

    def __init__(self, flat, bucket_bytes=2 << 20):
        self.flat = flat
        self.bucket = max(1, bucket_bytes // flat.element_size())
        self.pending, self.reduced, self.works = [], [], []
        self.launched = 0

    def _insert(self, ivals, lo, hi):
        out, placed = [], False
        for a, b in ivals:
            if b < lo or a > hi:
                out.append((a, b))
            else:
                lo, hi = min(a, lo), max(b, hi)
        out.append((lo, hi))
        out.sort()
        return out

    def _launch(self, lo, hi):
        self.works.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, async_op=True))
        self.reduced = self._insert(self.reduced, lo, hi)
        self.launched += 1

    def ready(self, tensors):
        if world_size() == 1:
            return
        base, es, n = self.flat.data_ptr(), self.flat.element_size(), self.flat.numel()
        for t in tensors:
            if t is None:
                continue
            off = (t.data_ptr() - base) // es
            if off < 0 or off >= n or not t.is_contiguous():
                continue                                   # not a view of this buffer
            lo, hi = off, off + t.numel()
            for a, b in self.reduced + self.pending:
                if lo < b and a < hi:
                    raise RuntimeError("BucketReducer: gradient range reported twice in one step")
            self.pending = self._insert(self.pending, lo, hi)
        keep = []
        for a, b in self.pending:
            if b - a >= self.bucket:
                self._launch(a, b)
            else:
                keep.append((a, b))
        self.pending = keep

    def finish(self):
        """All-reduce every element not reduced yet, wait for all launched reductions, reset for the next step."""
        if world_size() > 1:
            pos, n = 0, self.flat.numel()
            for a, b in list(self.reduced) + [(n, n)]:
                if a > pos:
                    self._launch(pos, a)
                pos = max(pos, b)
            for w in self.works:
                w.wait()
        self.pending, self.reduced, self.works = [], [], []


_reducer = None


def set_grad_reducer(r):
    """The optimiser that owns the flat gradient buffer registers its reducer here; the layers report finished
    weight gradients through ``grads_ready`` without knowing about the optimiser."""
    global _reducer
    _reducer = r


def grads_ready(tensors):
    if _reducer is not None:
        _reducer.ready(tensors)


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
