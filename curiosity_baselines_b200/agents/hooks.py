"""Agent hooks: drop-in for ``RecurrentCategoricalPgAgentBase.curiosity_step`` /
``curiosity_loss`` (rlpyt/agents/pg/categorical.py:83-130).  They can be bound onto the
reference agent class; they need ``self.model.curiosity_model``, ``self.device`` and the
number of actions (from ``self.distribution.dim`` or the model).

Host-to-device staging differs from the reference on purpose: uint8 frame stacks are copied
once per call with ``non_blocking`` from (ideally pinned) sampler buffers, and RND uploads only
the newest frame of each stack (7056 of 28224 bytes per sample) since nothing else is read
(rnd.py:130-131).
"""
from collections import namedtuple

import torch

AgentCuriosityStep = namedtuple("AgentCuriosityStep", ["r_int", "agent_curiosity_info"])
IcmInfo = namedtuple("IcmInfo", [])
NdigoInfo = namedtuple("NdigoInfo", ["prev_gru_state"])
RndInfo = namedtuple("RndInfo", [])


def _device(agent):
    dev = torch.device(getattr(agent, "device", "cuda"))
    if dev.type != "cuda":
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def _curiosity_model(agent):
    """``agent.model.curiosity_model``, also when rlpyt's ``data_parallel()`` wrapped the model in
    DistributedDataParallel (runners/sync_rl.py; the reference's hooks then fail on the missing attribute,
    agents/pg/categorical.py:102 -- SURVEY 8f rank 4)."""
    m = agent.model
    return getattr(m, "module", m).curiosity_model


def _onehot(agent, idx, dev):
    n = agent.distribution.dim if hasattr(agent, "distribution") else _curiosity_model(agent).action_size
    idx = idx.to(dev).long()
    return torch.zeros(idx.shape + (n,), dtype=torch.float32, device=dev).scatter_(-1, idx.unsqueeze(-1), 1)


def newest_frame_to_device(frames, dev):
    """[T,B,C,H,W] uint8 (host or device) -> device [T,B,1,H,W] holding only the newest frame."""
    if frames.is_cuda:
        return frames
    from .. import _lib as L
    T, B, C, H, W = frames.shape
    if frames.dtype != torch.uint8 or not frames.is_contiguous():
        return frames[:, :, -1:].contiguous().to(dev, non_blocking=True)
    out = torch.empty((T, B, 1, H, W), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        L.call("cb_h2d_2d", out.data_ptr(), H * W, frames.data_ptr() + (C - 1) * H * W, C * H * W,
               H * W, T * B, L.stream())
    return out


class FrameStager:
    """Double-buffered host->device staging of the newest frame of pinned uint8 stacks on a side
    stream.  Per batch:  ``buf = get()`` (compute stream waits for the upload),
    queue the compute that reads ``buf``, ``done()`` (marks the point after which ``buf`` may be
    overwritten), ``prefetch(next_frames)``: the upload of batch i+1 runs under the compute of batch i and
    only waits for batch i-1 to release its buffer (an event, not a whole-stream dependency)."""

    def __init__(self, device):
        self.dev = torch.device(device)
        if self.dev.index is None:
            self.dev = torch.device("cuda", torch.cuda.current_device())
        self.stream = torch.cuda.Stream(device=self.dev)
        self.bufs, self.events, self.free, self.k = [None, None], [None, None], [None, None], 0
        self.extras = [None, None]
        self.pending = self.current = None

    def prefetch(self, frames, extras=()):
        """``extras``: the batch's small pinned host tensors (done, value, ...).  They ride on the same
        stream AHEAD of the frames: separate small uploads on the compute stream would queue behind the
        next batch's 0.9 GB frame copy in the copy engine and stall the step (measured: 27 -> 40 ms)."""
        from .. import _lib as L
        T, B, C, H, W = frames.shape
        i = self.k
        if self.bufs[i] is None or self.bufs[i].shape != (T, B, 1, H, W):
            self.bufs[i] = torch.empty((T, B, 1, H, W), dtype=torch.uint8, device=self.dev)
            self.stream.wait_stream(torch.cuda.current_stream(self.dev))
        if self.free[i] is not None:
            self.stream.wait_event(self.free[i])       # the last reader of buffer i has finished
        with torch.cuda.stream(self.stream):
            if self.extras[i] is None or len(self.extras[i]) != len(extras):
                self.extras[i] = [torch.empty(t.shape, dtype=t.dtype, device=self.dev) for t in extras]
            for d, t in zip(self.extras[i], extras):
                d.copy_(t, non_blocking=True)
            L.call("cb_h2d_2d", self.bufs[i].data_ptr(), H * W, frames.data_ptr() + (C - 1) * H * W, C * H * W, H * W,
                   T * B, self.stream.cuda_stream)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.events[i] = ev
        self.pending = i
        self.k ^= 1

    def get(self):
        i = self.pending
        torch.cuda.current_stream(self.dev).wait_event(self.events[i])
        self.current = i
        return self.bufs[i]

    def get_extras(self):
        """Device copies of the ``extras`` of the batch returned by the last ``get()``."""
        return self.extras[self.current]

    def done(self):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.dev))
        self.free[self.current] = ev


class BatchPrefetcher:
    """Whole-batch host->device staging on a side stream for the full PPO iteration (ICM / Disagreement read all
    four frames of observation AND next_observation, 2 x T x B x 28 224 bytes): ``prefetch(host_tensors)`` queues
    the upload of the NEXT batch from pinned memory while the current one is being consumed; ``get()`` makes the
    compute stream wait for it and hands the device tensors over.  Two batches are resident at a time."""

    def __init__(self, device):
        self.dev = torch.device(device)
        if self.dev.index is None:
            self.dev = torch.device("cuda", torch.cuda.current_device())
        self.stream = torch.cuda.Stream(device=self.dev)
        self.pending = None

    def prefetch(self, host_tensors):
        with torch.cuda.stream(self.stream):
            dev_tensors = [t.to(self.dev, non_blocking=True) for t in host_tensors]
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.pending = (dev_tensors, ev)

    def get(self):
        dev_tensors, ev = self.pending
        self.pending = None
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_event(ev)
        for t in dev_tensors:
            t.record_stream(cur)          # allocated on the side stream, consumed on the compute stream
        return dev_tensors


@torch.no_grad()
def curiosity_step(self, curiosity_type, *args):
    dev = _device(self)
    cm = _curiosity_model(self)
    if curiosity_type in ("icm", "disagreement"):
        observation, next_observation, actions = args
        r_int = cm.compute_bonus(observation.to(dev, non_blocking=True),
                                 next_observation.to(dev, non_blocking=True), _onehot(self, actions, dev))
        info = IcmInfo()
    elif curiosity_type == "ndigo":
        observation, prev_actions, actions = args
        r_int = cm.compute_bonus(observation.to(dev, non_blocking=True), _onehot(self, prev_actions, dev),
                                 _onehot(self, actions, dev))
        info = NdigoInfo(prev_gru_state=None)
    elif curiosity_type == "rnd":
        next_observation, done = args
        r_int = cm.compute_bonus(newest_frame_to_device(next_observation, dev), done.to(dev))
        info = RndInfo()
    else:
        raise ValueError(curiosity_type)
    return AgentCuriosityStep(r_int=r_int.to("cpu"), agent_curiosity_info=info)


def curiosity_loss(self, curiosity_type, *args):
    dev = _device(self)
    cm = _curiosity_model(self)
    if curiosity_type in ("icm", "disagreement"):
        observation, next_observation, actions, valid = args
        inv, fwd = cm.compute_loss(observation.to(dev, non_blocking=True),
                                   next_observation.to(dev, non_blocking=True),
                                   _onehot(self, actions, dev), valid.to(dev))
        return inv.to("cpu"), fwd.to("cpu")
    if curiosity_type == "ndigo":
        observations, prev_actions, actions, valid = args
        loss = cm.compute_loss(observations.to(dev, non_blocking=True), _onehot(self, prev_actions, dev),
                               _onehot(self, actions, dev), valid.to(dev))
        return loss.to("cpu")
    if curiosity_type == "rnd":
        next_observation, valid = args
        loss = cm.compute_loss(newest_frame_to_device(next_observation, dev), valid.to(dev))
        return loss.to("cpu")
    raise ValueError(curiosity_type)
