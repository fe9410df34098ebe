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


class BatchStaging:
    """Host -> device staging of one training iteration's ``Samples`` on the REAL call path: the
    reference hands the same sampler buffers to ``curiosity_step`` (the bonus, algos/pg/base.py:64-73) and to
    ``curiosity_loss`` (every epoch / minibatch, ppo.py:93-110, 226-238).  Each host tensor crosses PCIe ONCE per
    iteration:

      * entries are keyed on (address, shape, dtype, newest-frame-only) and validated with a strided fingerprint of
        the host bytes, so a buffer that was rewritten in place (the sampler reuses its buffers every iteration, and
        shared-memory writes do not bump ``_version``) is uploaded again instead of being served stale;
      * every upload bumps ``generation`` -- the models key their feature caches on it;
      * with ``pin_host`` (off by default; ``integration.install(pin_sampler_buffers=True)`` turns it on for rlpyt,
        whose sample buffers are numpy shared memory allocated once for the whole run, samplers/buffer.py) large
        unpinned buffers are page-locked in place once with ``cudaHostRegister`` so the copies run at PCIe speed,
        asynchronously, on a side stream.  Only for buffers that outlive the staging object: freeing registered
        memory without ``release_host_registrations()`` is an error;
      * ``prefetch`` stages a batch ahead of its iteration (double-buffered runners, rlpyt's AsyncRl): the upload of
        batch i+1 then runs under the compute of batch i.
    RND reads only the newest frame of each stack (rnd.py:130-131): 7056 of 28224 bytes per sample are uploaded with
    one strided DMA (``cb_h2d_2d``)."""

    PIN_MIN_BYTES = 1 << 20
    default_pin_host = False
    default_assume_shifted = False

    def __init__(self, device, pin_host=None, assume_shifted=None):
        self.pin_host = self.default_pin_host if pin_host is None else bool(pin_host)
        # rlpyt's collectors write next_observation[t] and observation[t+1] from the same array
        # (samplers/parallel/cpu/collectors.py:118,156): with ``assume_shifted`` only T+1 steps cross PCIe (see
        # get_shifted).  Opt-in (``integration.install(shifted_next_observation=True)``): it relies on that contract
        # of the sampler, spot-checked on the host.
        self.assume_shifted = self.default_assume_shifted if assume_shifted is None else bool(assume_shifted)
        self.dev = torch.device(device)
        if self.dev.index is None:
            self.dev = torch.device("cuda", torch.cuda.current_device())
        self.stream = torch.cuda.Stream(device=self.dev)
        self.entries = {}
        self.iteration = 0
        self.generation = 0
        self.bytes_uploaded = 0
        self.hits = 0
        self._registered = {}

    # ------------------------------------------------------------------ host side
    @staticmethod
    def _fingerprint(t):
        """64 evenly spaced 8-byte words + size: detects in-place rewrites of a sampler buffer for microseconds."""
        import numpy as np
        flat = t.detach().reshape(-1)
        nbytes = flat.numel() * flat.element_size()
        raw = flat.view(torch.uint8) if flat.dtype != torch.bool else flat.to(torch.uint8)
        words = raw[: nbytes // 8 * 8].view(torch.int64)
        if words.numel() == 0:
            return (nbytes, tuple(raw.tolist()))
        idx = np.linspace(0, words.numel() - 1, num=min(64, words.numel())).astype(np.int64)
        return (nbytes, tuple(words[torch.from_numpy(idx)].tolist()))

    def _pin(self, t):
        """Page-lock a large host buffer in place (once); returns whether async copies from it are possible."""
        if t.is_pinned():
            return True
        nbytes = t.numel() * t.element_size()
        if not self.pin_host or nbytes < self.PIN_MIN_BYTES or not t.is_contiguous():
            return False
        key = (t.data_ptr(), nbytes)
        if key not in self._registered:
            rt = torch.cuda.cudart()
            rc = rt.cudaHostRegister(t.data_ptr(), nbytes, 0)
            self._registered[key] = int(rc) == 0
            if int(rc) != 0:
                rt.cudaGetLastError()         # a refused registration must not poison the next CUDA call
        return self._registered[key]

    def release_host_registrations(self):
        for (ptr, _), ok in self._registered.items():
            if ok:
                torch.cuda.cudart().cudaHostUnregister(ptr)
        self._registered.clear()

    # ------------------------------------------------------------------ uploads
    def new_iteration(self):
        """Called by the first consumer of an iteration (the bonus): drops the previous iteration's device copies;
        batches staged ahead with ``prefetch`` stay."""
        self.iteration += 1
        self.entries = {k: e for k, e in self.entries.items() if e["iteration"] >= self.iteration}

    def _upload(self, t, newest_only, iteration, fp=None):
        from .. import _lib as L
        cur = torch.cuda.current_stream(self.dev)
        if newest_only:
            T, B, C, H, W = t.shape
            out = torch.empty((T, B, 1, H, W), dtype=t.dtype, device=self.dev)
        else:
            out = torch.empty(t.shape, dtype=t.dtype, device=self.dev)
        self.stream.wait_stream(cur)                 # the block may have been used by earlier compute-stream work
        pinned = self._pin(t)
        with torch.cuda.stream(self.stream):
            if newest_only and t.is_contiguous() and t.dtype == torch.uint8:
                L.call("cb_h2d_2d", out.data_ptr(), H * W, t.data_ptr() + (C - 1) * H * W, C * H * W, H * W, T * B,
                       self.stream.cuda_stream)
            elif newest_only:
                out.copy_(t[:, :, -1:], non_blocking=pinned)
            else:
                out.copy_(t, non_blocking=pinned)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        out.record_stream(self.stream)
        self.generation += 1
        out._b200_staged = self.generation       # filled once, never refilled: the models may trust object identity
        self.bytes_uploaded += out.numel() * out.element_size()
        return dict(tensor=out, event=ev, fp=fp if fp is not None else self._fingerprint(t), iteration=iteration,
                    generation=self.generation)

    def _key(self, t, newest_only, iteration):
        # the iteration is part of the key: a runner that refills ONE host buffer can stage iteration i+1 while
        # iteration i's copy is still being consumed
        return (t.data_ptr(), tuple(t.shape), t.dtype, newest_only, iteration)

    def prefetch(self, t, newest_only=False):
        """Stage ``t`` for the NEXT iteration without making the compute stream wait."""
        if t.is_cuda:
            return
        self.entries[self._key(t, newest_only, self.iteration + 1)] = self._upload(t, newest_only, self.iteration + 1)

    def get(self, t, newest_only=False):
        """Device copy of ``t`` (uploaded now unless this iteration already did); the compute stream waits for it."""
        if t.is_cuda:
            return t
        key, fp = self._key(t, newest_only, self.iteration), self._fingerprint(t)
        e = self.entries.get(key)
        if e is None or e["fp"] != fp:
            e = self._upload(t, newest_only, self.iteration, fp)
            self.entries[key] = e
        else:
            self.hits += 1
        torch.cuda.current_stream(self.dev).wait_event(e["event"])
        return e["tensor"]

    @staticmethod
    def _spot_check_shifted(obs, nxt, words=4096):
        """next_observation[:-1] == observation[1:] on ``words`` evenly spaced 8-byte words (host side, microseconds).
        Catches a caller whose two tensors are unrelated; it is NOT a proof -- the proof is the sampler's contract."""
        import numpy as np
        if obs.shape != nxt.shape or obs.dtype != nxt.dtype or obs.shape[0] < 2:
            return obs.shape == nxt.shape and obs.dtype == nxt.dtype
        a = obs[1:].reshape(-1).view(torch.uint8)
        b = nxt[:-1].reshape(-1).view(torch.uint8)
        nw = a.numel() // 8
        if nw == 0:
            return bool(torch.equal(a, b))
        idx = torch.from_numpy(np.linspace(0, nw - 1, num=min(words, nw)).astype(np.int64))
        return bool(torch.equal(a[: nw * 8].view(torch.int64)[idx], b[: nw * 8].view(torch.int64)[idx]))

    def get_shifted(self, observation, next_observation):
        """(obs, next_obs) as two views, one time step apart, of ONE device buffer holding the T+1 distinct steps of a
        rollout: ``observation`` [T,B,...] is uploaded in full, of ``next_observation`` only the last step.  Halves the
        PCIe traffic of ICM / Disagreement and lets the models run their encoder once (icm.py ``shifted``).  Falls
        back to two plain uploads when the host-side spot check finds the tensors unrelated."""
        if observation.is_cuda or next_observation.is_cuda:
            return self.get(observation), self.get(next_observation)
        key = self._key(observation, "shifted", self.iteration)
        fp = (self._fingerprint(observation), self._fingerprint(next_observation[-1]))
        e = self.entries.get(key)
        if e is None or e["fp"] != fp:
            if not (observation.is_contiguous() and next_observation.is_contiguous()
                    and self._spot_check_shifted(observation, next_observation)):
                return self.get(observation), self.get(next_observation)
            e = self._upload_shifted(observation, next_observation, self.iteration, fp)
            self.entries[key] = e
        else:
            self.hits += 1
        torch.cuda.current_stream(self.dev).wait_event(e["event"])
        return e["tensor"]

    def prefetch_shifted(self, observation, next_observation):
        if observation.is_cuda or not self._spot_check_shifted(observation, next_observation):
            self.prefetch(observation); self.prefetch(next_observation)
            return
        fp = (self._fingerprint(observation), self._fingerprint(next_observation[-1]))
        self.entries[self._key(observation, "shifted", self.iteration + 1)] = \
            self._upload_shifted(observation, next_observation, self.iteration + 1, fp)

    def _upload_shifted(self, observation, next_observation, iteration, fp):
        T = observation.shape[0]
        cur = torch.cuda.current_stream(self.dev)
        buf = torch.empty((T + 1,) + tuple(observation.shape[1:]), dtype=observation.dtype, device=self.dev)
        self.stream.wait_stream(cur)
        p1, p2 = self._pin(observation), self._pin(next_observation)
        with torch.cuda.stream(self.stream):
            buf[:T].copy_(observation, non_blocking=p1)
            buf[T].copy_(next_observation[T - 1], non_blocking=p2)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        buf.record_stream(self.stream)
        self.generation += 1
        self.bytes_uploaded += buf.numel() * buf.element_size()
        obs, nxt = buf[:T], buf[1:]
        obs._b200_staged = nxt._b200_staged = self.generation
        return dict(tensor=(obs, nxt), event=ev, fp=fp, iteration=iteration, generation=self.generation, buffer=buf)

    def alias(self, host_tensor, dev_tensor):
        """Remember that ``host_tensor`` is a read-back of ``dev_tensor`` (process_returns hands ``valid`` back to the
        caller on the host and the caller passes it on to ``curiosity_loss``, ppo.py:93-110): ``get`` then returns the
        device original instead of uploading the copy again."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.dev))
        self.entries[self._key(host_tensor, False, self.iteration)] = dict(
            tensor=dev_tensor, event=ev, fp=self._fingerprint(host_tensor), iteration=self.iteration,
            generation=self.generation)

    def generation_of(self, dev_tensor):
        """Generation stamp of a staged device tensor (None if it is not one of ours): the curiosity models compare it
        instead of ``_version``, which a raw-pointer upload never bumps."""
        for e in self.entries.values():
            if e["tensor"] is dev_tensor:
                return e["generation"]
        return None


def staging(agent):
    """The agent's ``BatchStaging`` (created on first use, on the agent's device)."""
    st = getattr(agent, "_b200_staging", None)
    if st is None:
        st = BatchStaging(_device(agent))
        try:
            agent._b200_staging = st
        except AttributeError:
            pass
    return st


def prefetch_samples(agent, curiosity_type, samples):
    """Stage the frame tensors of the NEXT iteration's ``samples`` while the current one is being optimised
    (double-buffered runners)."""
    st = staging(agent)
    if curiosity_type == "rnd":
        st.prefetch(samples.env.next_observation, newest_only=True)
    elif curiosity_type in ("icm", "disagreement"):
        if st.assume_shifted:
            st.prefetch_shifted(samples.env.observation, samples.env.next_observation)
        else:
            st.prefetch(samples.env.observation)
            st.prefetch(samples.env.next_observation)
    elif curiosity_type == "ndigo":
        st.prefetch(samples.env.observation)


def _frames_pair(st, observation, next_observation):
    if st.assume_shifted:
        return st.get_shifted(observation, next_observation)
    return st.get(observation), st.get(next_observation)


def _onehot_staged(agent, st, idx, dev):
    """One-hot of an int64 action tensor; a host tensor is uploaded once per iteration through the staging cache."""
    return _onehot(agent, st.get(torch.as_tensor(idx)), dev)


@torch.no_grad()
def curiosity_step(self, curiosity_type, *args):
    dev = _device(self)
    cm = _curiosity_model(self)
    st = staging(self)
    st.new_iteration()                      # the bonus is the first consumer of an iteration's samples
    if curiosity_type in ("icm", "disagreement"):
        observation, next_observation, actions = args
        obs_d, nxt_d = _frames_pair(st, observation, next_observation)
        r_int = cm.compute_bonus(obs_d, nxt_d, _onehot_staged(self, st, actions, dev))
        info = IcmInfo()
    elif curiosity_type == "ndigo":
        observation, prev_actions, actions = args
        r_int = cm.compute_bonus(st.get(observation), _onehot_staged(self, st, prev_actions, dev),
                                 _onehot_staged(self, st, actions, dev))
        info = NdigoInfo(prev_gru_state=None)
    elif curiosity_type == "rnd":
        next_observation, done = args
        frames = next_observation if next_observation.is_cuda else st.get(next_observation, newest_only=True)
        r_int = cm.compute_bonus(frames, st.get(torch.as_tensor(done)).to(torch.float32))
        info = RndInfo()
    else:
        raise ValueError(curiosity_type)
    return AgentCuriosityStep(r_int=r_int.to("cpu"), agent_curiosity_info=info)


def curiosity_loss(self, curiosity_type, *args):
    dev = _device(self)
    cm = _curiosity_model(self)
    st = staging(self)
    if curiosity_type in ("icm", "disagreement"):
        observation, next_observation, actions, valid = args
        obs_d, nxt_d = _frames_pair(st, observation, next_observation)
        inv, fwd = cm.compute_loss(obs_d, nxt_d, _onehot_staged(self, st, actions, dev), st.get(torch.as_tensor(valid)))
        return inv.to("cpu"), fwd.to("cpu")
    if curiosity_type == "ndigo":
        observations, prev_actions, actions, valid = args
        loss = cm.compute_loss(st.get(observations), _onehot_staged(self, st, prev_actions, dev),
                               _onehot_staged(self, st, actions, dev), st.get(torch.as_tensor(valid)))
        return loss.to("cpu")
    if curiosity_type == "rnd":
        next_observation, valid = args
        frames = next_observation if next_observation.is_cuda else st.get(next_observation, newest_only=True)
        loss = cm.compute_loss(frames, st.get(torch.as_tensor(valid)))
        return loss.to("cpu")
    raise ValueError(curiosity_type)
