#!/usr/bin/env python
"""Benchmark of the intrinsic-reward hot path (BASELINE.json metric: intrinsic-reward samples/sec, bonus + update).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload rnd|icm|disagreement|ndigo|ppo_icm|ppo_rnd] [--only] [--T 128] [--B per-GPU columns]

Headline (default): BASELINE configs[1] -- RND bonus + update at T=128, B=1024 per GPU, 4x84x84 uint8 frames.
One "step" = compute_bonus -> reward += r_int -> GAE / valid -> one epoch of compute_loss forward + backward over all T*B
samples -> gradient all-reduce -> global-norm clip -> Adam.  One JSON line on rank 0 with

  value        device-timed, inputs resident in HBM, weak scaling (B per GPU fixed); CUDA events, max over ranks
  strong       the same step with the GLOBAL batch fixed (B_global = B, sharded B/N per GPU)            
  e2e          through the reference-facing calls (process_returns -> agent.curiosity_step, agent.curiosity_loss,
               optimiser) with HOST buffers: frames uploaded from pinned memory by agents/hooks.BatchStaging, results
               read back; the next batch's upload is staged under this batch's compute (double-buffered runner)
  roofline     the kernel with the largest share of the step + a per-kernel table measured live (CUDA events)
  cpu_baseline the reference's own CPU implementation on the host cores (kind "reference": the unmodified reference
               from /root/reference or the vendored baseline/_ref; kind "port": the oracle, when neither is present)
  configs      short runs of the other BASELINE configs (icm, disagreement sharded, ppo_icm, ppo_rnd, ndigo), each with
               value / e2e / roofline / cpu_baseline  (skipped with --only)

`--impl reference` times the reference CPU path of the chosen workload (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from collections import namedtuple
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

METRIC = "intrinsic-reward samples/sec (bonus+update)"

# ---------------------------------------------------------------------------------------------- algorithmic FLOPs
# per sample, FLOP = 2 MAC, Burda encoder: unit of work = 2 * forward + backward
RND_MAC_T = 64 * 32 * 400 + 512 * 64 * 81 + 576 * 64 * 49 + 3136 * 512
RND_MAC_P = RND_MAC_T + 2 * 512 * 512
RND_FLOP = 2 * (2 * (RND_MAC_T + RND_MAC_P) + 2 * RND_MAC_P - 64 * 32 * 400)
ENC_MAC = 256 * 32 * 400 + 512 * 64 * 81 + 576 * 64 * 49 + 3136 * 512          # 4-channel conv1
POL_NONE_MAC = 256 * 16 * 400 + 256 * 32 * 100 + 3200 * 512                      # Conv2dHeadModel 16/32 + fc


def icm_flops(A, E):
    F = 512
    resf, inv = 10 * (F + A) * F, 2 * F * F + F * A
    fwd = 2 * ENC_MAC + inv + E * resf
    bwd = 2 * (2 * ENC_MAC - 256 * 32 * 400) + 2 * inv + E * (2 * resf - (F + A) * F)
    return 2 * (2 * fwd + bwd)


def icm_flops_dedup(A, E):
    """The same unit when next_obs[t] == obs[t+1] is exploited: ONE encoder pass (forward and backward) per phase."""
    F = 512
    resf, inv = 10 * (F + A) * F, 2 * F * F + F * A
    fwd = ENC_MAC + inv + E * resf
    bwd = (2 * ENC_MAC - 256 * 32 * 400) + 2 * inv + E * (2 * resf - (F + A) * F)
    return 2 * (2 * fwd + bwd)


def policy_macs(A, enc_mac, conv1_mac, F=512, H=512):
    lstm, heads = (F + A) * 4 * H + H * 4 * H, H * (A + 1)
    fwd = enc_mac + lstm + heads
    bwd = (2 * enc_mac - conv1_mac) + 2 * lstm - A * 4 * H + 2 * heads
    return fwd, bwd


def ppo_icm_flops(A, epochs, dedup=False):
    F = 512
    pf, pb = policy_macs(A, ENC_MAC, 256 * 32 * 400)
    resf, inv = 10 * (F + A) * F, 2 * F * F + F * A
    passes = 1 if dedup else 2
    icm_fwd = passes * ENC_MAC + inv + resf
    icm_bwd = passes * (2 * ENC_MAC - 256 * 32 * 400) + 2 * inv + (2 * resf - (F + A) * F)
    return 2 * (icm_fwd + epochs * (pf + pb + icm_fwd + icm_bwd))


def ppo_rnd_flops(A, epochs):
    pf, pb = policy_macs(A, POL_NONE_MAC, 256 * 16 * 400)
    rnd_fwd = RND_MAC_T + RND_MAC_P
    rnd_bwd = 2 * RND_MAC_P - 64 * 32 * 400
    return 2 * (rnd_fwd + epochs * (pf + pb + rnd_fwd + rnd_bwd))


def ndigo_flops(A=5):
    enc = 3 * 16 * 9 * 25 + 16 * 16 * 9 * 16 + 256 * 256
    gru = (256 + A) * 384 + 128 * 384
    pred = sum((128 + k * A) * 64 + 64 * 75 for k in range(1, 11))
    fwd_loss = enc + gru + pred
    fwd_bonus = enc + gru + 2 * ((128 + 4 * A) * 64 + 64 * 75)
    return 2 * (fwd_bonus + 3 * fwd_loss)


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], d["bf16_tflops_sustained"], "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- distributed helpers
class Ctx:
    def __init__(self):
        self.rank = int(os.environ.get("RANK", 0))
        self.world = int(os.environ.get("WORLD_SIZE", 1))
        self.local = int(os.environ.get("LOCAL_RANK", 0))
        self.dev = None

    def init_gpu(self):
        import torch.distributed as dist
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1 and not dist.is_initialized():
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = torch.tensor([float(x)], device=self.dev)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    def sum_over_ranks(self, x):
        t = torch.tensor([float(x)], device=self.dev, dtype=torch.float64)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t)
        return float(t)


def timed(ctx, fn, warmup, steps):
    """W untimed + exactly K timed calls of ``fn``; barrier + synchronize on both sides; CUDA events on the launching
    stream; max over ranks.  Returns (ms per step, last result)."""
    out = None
    for _ in range(warmup):
        out = fn()
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = fn()
    e1.record()
    ctx.barrier()
    return ctx.max_over_ranks(e0.elapsed_time(e1)) / steps, out


def wall_timed(ctx, fn, warmup, steps):
    """Host-facing calls block on their own device->host reads: wall clock between barriers, max over ranks."""
    out = None
    for _ in range(warmup):
        out = fn()
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = fn()
    ctx.barrier()
    return ctx.max_over_ranks((time.perf_counter() - t0) * 1e3) / steps, out


# ---------------------------------------------------------------------------------------------- synthetic rollouts
Env = namedtuple("Env", ["reward", "done", "observation", "next_observation", "prev_reward"])
AgentInfo = namedtuple("AgentInfo", ["dist_info", "value", "prev_rnn_state"])
DistInfo = namedtuple("DistInfo", ["prob"])
RnnState = namedtuple("RnnState", ["h", "c"])
AgentS = namedtuple("AgentS", ["action", "prev_action", "agent_info", "bootstrap_value"])
Samples = namedtuple("Samples", ["agent", "env"])


def synth(dev, seed, T, B, A, frames="u8", next_frames=True, rnn=False):
    """Device-resident synthetic rollout of the shapes rlpyt's Samples carry."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    d = {}
    if frames == "u8":
        d["obs"] = torch.randint(0, 256, (T, B, 4, 84, 84), dtype=torch.uint8, device=dev, generator=g)
        if next_frames:
            d["nxt"] = torch.randint(0, 256, (T, B, 4, 84, 84), dtype=torch.uint8, device=dev, generator=g)
    else:
        d["obs"] = (torch.rand(T, B, 3, 5, 5, device=dev, generator=g) < 0.3).float()
    d["act"] = torch.randint(0, A, (T, B), device=dev, generator=g)
    d["prev"] = torch.randint(0, A, (T, B), device=dev, generator=g)
    d["old_prob"] = torch.softmax(torch.randn(T, B, A, device=dev, generator=g), -1)
    first = torch.randint(1, T + 1, (B,), device=dev, generator=g)
    has = torch.rand(B, device=dev, generator=g) < 0.01
    first = torch.where(has, first, torch.full_like(first, T))
    d["done"] = (torch.arange(T, device=dev).unsqueeze(1) >= first.unsqueeze(0)).float()
    d["value"] = torch.randn(T, B, device=dev, generator=g)
    d["boot"] = torch.randn(B, device=dev, generator=g)
    d["mask"] = (torch.rand(T, B, device=dev, generator=g) > 0.25).float()
    if rnn:
        d["h0"] = torch.randn(B, 1, 512, device=dev, generator=g) * 0.1
        d["c0"] = torch.randn(B, 1, 512, device=dev, generator=g) * 0.1
    return d


def host_samples(d, T, B, rnn=False):
    """The same rollout as pinned host tensors in rlpyt's ``Samples`` layout (what the sampler hands the algorithm)."""
    pin = lambda t: t.cpu().pin_memory()
    if "nxt" in d and d["nxt"].data_ptr() == d["obs"].data_ptr() + d["obs"][0].numel() * d["obs"].element_size():
        both = torch.empty((T + 1,) + tuple(d["obs"].shape[1:]), dtype=d["obs"].dtype).pin_memory()      # collector layout
        both[:T].copy_(d["obs"]); both[T].copy_(d["nxt"][T - 1])
        obs, nxt = both[:T], both[1:]
    else:
        obs = pin(d["obs"])
        nxt = pin(d["nxt"]) if "nxt" in d else obs
    rs = None
    if rnn:
        rs = RnnState(h=pin(d["h0"]).unsqueeze(0), c=pin(d["c0"]).unsqueeze(0))      # [T=1 slice][B,N,H]
    info = AgentInfo(dist_info=DistInfo(prob=pin(d["old_prob"])), value=pin(d["value"]), prev_rnn_state=rs)
    ag = AgentS(action=pin(d["act"]), prev_action=pin(d["prev"]), agent_info=info, bootstrap_value=pin(d["boot"]).reshape(1, B))
    env = Env(reward=torch.zeros(T, B).pin_memory(), done=pin(d["done"]).bool(), observation=obs, next_observation=nxt,
              prev_reward=torch.zeros(T, B).pin_memory())
    return Samples(agent=ag, env=env)


def nbytes(*ts):
    return int(sum(t.numel() * t.element_size() for t in ts))


# ---------------------------------------------------------------------------------------------- workloads
class HookAgent:
    """What agents/hooks.py needs of an rlpyt agent (agents/pg/categorical.py:83-130 are bound as its methods)."""
    recurrent = True

    def __init__(self, model, dev, A):
        from curiosity_baselines_b200.agents import hooks
        self.model = SimpleNamespace(curiosity_model=model)
        self.device = dev
        self.distribution = SimpleNamespace(dim=A)
        self._hooks = hooks

    def curiosity_step(self, *a):
        return self._hooks.curiosity_step(self, *a)

    def curiosity_loss(self, *a):
        return self._hooks.curiosity_loss(self, *a)


class CuriosityWorkload:
    """bonus + returns + one-epoch update of one curiosity model (RND / ICM / Disagreement / NDIGO)."""

    def __init__(self, kind, ctx, T, B, A=None):
        from curiosity_baselines_b200.curiosity.icm import ICM, Disagreement
        from curiosity_baselines_b200.curiosity.ndigo import NDIGO
        from curiosity_baselines_b200.curiosity.rnd import RND
        from curiosity_baselines_b200.trainer import RndStep, IcmStep, FlatAdam
        self.kind, self.ctx, self.T, self.B = kind, ctx, T, B
        dev = ctx.dev
        torch.manual_seed(0)                               # identical weights on every rank
        if kind == "rnd":
            self.A = A or 4
            self.model = RND(image_shape=(4, 84, 84), prediction_beta=1.0, drop_probability=0.25, gamma=0.99).to(dev)
            self.driver = RndStep(self.model, lr=1e-4)
            self.flop = RND_FLOP
            self.d = synth(dev, 1234 + ctx.rank, T, B, self.A, next_frames=False)
            self.what = "RND (Burda conv target/predictor) bonus+update"
        elif kind in ("icm", "disagreement"):
            self.A = A or (4 if kind == "icm" else 7)
            cls = ICM if kind == "icm" else Disagreement
            self.model = cls(image_shape=(4, 84, 84), action_size=self.A, feature_encoding="idf_burda").to(dev)
            self.driver = IcmStep(self.model, lr=1e-4)
            self.flop = icm_flops(self.A, 1 if kind == "icm" else 4)
            self.d = synth(dev, 1234 + ctx.rank, T, B, self.A)
            self.d["onehot"] = torch.nn.functional.one_hot(self.d["act"], self.A).float()
            self.what = ("ICM" if kind == "icm" else "Disagreement (4 forward models, grouped GEMMs)") + " bonus+update"
        elif kind == "ndigo":
            self.A = A or 5
            self.model = NDIGO(image_shape=(3, 5, 5), action_size=self.A, horizon=3, feature_encoding="idf_maze",
                               num_predictors=10).to(dev)
            self.opt = FlatAdam(self.model.parameters(), lr=1e-4)
            self.driver = None
            self.flop = ndigo_flops(self.A)
            self.d = synth(dev, 1234 + ctx.rank, T, B, self.A, frames="maze")
            oh = lambda t: torch.nn.functional.one_hot(t, self.A).float()
            self.d["prev1h"], self.d["act1h"] = oh(self.d["prev"]), oh(self.d["act"])
            self.what = "NDIGO (3x5x5 frames, horizon 3, 10 predictors) bonus+update"
        else:
            raise ValueError(kind)
        self.opt = self.driver.opt if self.driver is not None else self.opt
        self.agent = HookAgent(self.model, dev, self.A)

    @property
    def frame_desc(self):
        return "3x5x5 binary frames" if self.kind == "ndigo" else "4x84x84 uint8"

    def enable_graph(self):
        """Capture the fused step in a CUDA graph (trainer.GraphedStep); False when capture is not possible."""
        from curiosity_baselines_b200.trainer import GraphedStep
        if self.driver is None:
            return False
        d, T, B = self.d, self.T, self.B
        self._zero = torch.zeros(T, B, device=self.ctx.dev)
        self._rew = torch.zeros(T, B, device=self.ctx.dev)
        if self.kind == "rnd":
            ins = (d["obs"], d["done"], self._rew, d["value"], d["boot"], d["mask"])
        else:
            ins = (d["obs"], d["nxt"], d["onehot"], d["done"], self._rew, d["value"], d["boot"])
        try:
            self.graphed = GraphedStep(self.driver, ins)
            self._graph_ins = ins
            return True
        except Exception as e:                                  # pragma: no cover
            print(f"[bench] CUDA graph capture failed ({type(e).__name__}: {str(e)[:200]}); eager launches", file=sys.stderr)
            self.graphed = None
            return False

    def step(self):
        d, T, B = self.d, self.T, self.B
        if getattr(self, "graphed", None) is not None:
            self._rew.copy_(self._zero)                         # no_extrinsic: the step adds the bonus in place
            return self.graphed(*self._graph_ins)[0]
        rew = torch.zeros(T, B, device=self.ctx.dev)           # no_extrinsic
        if self.kind == "rnd":
            return self.driver(d["obs"], d["done"], rew, d["value"], d["boot"], d["mask"])[0]
        if self.kind in ("icm", "disagreement"):
            return self.driver(d["obs"], d["nxt"], d["onehot"], d["done"], rew, d["value"], d["boot"])[0]
        from curiosity_baselines_b200.algos.pg_base import returns_on_device
        m = self.model
        r_int = m.compute_bonus(d["obs"], d["prev1h"], d["act1h"])
        rew.add_(r_int)
        _, _, valid, _ = returns_on_device(rew, d["done"], d["value"], d["boot"], 0.99, 0.95)
        self.opt.grad.zero_()
        loss = m.compute_loss(d["obs"], d["prev1h"], d["act1h"], valid)
        loss.backward()
        if self.ctx.world > 1:
            self.opt.grad.div_(self.ctx.world)              # per-rank means -> global mean after the SUM all-reduce
        out = self.opt.log_scalar(0, loss.detach() / self.ctx.world)
        self.opt.step()
        m.invalidate_weights() if hasattr(m, "invalidate_weights") else None
        return out

    # ---- host-facing path: process_returns -> agent.curiosity_step ; agent.curiosity_loss ; optimiser
    def e2e_prepare(self):
        from curiosity_baselines_b200.algos.pg_base import process_returns
        self.samples = host_samples(self.d, self.T, self.B)
        self.process_returns = process_returns
        self.algo = SimpleNamespace(curiosity_type=self.kind, agent=self.agent, normalize_reward=False,
                                    normalize_advantage=False, gae_lambda=0.95, discount=0.99, mid_batch_reset=False,
                                    kernel_params=None)
        s = self.samples
        small = nbytes(s.env.done, s.agent.agent_info.value, s.agent.bootstrap_value, s.env.reward, s.agent.action)
        if self.kind == "rnd":
            self.h2d = self.T * self.B * 84 * 84 + small + self.T * self.B * 4
        elif self.kind == "ndigo":
            self.h2d = nbytes(s.env.observation) + small + nbytes(s.agent.prev_action)
        else:
            shifted = s.env.next_observation.data_ptr() == s.env.observation.data_ptr() + s.env.observation[0].numel()
            self.h2d = (nbytes(s.env.observation) * (self.T + 1) // self.T if shifted else
                        nbytes(s.env.observation, s.env.next_observation)) + small
        self.d2h = 4 * self.T * self.B * 4 + 8              # r_int, return, advantage, valid + the loss scalars
        self.agent._hooks.prefetch_samples(self.agent, self.kind, s)

    def e2e_step(self):
        s, kind, ag = self.samples, self.kind, self.agent
        s.env.reward.zero_()
        ret, adv, valid = self.process_returns(self.algo, s)
        # the NEXT batch uploads under this iteration's loss / backward (issued after this iteration's own small
        # uploads: one DMA engine serves host->device copies in order)
        ag._hooks.prefetch_samples(ag, kind, s)
        if kind == "rnd":
            loss = ag.curiosity_loss("rnd", s.env.next_observation, valid)
        elif kind == "ndigo":
            loss = ag.curiosity_loss("ndigo", s.env.observation, s.agent.prev_action, s.agent.action, valid)
        else:
            inv, fwd = ag.curiosity_loss(kind, s.env.observation, s.env.next_observation, s.agent.action, valid)
            loss = inv + fwd
        self.opt.grad.zero_()
        loss.backward()
        if self.ctx.world > 1:
            self.opt.grad.div_(self.ctx.world)              # DDP convention on the autograd path: average of rank means
        self.opt.step()
        if hasattr(self.model, "invalidate_weights"):
            self.model.invalidate_weights()
        return float(loss)

    def set_reuse(self, on):
        if hasattr(self.model, "reuse_features"):
            self.model.reuse_features = on
            self.model._cache = None
            return True
        return False

    def reuse_hits(self):
        return dict(getattr(self.model, "reuse_hits", {}))

    def enable_dedup(self):
        """Lay the rollout out as the collectors produce it: next_obs[t] IS obs[t+1] (two views of one T+1 buffer)."""
        if self.kind not in ("icm", "disagreement"):
            return False
        d, T = self.d, self.T
        buf = torch.empty((T + 1,) + tuple(d["obs"].shape[1:]), dtype=torch.uint8, device=self.ctx.dev)
        buf[:T].copy_(d["obs"])
        buf[T].copy_(d["nxt"][T - 1])
        d["obs"], d["nxt"] = buf[:T], buf[1:]
        self.agent._hooks.staging(self.agent).assume_shifted = True
        self.flop_dedup = icm_flops_dedup(self.A, 1 if self.kind == "icm" else 4)
        return True


class PpoWorkload:
    """Full PPO iteration (ppo.py:72-190): bonus -> returns -> epochs x minibatches of [policy forward, PPO loss, policy
    backward, curiosity loss forward/backward, one clip + Adam over policy U curiosity parameters]."""

    def __init__(self, kind, ctx, T, B, A=None, epochs=3, minibatches=1):
        import numpy as np
        from curiosity_baselines_b200.algos.ppo import PPO
        from curiosity_baselines_b200.pg.atari_lstm_model import AtariLstmModel
        self.kind, self.ctx, self.T, self.B, self.A = kind, ctx, T, B, A or 4
        self.epochs, self.minibatches = epochs, minibatches
        dev = ctx.dev
        torch.manual_seed(0)
        if kind == "ppo_icm":
            ck = dict(curiosity_alg="icm", feature_encoding="idf_burda", batch_norm=False, prediction_beta=1.0,
                      forward_loss_wt=0.2)
            self.ctype, self.flop = "icm", ppo_icm_flops(self.A, epochs)
            self.what = "ICM + LSTM PPO full iteration"
        else:
            ck = dict(curiosity_alg="rnd", feature_encoding="none", prediction_beta=1.0, drop_probability=0.25, gamma=0.99)
            self.ctype, self.flop = "rnd", ppo_rnd_flops(self.A, epochs)
            self.what = "RND + LSTM PPO (GAE) full iteration"
        self.model = AtariLstmModel((4, 84, 84), self.A, curiosity_kwargs=ck).to(dev)
        with torch.no_grad():       # damp the policy encoder so raw 0..255 pixels give O(1) features (non-saturated gates)
            for prm in self.model.conv.parameters():
                prm.mul_(0.65)
        self.algo = PPO(discount=0.99, learning_rate=1e-4, gae_lambda=0.95, minibatches=minibatches, epochs=epochs,
                        curiosity_type=self.ctype, linear_lr_schedule=False)
        self.algo.initialize(self.model, n_itr=10 ** 6, rng=np.random.RandomState(ctx.rank))
        self.opt = self.algo.opt
        self.d = synth(dev, 1234 + ctx.rank, T, B, self.A, rnn=True)
        self.agent = SimpleNamespace(device=dev, model=self.model)

    frame_desc = "4x84x84 uint8"

    def step(self):
        from curiosity_baselines_b200.algos.ppo import PpoBatch
        d = self.d
        batch = PpoBatch(d["obs"], d["nxt"], d["prev"], d["act"], d["old_prob"], torch.zeros(self.T, self.B, device=self.ctx.dev),
                         d["done"], d["value"], d["boot"], (d["h0"], d["c0"]))
        return self.algo.optimize_agent(0, batch)[-1].loss

    def e2e_prepare(self):
        from curiosity_baselines_b200.agents import hooks
        self.hooks = hooks
        self.samples = host_samples(self.d, self.T, self.B, rnn=True)
        s = self.samples
        shifted = s.env.next_observation.data_ptr() == s.env.observation.data_ptr() + s.env.observation[0].numel()
        frames = nbytes(s.env.observation) * (self.T + 1) // self.T if shifted else nbytes(s.env.observation, s.env.next_observation)
        if self.ctype == "rnd":
            frames = nbytes(s.env.observation) + nbytes(s.env.next_observation) // 4
        self.h2d = frames + nbytes(s.agent.action, s.agent.prev_action,
                          s.agent.agent_info.dist_info.prob, s.agent.agent_info.value, s.env.done, s.env.reward,
                          s.agent.bootstrap_value, s.agent.agent_info.prev_rnn_state.h, s.agent.agent_info.prev_rnn_state.c)
        self.d2h = 2 * self.T * self.B * 4 + 4 * 12 * self.epochs * self.minibatches
        self._prefetch()

    def _prefetch(self):
        st = self.hooks.staging(self.agent)
        if self.ctype == "rnd":
            st.prefetch(self.samples.env.observation)
            st.prefetch(self.samples.env.next_observation, newest_only=True)
        elif getattr(self, "hooks_staging_shifted", False):
            st.assume_shifted = True
            st.prefetch_shifted(self.samples.env.observation, self.samples.env.next_observation)
        else:
            st.prefetch(self.samples.env.observation)
            st.prefetch(self.samples.env.next_observation)

    def e2e_step(self):
        from curiosity_baselines_b200.algos.ppo import optimize_agent_from_samples
        self.samples.env.reward.zero_()
        # the NEXT batch is staged right after this iteration's own (small) uploads: one DMA engine serves host->device
        # copies in order, so issuing it first would stall them behind 7-15 GB of frames
        info = optimize_agent_from_samples(self.algo, self.agent, 0, self.samples, after_staging=self._prefetch)
        return info.loss[-1]

    def set_reuse(self, on):
        cm = self.model.curiosity_model
        cm.reuse_features = on
        cm._cache = None
        return True

    def reuse_hits(self):
        return dict(self.model.curiosity_model.reuse_hits)

    def enable_dedup(self):
        if self.ctype != "icm":
            return False
        d, T = self.d, self.T
        buf = torch.empty((T + 1,) + tuple(d["obs"].shape[1:]), dtype=torch.uint8, device=self.ctx.dev)
        buf[:T].copy_(d["obs"])
        buf[T].copy_(d["nxt"][T - 1])
        d["obs"], d["nxt"] = buf[:T], buf[1:]
        self.hooks_staging_shifted = True
        self.flop_dedup = ppo_icm_flops(self.A, self.epochs, dedup=True)
        return True


def make_workload(kind, ctx, T, B, args):
    if kind.startswith("ppo_"):
        return PpoWorkload(kind, ctx, T, B, args.A, args.epochs, args.minibatches)
    return CuriosityWorkload(kind, ctx, T, B, args.A)


DEFAULT_B = {"rnd": 1024, "icm": 1024, "disagreement": 1024, "ndigo": 4096, "ppo_icm": 2048, "ppo_rnd": 1024}
CONFIG_ID = {"rnd": "configs[1] (curiosity-only unit)", "icm": "configs[0] shape at B=1024", "disagreement": "configs[2]",
             "ndigo": "configs[4]", "ppo_icm": "configs[3]", "ppo_rnd": "configs[1] as a full PPO iteration"}


# ---------------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference(kind, T, B, A, epochs, steps, warmup, seconds_cap):
    """The reference's own CPU implementation of the workload on all host threads: the UNMODIFIED reference (from
    /root/reference or the vendored baseline/_ref) when importable, else the oracle port.  Returns
    (samples/s, ms/step, steps done, kind)."""
    torch.set_num_threads(os.cpu_count())
    import recipes as R
    from oracle import refshim
    step, which = None, "port"
    if refshim.available():
        try:
            step = _reference_step(kind, T, B, A, epochs, R, refshim)
            which = "reference"
        except Exception as e:                                      # pragma: no cover
            print(f"[bench] reference import failed ({type(e).__name__}: {e}); timing the oracle port", file=sys.stderr)
    if step is None:
        step = _port_step(kind, T, B, A, epochs, R)
    for _ in range(warmup):
        step()
    t0, k = time.perf_counter(), 0
    for _ in range(steps):
        step()
        k += 1
        if time.perf_counter() - t0 > seconds_cap:
            break
    dt = (time.perf_counter() - t0) / k
    return T * B / dt, dt * 1e3, k, which


def _cpu_inputs(kind, T, B, A, R):
    if kind == "ndigo":
        obs = R.make_binary_frames(3, T, B)
        nxt = obs
    else:
        obs, nxt = R.make_frames(3, T, B), R.make_frames(4, T, B)
    return dict(obs=obs, nxt=nxt, prev=R.make_actions(5, T, B, A), act=R.make_actions(6, T, B, A),
                old_prob=R.make_probs(7, T, B, A), done=R.make_suffix_done(8, T, B, frac=0.01),
                value=R.make_normal(9, T, B), boot=R.make_normal(10, B), mask=R.make_mask(11, (T, B), 0.25))


def _reference_step(kind, T, B, A, epochs, R, refshim):
    """One unit of work driven through the reference's own objects."""
    refshim.import_reference()
    x = _cpu_inputs(kind, T, B, A, R)
    if kind.startswith("ppo_"):
        ctype = kind[4:]
        ck = (dict(curiosity_alg="icm", feature_encoding="idf_burda", batch_norm=False, prediction_beta=1.0,
                   forward_loss_wt=0.2, device="cpu") if ctype == "icm" else
              dict(curiosity_alg="rnd", feature_encoding="none", prediction_beta=1.0, drop_probability=0.25, gamma=0.99,
                   device="cpu"))
        agent, algo = refshim.make_agent_and_algo(ck, A, T, B, n_itr=10 ** 6,
                                                  ppo_kwargs=dict(epochs=epochs, minibatches=1, gae_lambda=0.95,
                                                                  learning_rate=1e-4, linear_lr_schedule=False))
        agent.train_mode(0)

        def step():
            s = refshim.make_samples(x["obs"], x["nxt"], x["act"], x["prev"], x["old_prob"], x["value"], x["boot"],
                                     torch.zeros(T, B), x["done"])
            algo.optimize_agent(0, s)
        return step
    from rlpyt.algos.utils import generalized_advantage_estimation, valid_from_done
    onehot = lambda t: torch.nn.functional.one_hot(t, A).float()
    if kind == "rnd":
        from rlpyt.models.curiosity.rnd import RND
        m = RND(image_shape=(4, 84, 84), prediction_beta=1.0, drop_probability=0.25, gamma=0.99, device="cpu")
        bonus = lambda: m.compute_bonus(x["nxt"], x["done"])
        loss = lambda valid: m.compute_loss(x["nxt"], valid)
    elif kind in ("icm", "disagreement"):
        if kind == "icm":
            from rlpyt.models.curiosity.icm import ICM as Cls
            m = Cls(image_shape=(4, 84, 84), action_size=A, feature_encoding="idf_burda")
        else:
            from rlpyt.models.curiosity.disagreement import Disagreement as Cls
            m = Cls(image_shape=(4, 84, 84), action_size=A, feature_encoding="idf_burda", device="cpu")
        bonus = lambda: m.compute_bonus(x["obs"], x["nxt"], onehot(x["act"]))
        loss = lambda valid: sum(m.compute_loss(x["obs"], x["nxt"], onehot(x["act"]), valid))
    else:
        from rlpyt.models.curiosity.ndigo import NDIGO
        m = NDIGO(image_shape=(3, 5, 5), action_size=A, horizon=3, feature_encoding="idf_maze", num_predictors=10, device="cpu")
        bonus = lambda: m.compute_bonus(x["obs"], onehot(x["prev"]), onehot(x["act"]))
        loss = lambda valid: m.compute_loss(x["obs"], onehot(x["prev"]), onehot(x["act"]), valid)
    train = [p for p in m.parameters() if p.requires_grad]
    opt = torch.optim.Adam(train, lr=1e-4)

    def step():
        with torch.no_grad():
            r_int = bonus()
        generalized_advantage_estimation(r_int.clone(), x["value"], x["done"], x["boot"], 0.99, 0.95)
        valid = valid_from_done(x["done"])
        opt.zero_grad()
        loss(valid).backward()
        torch.nn.utils.clip_grad_norm_(train, 1.0)
        opt.step()
    return step


def _port_step(kind, T, B, A, epochs, R):
    """The same unit of work on the oracle (torch CPU fp32 restatement), used when the reference is not importable."""
    from oracle import curiosity_oracle as O
    x = _cpu_inputs(kind, T, B, A, R)
    onehot = lambda t: R.onehot(t, A)
    pol = None
    ctype = kind[4:] if kind.startswith("ppo_") else kind
    if kind.startswith("ppo_"):
        enc = "idf_burda" if ctype == "icm" else "none"
        pol = {k: v.clone().requires_grad_(True) for k, v in R.make_policy_params(A, 1, feature_encoding=enc).items()}
    if ctype == "rnd":
        q = {k: v.clone().requires_grad_(k.startswith("forward")) for k, v in R.make_params(R.rnd_shapes(), 2, gain=1.4).items()}
        o = O.RndOracle(drop_probability=0.25)
        bonus = lambda: o.compute_bonus(q, x["nxt"], x["done"])
        closs = lambda valid: o.compute_loss(q, x["nxt"], valid, x["mask"])
    elif ctype in ("icm", "disagreement"):
        dis = ctype == "disagreement"
        shapes = R.disagreement_shapes(A, "idf_burda") if dis else R.icm_shapes(A, "idf_burda")
        q = {k: v.clone().requires_grad_(True) for k, v in R.make_params(shapes, 2).items()}
        fb, fl = (O.disagreement_bonus, O.disagreement_loss) if dis else (O.icm_bonus, O.icm_loss)
        bonus = lambda: fb(q, x["obs"], x["nxt"], onehot(x["act"]))
        closs = lambda valid: sum(fl(q, x["obs"], x["nxt"], onehot(x["act"]), valid))
    else:
        q = {k: v.clone().requires_grad_(True) for k, v in R.make_params(R.ndigo_shapes(A), 2).items()}
        bonus = lambda: O.ndigo_bonus(q, x["obs"], onehot(x["prev"]), onehot(x["act"]), 3)
        closs = lambda valid: O.ndigo_loss(q, x["obs"], onehot(x["prev"]), onehot(x["act"]))
    train = [v for v in q.values() if v.requires_grad] + (list(pol.values()) if pol else [])
    opt = torch.optim.Adam(train, lr=1e-4)

    def step():
        with torch.no_grad():
            r_int = bonus()
        adv, ret = O.generalized_advantage_estimation(r_int.clone(), x["value"], x["done"], x["boot"], 0.99, 0.95)
        valid = O.valid_from_done(x["done"])
        for _ in range(epochs if pol else 1):
            opt.zero_grad()
            loss = closs(valid)
            if pol:
                enc = "idf_burda" if ctype == "icm" else "none"
                pi, v, _ = O.policy_forward(pol, x["obs"], onehot(x["prev"]), None, feature_encoding=enc)
                loss = loss + O.ppo_loss(pi, v, x["act"], x["old_prob"], ret, adv, valid, 0.1, 1.0, 0.01)[0]
            loss.backward()
            torch.nn.utils.clip_grad_norm_(train, 1.0)
            opt.step()
    return step


def cpu_baseline_record(kind, T, A, epochs, cpu_B, seconds):
    B = cpu_B * (8 if kind == "ndigo" else 1)
    v, ms, k, which = cpu_reference(kind, T, B, A, epochs, 1000, 1, seconds)
    return {"value": v, "unit": "samples/s", "cores": os.cpu_count(), "cpu_model": cpu_model(), "kind": which,
            "sample": f"T={T}, B={B} ({T * B} samples/step) of the same workload, {k} steps, {ms:.0f} ms/step, all host threads"}


# ---------------------------------------------------------------------------------------------- per-kernel rooflines
def kernel_table(model, T, B, ms_per_step, dev):
    """Event-timed single launches of the RND step's GEMM-shaped kernels at the step's own shapes, each with its
    algorithmic bytes (every operand read or written once: bf16 activations, uint8 frames) and FLOPs, the number of
    launches per step and its share of the step.  ncu DRAM traffic per launch is attached where a capture of the same
    kernel at the same size is recorded in profiles/r2_kernel_traffic.json."""
    from curiosity_baselines_b200 import _lib as L
    from curiosity_baselines_b200 import dist as D
    hbm, tf_burst, tf_sust, src = peaks()
    n = T * B
    reducer, D._reducer = D._reducer, None          # single launches, no data-parallel gradient exchange
    pred, targ = model._chains
    c1, c2, c3, fc, l9, l11 = pred.layers
    bf = lambda *s: torch.randn(*s, device=dev).to(torch.bfloat16)
    frames = torch.randint(0, 256, (n, 1, 84, 84), dtype=torch.uint8, device=dev)
    st1, nm = (7056, 84, 1, 7056), model._norm
    a1, a2, a3, h = bf(n * 400, 32), bf(n * 81, 64), bf(n * 49, 64), bf(n, 512)
    fl12 = 2 * (64 * 32 * 400 + 512 * 64 * 81)
    if getattr(model, "fuse_conv12", False):        # default: frames normalised once per pass, conv1 -> conv2 of both networks fused
        img = torch.empty(n * model.S2D_BYTES, dtype=torch.uint8, device=dev)
        head = [
            ("normalise frames (block order)", lambda: L.call("cb_rnd_normalize_s2d", frames.data_ptr(), 7056, n, L.ptr(nm[0]),
                                                               L.ptr(nm[1]), L.ptr(img), L.stream()), 2, 7056 + 14112, 4 * 7056),
            ("conv1+conv2 twin fwd (target+predictor)", lambda: model._conv12_twin(img, n, False), 1, 14112 + 2 * 10368, 2 * fl12),
            ("conv1+conv2 twin fwd, predictor a1 kept", lambda: model._conv12_twin(img, n, True), 1, 14112 + 2 * 10368 + 25600, 2 * fl12),
        ]
    else:
        head = [
            ("conv1 twin fwd (target+predictor)", lambda: c1.forward(frames.data_ptr(), n, in_dtype=L.U8, strides=st1, norm=nm, twin=targ.layers[0]),
             2, 7056 + 2 * 25600, 2 * 2 * 64 * 32 * 400),
            ("conv2 fwd", lambda: c2.forward(a1, n), 4, 25600 + 10368, 2 * 512 * 64 * 81),
        ]
    cases = head + [
        # name, fn, launches per step, algorithmic bytes per sample, FLOP per sample
        ("conv3 fwd", lambda: c3.forward(a2, n), 4, 10368 + 6272, 2 * 576 * 64 * 49),
        ("fc 3136->512 fwd", lambda: fc.forward(a3, n), 4, 6272 + 1024, 2 * 3136 * 512),
        ("linear 512x512 fwd", lambda: l9.forward(h, n), 4, 2048, 2 * 512 * 512),
        ("linear 512x512 dgrad", lambda: l9.dgrad(h, n, x_saved=h, x_act="relu"), 2, 3072, 2 * 512 * 512),
        ("fc dgrad", lambda: fc.dgrad(h, n, x_saved=a3, x_act="lrelu"), 1, 1024 + 2 * 6272, 2 * 3136 * 512),
        ("conv3 dgrad", lambda: c3.dgrad(a3, n, x_saved=a2, x_act="lrelu"), 1, 6272 + 2 * 10368, 2 * 576 * 64 * 81),
        ("conv2 dgrad", lambda: c2.dgrad(a2, n, x_saved=a1, x_act="lrelu"), 1, 10368 + 2 * 25600, 2 * 512 * 64 * 81),
        ("linear 512x512 wgrad", lambda: l9.wgrad(h, h, n), 2, 2048, 2 * 512 * 512),
        ("fc wgrad", lambda: fc.wgrad(a3, h, n), 1, 6272 + 1024, 2 * 3136 * 512),
        ("conv3 wgrad", lambda: c3.wgrad(a2, a3, n), 1, 10368 + 6272, 2 * 576 * 64 * 49),
        ("conv2 wgrad", lambda: c2.wgrad(a1, a2, n), 1, 25600 + 10368, 2 * 512 * 64 * 81),
        ("conv1 wgrad", lambda: c1.wgrad(frames.data_ptr(), a1, n, in_dtype=L.U8, strides=st1, norm=nm), 1, 7056 + 25600, 2 * 64 * 32 * 400),
    ]
    traffic = {}
    tp = os.path.join(ROOT, "profiles", "r2_kernel_traffic.json")
    if os.path.exists(tp) and n == 131072:
        traffic = json.load(open(tp))
    rows = []
    for name, fn, per_step, bytes_ps, flop_ps in cases:
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 4
        gbs, tfl = bytes_ps * n / ms / 1e6, flop_ps * n / ms / 1e9
        rows.append({"kernel": name, "ms": round(ms, 4), "launches_per_step": per_step,
                     "share_of_step": round(per_step * ms / ms_per_step, 4), "algorithmic_bytes": int(bytes_ps * n),
                     "hbm_gbs": round(gbs, 1), "hbm_frac": round(gbs / hbm, 4), "tflops": round(tfl, 1),
                     "tensor_frac_of_burst": round(tfl / tf_burst, 4),
                     "bound": "hbm" if flop_ps / bytes_ps < tf_burst * 1e3 / hbm else "tensor",
                     "traffic": traffic.get(name)})
    for p in pred.layers:                                        # the timing calls left gradients behind
        for prm in (p.module.weight, p.module.bias):
            if prm.grad is not None:
                prm.grad.zero_()
    D._reducer = reducer
    rows.sort(key=lambda r: -r["share_of_step"])
    top = rows[0]
    frac = top["hbm_frac"] if top["bound"] == "hbm" else top["tensor_frac_of_burst"]
    roofline = {"bound": top["bound"], "kernel": top["kernel"],
                "achieved": top["hbm_gbs"] if top["bound"] == "hbm" else top["tflops"],
                "peak": hbm if top["bound"] == "hbm" else tf_burst, "unit": "GB/s" if top["bound"] == "hbm" else "TFLOP/s",
                "frac": frac, "traffic": top["traffic"], "kernel_ms": top["ms"], "share_of_step": top["share_of_step"],
                "algorithmic_bytes": top["algorithmic_bytes"],
                "peak_source": f"{src} " + ("HBM copy bandwidth" if top["bound"] == "hbm" else "bf16 burst"),
                "kernels": rows[:8],
                "note": "by arithmetic intensity these 32/64-channel layers sit on the HBM side of the roofline; the measured "
                        "limiter of the window / first-conv kernels is the tensor cores' operand fetch from shared memory "
                        "(~(M+N)/4 cycles per K=16 tcgen05.mma: 52 at M128 N64, against 32 for the math), see "
                        "profiles/r2_fused_conv12_study.txt"}
    return roofline


def whole_step_roofline(flop, T, B, ms):
    hbm, tf_burst, tf_sust, src = peaks()
    tfl = flop * T * B / (ms * 1e-3) / 1e12
    return {"bound": "tensor", "kernel": "whole step", "achieved": tfl, "peak": tf_sust, "unit": "TFLOP/s",
            "frac": tfl / tf_sust, "traffic": None, "peak_source": f"{src} bf16 sustained", "mflop_per_sample": flop / 1e6}


# ---------------------------------------------------------------------------------------------- one measured workload
def h2d_bandwidth(ctx, nbytes_=1 << 30):
    """Aggregate pinned host -> device copy rate with every rank copying at once (names the e2e limiter at N > 1)."""
    h = torch.empty(nbytes_, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes_, dtype=torch.uint8, device=ctx.dev)
    d.copy_(h, non_blocking=True)
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    ctx.barrier()
    dt = ctx.max_over_ranks(time.perf_counter() - t0) / 3
    return nbytes_ / dt / 1e9


def run_workload(kind, ctx, args, T, B, steps, warmup, headline):
    """Device-timed value (+ strong-scaling leg, feature-reuse leg), host-facing e2e, rooflines.  Returns a dict (rank 0
    keeps it) without the CPU baseline, which is timed after the GPU work of every workload is done."""
    from curiosity_baselines_b200 import _lib as L
    w = make_workload(kind, ctx, T, B, args)
    n_global = T * B * ctx.world
    sampler = ClockSampler(ctx.local) if (headline and ctx.rank == 0) else None
    if sampler:
        sampler.start()
    for _ in range(warmup):
        w.step()
    ctx.barrier()
    l0 = L.launch_count()
    ms, loss = timed(ctx, w.step, 0, steps)
    launches = L.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    loss_val = float(loss)
    rec = {"workload": f"{w.what}, T={T}, B={B}/GPU, {w.frame_desc}" + (f", A={w.A}" if kind != "rnd" else ""),
           "baseline_config": CONFIG_ID[kind], "value": n_global / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms,
           "steps": steps, "warmup": warmup, "gpu_launches": launches, "loss": loss_val, "clocks": clocks,
           "scaling": "weak", "global_B": B * ctx.world}
    if kind.startswith("ppo_"):
        rec["workload"] += f", {w.epochs} epochs x {w.minibatches} minibatches"
    rec["step_roofline"] = whole_step_roofline(w.flop, T, B, ms)
    # ---- the same step replayed from a CUDA graph (one launch per step instead of ~85-300)
    if args.graph and hasattr(w, "enable_graph") and w.enable_graph():
        gms, gloss = timed(ctx, w.step, 3, steps)
        rec["cuda_graph"] = {"ms_per_step": gms, "value": n_global / (gms * 1e-3), "unit": "samples/s",
                             "loss": float(gloss), "note": "trainer.GraphedStep: the whole step captured once, replayed per step"}
        w.graphed = None
    # ---- RND with the separate first / second convolution kernels (the fused twin kernel is the default)
    if headline and kind == "rnd" and getattr(w.model, "fuse_conv12", None) is True:
        w.model.fuse_conv12 = False
        fms, _ = timed(ctx, w.step, 2, max(3, steps // 2))
        w.model.fuse_conv12 = True
        rec["with_separate_conv12"] = {"ms_per_step": fms, "value": n_global / (fms * 1e-3), "unit": "samples/s",
                                       "note": "RND.fuse_conv12 = False: twin first-conv kernel + window conv2 per network "
                                               "instead of cb_rnd_normalize_s2d + cb_rnd_conv12_twin_fwd"}
    # ---- the same iteration with exact feature reuse between bonus and loss (NOT the headline value)
    if not args.no_reuse and w.set_reuse(True):
        h0 = w.reuse_hits()
        rms, _ = timed(ctx, w.step, 2, max(3, steps // 2))
        h1 = w.reuse_hits()
        rec["with_feature_reuse"] = {"ms_per_step": rms, "value": n_global / (rms * 1e-3), "unit": "samples/s",
                                     "hits": {k: h1[k] - h0.get(k, 0) for k in h1},
                                     "note": "the loss pass reuses the bonus pass's forward of the same batch (exact; "
                                             "guarded by object identity of staged tensors or a device-side comparison); "
                                             "not the headline value"}
        w.set_reuse(False)
    # ---- end to end through the host-facing calls
    # (a PPO batch is 2 x 7.4 GB of pinned host memory per rank: the host-facing leg of the PPO workloads is measured
    # at N <= 2 only so that an 8-rank run does not pin > 100 GB)
    if not args.no_e2e and not (kind.startswith("ppo_") and ctx.world > 2):
        w.e2e_prepare()
        k = max(3, min(steps, 8))
        ems, _ = wall_timed(ctx, w.e2e_step, 2, k)
        rec["e2e"] = {"value": n_global / (ems * 1e-3), "unit": "samples/s", "ms_per_step": ems, "steps": k,
                      "h2d_bytes_per_step": w.h2d, "d2h_bytes_per_step": w.d2h,
                      "path": ("algos.pg_base.process_returns -> agents.hooks.curiosity_step, agents.hooks.curiosity_loss "
                               "(autograd) -> FlatAdam" if not kind.startswith("ppo_") else
                               "algos.ppo.optimize_agent_from_samples (what install(policy=True, algo=True) binds to "
                               "rlpyt's PPO.optimize_agent)"),
                      "staging": "pinned host Samples, uploaded once per iteration by hooks.BatchStaging on a side stream; "
                                 "the next batch is staged under this batch's compute"}
        w.samples = None
    # ---- observation de-duplication (next_obs[t] == obs[t+1], the layout rlpyt's collectors produce): T+1 steps are
    # uploaded and encoded instead of 2T.  Reported separately; the headline value above encodes both tensors in full.
    if not args.no_dedup and w.enable_dedup():
        dms, _ = timed(ctx, w.step, 2, max(3, steps // 2))
        hbm, tf_burst, tf_sust, src = peaks()
        dd = {"ms_per_step": dms, "value": n_global / (dms * 1e-3), "unit": "samples/s",
              "frac_of_sustained_vs_full_flops": w.flop * T * B / (dms * 1e-3) / 1e12 / tf_sust,
              "frac_of_sustained_vs_dedup_flops": w.flop_dedup * T * B / (dms * 1e-3) / 1e12 / tf_sust,
              "note": "exact: the encoder runs once over the T+1 distinct steps (icm.py shifted()); not the headline value"}
        if not args.no_e2e and not (kind.startswith("ppo_") and ctx.world > 2):
            w.e2e_prepare()
            k = max(3, min(steps, 8))
            ems, _ = wall_timed(ctx, w.e2e_step, 2, k)
            dd["e2e"] = {"value": n_global / (ems * 1e-3), "unit": "samples/s", "ms_per_step": ems,
                         "h2d_bytes_per_step": w.h2d, "note": "hooks.BatchStaging.get_shifted: T+1 steps cross PCIe"}
            w.samples = None
        rec["with_obs_dedup"] = dd
    if headline:
        rec["roofline"] = kernel_table(w.model, T, B, ms, ctx.dev)
    A, flop = w.A, w.flop
    del w
    torch.cuda.empty_cache()
    # ---- strong scaling: global batch fixed, B/N columns per GPU
    if kind == "disagreement" and ctx.world > 1:
        rec["scaling"] = "strong"                       # configs[2]: the B=1024 batch is already sharded over the GPUs
    elif ctx.world > 1 and not args.no_strong and B % ctx.world == 0:
        Bs = B // ctx.world
        ws = make_workload(kind, ctx, T, Bs, args)
        sms, _ = timed(ctx, ws.step, warmup, steps)
        rec["strong"] = {"scaling": "strong", "global_B": B, "B_per_gpu": Bs, "ms_per_step": sms,
                         "value": T * B / (sms * 1e-3), "unit": "samples/s",
                         "speedup_needs": "divide by the N=1 value of the same workload (this run's weak N=1 line)"}
        del ws
        torch.cuda.empty_cache()
    rec["_A"] = A
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", "--model", dest="workload", default="rnd",
                    choices=["rnd", "icm", "disagreement", "ndigo", "ppo_icm", "ppo_rnd"])
    ap.add_argument("--only", action="store_true", help="measure the chosen workload only (skip the other configs)")
    ap.add_argument("--T", type=int, default=128)
    ap.add_argument("--B", type=int, default=None, help="env columns per GPU (weak scaling); default per workload")
    ap.add_argument("--A", type=int, default=None)
    ap.add_argument("--epochs", type=int, default=3, help="PPO epochs per iteration (launcher default 3)")
    ap.add_argument("--minibatches", type=int, default=1)
    ap.add_argument("--engine", default="auto")
    ap.add_argument("--cpu-B", type=int, default=8)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-reuse", action="store_true")
    ap.add_argument("--no-strong", action="store_true")
    ap.add_argument("--no-dedup", action="store_true")
    ap.add_argument("--graph", action="store_true", help="also time the step replayed from a CUDA graph")
    args = ap.parse_args()
    ctx = Ctx()
    kind, T = args.workload, args.T
    B = args.B or DEFAULT_B[kind]

    if args.impl == "reference":
        if ctx.rank != 0:
            return
        A = args.A or {"disagreement": 7, "ndigo": 5}.get(kind, 4)
        cb = args.cpu_B * (8 if kind == "ndigo" else 1)
        v, ms, k, which = cpu_reference(kind, T, cb, A, args.epochs, args.steps, args.warmup, seconds_cap=120.0)
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": "samples/s", "n_gpus": args.gpus, "steps": k,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{kind} bonus+update, T={T}, {CONFIG_ID[kind]}; CPU sample B={cb}"},
            "cpu_baseline": {"value": v, "unit": "samples/s", "cores": os.cpu_count(), "cpu_model": cpu_model(),
                             "kind": which, "sample": f"T={T}, B={cb} ({T * cb} samples/step), {k} steps, all host threads"},
            "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    from curiosity_baselines_b200 import layers as LY
    ctx.init_gpu()
    numa = None
    if ctx.world > 1:
        from curiosity_baselines_b200.dist import bind_to_gpu_numa_node
        numa = bind_to_gpu_numa_node(ctx.local)          # host staging buffers local to the GPU's socket, where exposed
    LY.set_engine(args.engine)
    head = run_workload(kind, ctx, args, T, B, args.steps, args.warmup, headline=(kind == "rnd"))
    h2d_gbs = h2d_bandwidth(ctx) if not args.no_e2e else None
    others = []
    if not args.only:
        sub_steps, sub_warm = max(3, args.steps // 4), 3
        for k2 in [k for k in ("icm", "disagreement", "ppo_icm", "ppo_rnd", "ndigo") if k != kind]:
            B2 = DEFAULT_B[k2]
            if k2 == "disagreement":                     # configs[2]: the B=1024 batch sharded over the GPUs
                B2 = max(1, DEFAULT_B[k2] // ctx.world)
            steps2 = max(2, sub_steps // 2) if k2.startswith("ppo_") else sub_steps
            others.append((k2, run_workload(k2, ctx, args, T, B2, steps2, sub_warm, headline=False)))
    if ctx.world > 1:
        import torch.distributed as dist
        ctx.barrier()
        dist.destroy_process_group()
    if ctx.rank != 0:
        return
    # ---- CPU baselines last (rank 0 only; the GPUs are idle by now)
    def finish(k, rec):
        A = rec.pop("_A")
        if not args.no_cpu_baseline:
            rec["cpu_baseline"] = cpu_baseline_record(k, T, A, args.epochs, args.cpu_B,
                                                      args.cpu_seconds if k == kind else max(4.0, args.cpu_seconds / 2))
        return rec
    head = finish(kind, head)
    hbm, tf_burst, tf_sust, src = peaks()
    line = {
        "metric": METRIC, "value": head["value"], "unit": "samples/s", "n_gpus": ctx.world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": head["workload"], "baseline_config": head["baseline_config"], "global_B": head["global_B"],
                   "l2": "inputs (>= 0.9 GB of frames per GPU per step) larger than the 126 MB L2", "engine": args.engine,
                   "parallelism": f"dp{ctx.world} over B", "numa_node_rank0": numa,
                   "numa_note": None if numa is not None or ctx.world == 1 else
                   "the platform exposes no NUMA node for the GPU (sysfs numa_node = -1): ranks are not pinned"},
        "clocks": head["clocks"], "gpu_launches": head["gpu_launches"], "e2e": head.get("e2e"),
        "roofline": head.get("roofline") or head["step_roofline"], "step_roofline": head["step_roofline"],
        "cpu_baseline": head.get("cpu_baseline"), "with_feature_reuse": head.get("with_feature_reuse"),
        "with_obs_dedup": head.get("with_obs_dedup"), "with_separate_conv12": head.get("with_separate_conv12"), "cuda_graph": head.get("cuda_graph"), "strong": head.get("strong"), "loss": head["loss"], "host_to_device_gbs_all_ranks": None if h2d_gbs is None else
        {"per_gpu": h2d_gbs, "aggregate": h2d_gbs * ctx.world, "note": "1 GiB pinned copies, all ranks at once"},
        "configs": [dict(finish(k, r), name=k) for k, r in others],
    }
    print(json.dumps(line))


if __name__ == "__main__":
    main()
