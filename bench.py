#!/usr/bin/env python
"""Benchmark of the intrinsic-reward hot path (BASELINE.json metric): samples/s of
bonus + update on an RND [T,B] rollout batch, T=128, B=1024 per GPU, 4x84x84 uint8 frames.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--B 1024] [--T 128]

One "step" = RndStep.__call__: compute_bonus -> reward += r_int -> GAE/valid -> compute_loss
forward+backward over all T*B samples -> (all-reduce) -> clip -> Adam.  Prints ONE JSON line on
rank 0.  `--impl reference` times the CPU restatement of the reference (oracle/, torch CPU ops on
all host cores) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

# algorithmic FLOPs per sample, RND bonus + one-epoch update: 2*fwd + bwd
FWD_MAC_TARGET = 64 * 32 * 400 + 512 * 64 * 81 + 576 * 64 * 49 + 3136 * 512
FWD_MAC_PRED = FWD_MAC_TARGET + 2 * 512 * 512
FLOP_PER_SAMPLE = 2 * (2 * (FWD_MAC_TARGET + FWD_MAC_PRED)                       # bonus fwd + loss fwd
                       + 2 * FWD_MAC_PRED - 64 * 32 * 400)                       # bwd: wgrad all + dgrad (no conv1 dgrad)


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


# ------------------------------------------------------------------------------ reference arm
def cpu_reference(T, B, steps, warmup, seconds_cap=120.0):
    """The reference's CPU path for the same unit of work, restated in oracle/ (torch CPU fp32,
    all host threads): bonus -> reward/GAE -> loss.backward -> clip_grad_norm_ -> Adam.step."""
    import recipes as R
    from oracle import curiosity_oracle as O
    torch.set_num_threads(os.cpu_count())
    params = R.make_params(R.rnd_shapes(), 1, gain=1.4)
    p = {k: v.clone().requires_grad_(k.startswith("forward")) for k, v in params.items()}
    train = [v for k, v in p.items() if k.startswith("forward")]
    opt = torch.optim.Adam(train, lr=1e-4)
    m = O.RndOracle(drop_probability=0.25)
    frames = R.make_frames(2, T, B)
    done = R.make_suffix_done(3, T, B, frac=0.01)
    value, boot = R.make_normal(4, T, B), R.make_normal(5, B)
    mask = R.make_mask(6, (T, B), 0.25)

    def step():
        r_int = m.compute_bonus(p, frames, done)
        adv, ret = O.generalized_advantage_estimation(r_int.clone(), value, done, boot, 0.99, 0.95)
        valid = O.valid_from_done(done)
        opt.zero_grad()
        loss = m.compute_loss(p, frames, valid, mask)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(train, 1.0)
        opt.step()
        return float(loss)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    done_steps = 0
    for _ in range(steps):
        step()
        done_steps += 1
        if time.perf_counter() - t0 > seconds_cap:
            break
    dt = (time.perf_counter() - t0) / done_steps
    return T * B / dt, dt * 1e3, done_steps


# ------------------------------------------------------------------------------ ICM / Disagreement
# algorithmic MFLOP per sample, bonus + one-epoch update = 2*fwd + bwd (Burda encoder)
ENC_MAC = 256 * 32 * 400 + 512 * 64 * 81 + 576 * 64 * 49 + 3136 * 512          # 4-channel conv1


def icm_flops(A, E):
    F = 512
    resf = 10 * (F + A) * F
    inv = 2 * F * F + F * A
    fwd = 2 * ENC_MAC + inv + E * resf
    # backward: wgrad of every layer; dgrad of every layer whose input needs a gradient (no conv1 dgrad,
    # no dgrad into the detached features entering the forward models' first layer)
    enc_b = 2 * (2 * ENC_MAC - 256 * 32 * 400)
    inv_b = 2 * inv
    resf_b = E * (2 * resf - (F + A) * F)
    return 2 * (2 * fwd + enc_b + inv_b + resf_b)


def other_model(args, rank, world, local):
    import torch.distributed as dist
    from curiosity_baselines_b200 import _lib as L
    from curiosity_baselines_b200 import layers as LY
    from curiosity_baselines_b200.curiosity.icm import ICM, Disagreement
    from curiosity_baselines_b200.trainer import IcmStep
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    LY.set_engine(args.engine)
    T, B = args.T, args.B
    A = args.A or (4 if args.model == "icm" else 7)
    E = 1 if args.model == "icm" else 4
    torch.manual_seed(0)
    cls = ICM if args.model == "icm" else Disagreement
    model = cls(image_shape=(4, 84, 84), action_size=A, feature_encoding="idf_burda").to(dev)
    step = IcmStep(model, lr=1e-4)
    g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
    obs = torch.randint(0, 256, (T, B, 4, 84, 84), dtype=torch.uint8, device=dev, generator=g)
    nxt = torch.randint(0, 256, (T, B, 4, 84, 84), dtype=torch.uint8, device=dev, generator=g)
    act = torch.nn.functional.one_hot(torch.randint(0, A, (T, B), device=dev, generator=g), A).float()
    first = torch.randint(1, T + 1, (B,), device=dev, generator=g)
    has = torch.rand(B, device=dev, generator=g) < 0.01
    first = torch.where(has, first, torch.full_like(first, T))
    done = (torch.arange(T, device=dev).unsqueeze(1) >= first.unsqueeze(0)).float()
    value = torch.randn(T, B, device=dev, generator=g)
    boot = torch.randn(B, device=dev, generator=g)

    def one_step():
        return step(obs, nxt, act, done, torch.zeros(T, B, device=dev), value, boot)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        one_step()
    barrier()
    l0 = L.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        loss, r_int, adv, ret = one_step()
    ev1.record()
    barrier()
    launches = L.launch_count() - l0
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    # the same step with reuse_features (the reference passes the same observation / next_observation / action to
    # curiosity_step and curiosity_loss): reported separately, the value above recomputes everything
    reuse = None
    try:
        model.reuse_features = True
        for _ in range(2):
            one_step()
        barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0 = model.reuse_hits["full"]
        r0.record()
        for _ in range(args.steps):
            one_step()
        r1.record()
        barrier()
        rms = torch.tensor([r0.elapsed_time(r1)], device=dev)
        if world > 1:
            dist.all_reduce(rms, op=dist.ReduceOp.MAX)
        reuse = {"ms_per_step": float(rms) / args.steps, "value": world * T * B / (float(rms) / args.steps * 1e-3),
                 "unit": "samples/s", "full_hits": model.reuse_hits["full"] - h0,
                 "note": "loss pass reuses the bonus pass's forward (exact; guarded by cb_frames_differ); not the headline value"}
    finally:
        model.reuse_features = False
        model._cache = None
    hbm, tf_burst, tf_sust, src = peaks()
    flop = icm_flops(A, E)
    if rank == 0:
        tfl = flop * T * B / (ms_per_step * 1e-3) / 1e12
        print(json.dumps({
            "metric": "intrinsic-reward samples/sec (bonus+update)", "value": world * T * B / (ms_per_step * 1e-3),
            "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{args.model} bonus+update, T={T}, B={B}/GPU, 4x84x84 uint8, A={A}, E={E} forward models",
                       "global_B": B * world, "l2": "inputs (7.4 GB frames/GPU) larger than L2", "engine": args.engine,
                       "parallelism": f"dp{world} over B"},
            "clocks": clocks, "gpu_launches": launches, "e2e": None,
            "roofline": {"bound": "tensor", "kernel": "whole step", "achieved": tfl, "peak": tf_sust, "unit": "TFLOP/s",
                         "frac": tfl / tf_sust, "traffic": None, "peak_source": f"{src} bf16 sustained",
                         "mflop_per_sample": flop / 1e6},
            "cpu_baseline": None, "with_feature_reuse": reuse, "loss": float(loss)}))
    if world > 1:
        dist.destroy_process_group()


def ndigo_model(args, rank, world, local):
    """BASELINE configs[4]: NDIGO on PyColab-shaped 3x5x5 binary frames, A=5, horizon 3, 10 predictors -- the
    small-tensor regime (0.8 MFLOP per sample): bonus (no grad) + loss forward/backward + Adam.  Launch- and
    bandwidth-bound: this line is a measurement, not a tensor-roofline claim."""
    from curiosity_baselines_b200 import _lib as L
    from curiosity_baselines_b200.curiosity.ndigo import NDIGO
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    T, B, A = args.T, args.B, args.A or 5
    torch.manual_seed(0)
    model = NDIGO(image_shape=(3, 5, 5), action_size=A, horizon=3, feature_encoding="idf_maze", num_predictors=10).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
    obs = (torch.rand(T, B, 3, 5, 5, device=dev, generator=g) < 0.3).float()
    one = lambda: torch.nn.functional.one_hot(torch.randint(0, A, (T, B), device=dev, generator=g), A).float()
    prev, act, valid = one(), one(), torch.ones(T, B, device=dev)

    def one_step():
        r_int = model.compute_bonus(obs, prev, act)
        opt.zero_grad(set_to_none=True)
        loss = model.compute_loss(obs, prev, act, valid)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        return loss, r_int

    for _ in range(args.warmup):
        one_step()
    torch.cuda.synchronize()
    l0 = L.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        loss, r_int = one_step()
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / args.steps
    if rank == 0:
        print(json.dumps({
            "metric": "intrinsic-reward samples/sec (bonus+update)", "value": T * B / (ms * 1e-3), "unit": "samples/s",
            "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"ndigo bonus+update, T={T}, B={B}, 3x5x5 binary frames, A={A}, horizon 3, 10 predictors",
                       "note": "encoder / GRU (cb_gru_sequence_*) / predictors / BCE all in libcuriosity_b200"},
            "gpu_launches": L.launch_count() - l0, "e2e": None, "roofline": None, "cpu_baseline": None, "loss": float(loss)}))


# ------------------------------------------------------------------------------ B200 arm
# ------------------------------------------------------------------------------ ICM + LSTM PPO (BASELINE configs[3])
def cpu_reference_ppo(T, B, A, epochs, steps, warmup, seconds_cap=60.0):
    """The reference's CPU path of one PPO iteration (ICM bonus -> returns -> ``epochs`` x [policy forward, PPO
    loss, ICM loss, backward, clip_grad_norm_, Adam]) restated in oracle/ (torch CPU fp32, all host threads)."""
    import recipes as R
    from oracle import curiosity_oracle as O
    torch.set_num_threads(os.cpu_count())
    p = {k: v.clone().requires_grad_(True) for k, v in R.make_policy_params(A, 1).items()}
    q = {k: v.clone().requires_grad_(True) for k, v in R.make_params(R.icm_shapes(A, "idf_burda"), 2).items()}
    allp = list(p.values()) + list(q.values())
    opt = torch.optim.Adam(allp, lr=1e-4)
    obs, nxt = R.make_frames(3, T, B), R.make_frames(4, T, B)
    prev, act = R.make_actions(5, T, B, A), R.make_actions(6, T, B, A)
    old_prob = R.make_probs(7, T, B, A)
    done = R.make_suffix_done(8, T, B, frac=0.01)
    value, boot = R.make_normal(9, T, B), R.make_normal(10, B)
    h0 = (R.make_normal(11, 1, B, 512) * 0.1, R.make_normal(12, 1, B, 512) * 0.1)

    def step():
        with torch.no_grad():
            r_int = O.icm_bonus(q, obs, nxt, R.onehot(act, A), "idf_burda", 1.0)
        adv, ret = O.generalized_advantage_estimation(r_int.clone(), value, done, boot, 0.99, 0.95)
        valid = O.valid_from_done(done)
        for _ in range(epochs):
            opt.zero_grad()
            pi, v, _ = O.policy_forward(p, obs, R.onehot(prev, A), h0)
            loss = O.ppo_loss(pi, v, act, old_prob, ret, adv, valid, 0.1, 1.0, 0.01)[0]
            inv, fwd = O.icm_loss(q, obs, nxt, R.onehot(act, A), valid, "idf_burda", 0.2)
            (loss + inv + fwd).backward()
            torch.nn.utils.clip_grad_norm_(allp, 1.0)
            opt.step()

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    k = 0
    for _ in range(steps):
        step()
        k += 1
        if time.perf_counter() - t0 > seconds_cap:
            break
    dt = (time.perf_counter() - t0) / k
    return T * B / dt, dt * 1e3, k


def ppo_flops(A, epochs):
    """Algorithmic FLOP per sample of one PPO iteration (SURVEY 8d ii): ICM bonus forward + ``epochs`` x (policy
    forward/backward + ICM forward/backward).  Policy: Burda encoder + LSTM(512+A -> 512) + pi/value heads."""
    F, H = 512, 512
    lstm = (F + A) * 4 * H + H * 4 * H
    heads = H * (A + 1)
    pol_fwd = ENC_MAC + lstm + heads
    # backward: encoder wgrad + dgrad (no conv1 dgrad); LSTM: wgrad of both matrices, dgrad to the features and to h
    pol_bwd = (2 * ENC_MAC - 256 * 32 * 400) + 2 * lstm - A * 4 * H + 2 * heads
    icm_unit = icm_flops(A, 1) // 2                    # MAC of 2*fwd + bwd
    resf = 10 * (F + A) * F
    inv = 2 * F * F + F * A
    icm_fwd = 2 * ENC_MAC + inv + resf
    icm_bwd = icm_unit - 2 * icm_fwd
    return 2 * (icm_fwd + epochs * (pol_fwd + pol_bwd + icm_fwd + icm_bwd))


def ppo_model(args, rank, world, local):
    import numpy as np
    import torch.distributed as dist
    from curiosity_baselines_b200 import _lib as L
    from curiosity_baselines_b200 import layers as LY
    from curiosity_baselines_b200.algos.ppo import PPO, PpoBatch
    from curiosity_baselines_b200.pg.atari_lstm_model import AtariLstmModel
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    LY.set_engine(args.engine)
    T, B = args.T, args.B
    A = args.A or 4
    epochs, minibatches = args.epochs, args.minibatches
    torch.manual_seed(0)
    model = AtariLstmModel((4, 84, 84), A, curiosity_kwargs=dict(
        curiosity_alg="icm", feature_encoding="idf_burda", batch_norm=False, prediction_beta=1.0,
        forward_loss_wt=0.2)).to(dev)
    with torch.no_grad():       # damp the policy encoder so raw 0..255 pixels give O(1) features (non-saturated gates)
        for name, prm in model.conv.named_parameters():
            prm.mul_(0.65)
    algo = PPO(discount=0.99, learning_rate=1e-4, gae_lambda=0.95, minibatches=minibatches, epochs=epochs,
               curiosity_type="icm", linear_lr_schedule=False)
    algo.initialize(model, n_itr=10 ** 6, rng=np.random.RandomState(rank))
    g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
    obs = torch.randint(0, 256, (T, B, 4, 84, 84), dtype=torch.uint8, device=dev, generator=g)
    nxt = torch.randint(0, 256, (T, B, 4, 84, 84), dtype=torch.uint8, device=dev, generator=g)
    act = torch.randint(0, A, (T, B), device=dev, generator=g)
    prev = torch.randint(0, A, (T, B), device=dev, generator=g)
    old_prob = torch.softmax(torch.randn(T, B, A, device=dev, generator=g), -1)
    first = torch.randint(1, T + 1, (B,), device=dev, generator=g)
    has = torch.rand(B, device=dev, generator=g) < 0.01
    first = torch.where(has, first, torch.full_like(first, T))
    done = (torch.arange(T, device=dev).unsqueeze(1) >= first.unsqueeze(0)).float()
    value = torch.randn(T, B, device=dev, generator=g)
    boot = torch.randn(B, device=dev, generator=g)
    h0 = torch.randn(B, 1, 512, device=dev, generator=g) * 0.1
    c0 = torch.randn(B, 1, 512, device=dev, generator=g) * 0.1

    def one_step():
        batch = PpoBatch(obs, nxt, prev, act, old_prob, torch.zeros(T, B, device=dev), done, value, boot, (h0, c0))
        return algo.optimize_agent(0, batch)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        one_step()
    barrier()
    l0 = L.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        infos = one_step()
    ev1.record()
    barrier()
    launches = L.launch_count() - l0
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    # the policy pass alone (forward + PPO loss + backward), for the breakdown
    ret, adv, valid = algo.process_returns(PpoBatch(obs, nxt, prev, act, old_prob, torch.zeros(T, B, device=dev), done,
                                                    value, boot, (h0, c0)))
    algo_p = PPO(curiosity_type="none")
    algo_p.model = model
    for _ in range(2):
        algo_p.loss_and_grads_(obs, prev, act, ret, adv, valid, old_prob, (h0, c0))
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(3):
        algo_p.loss_and_grads_(obs, prev, act, ret, adv, valid, old_prob, (h0, c0))
    p1.record()
    barrier()
    policy_ms = p0.elapsed_time(p1) / 3
    # the same iteration with ICM.reuse_features: the first epoch's ICM forward is the bonus pass's (same inputs, same
    # weights; guarded by an equality check of the frames) -- reported separately, the value above recomputes it
    reuse = None
    cm = model.curiosity_model
    try:
        cm.reuse_features = True
        one_step()
        barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        hits0 = cm.reuse_hits["full"]
        r0.record()
        for _ in range(args.steps):
            one_step()
        r1.record()
        barrier()
        rms = torch.tensor([r0.elapsed_time(r1)], device=dev)
        if world > 1:
            dist.all_reduce(rms, op=dist.ReduceOp.MAX)
        reuse = {"ms_per_step": float(rms) / args.steps, "value": world * T * B / (float(rms) / args.steps * 1e-3),
                 "unit": "samples/s", "full_hits": cm.reuse_hits["full"] - hits0,
                 "note": "first-epoch ICM forward reused from the bonus pass (exact; guarded); not the headline value"}
    finally:
        cm.reuse_features = False
        cm._cache = None
    # end to end: every iteration uploads its batch from pinned host memory (both uint8 frame tensors + the small
    # rollout tensors) on the compute stream and reads the loss back
    e2e = None
    if not args.no_e2e:
        host = [t.cpu().pin_memory() for t in (obs, nxt, prev, act, old_prob, done, value, boot, h0, c0)]
        h2d = sum(t.numel() * t.element_size() for t in host)

        from curiosity_baselines_b200.agents.hooks import BatchPrefetcher
        pre = BatchPrefetcher(dev)
        pre.prefetch(host)

        def e2e_step():
            d = pre.get()
            pre.prefetch(host)               # the next batch's upload runs under this iteration's compute
            batch = PpoBatch(d[0], d[1], d[2], d[3], d[4], torch.zeros(T, B, device=dev), d[5], d[6], d[7], (d[8], d[9]))
            out = algo.optimize_agent(0, batch)
            return float(out[-1].loss)

        e2e_step()
        barrier()
        k = max(2, args.steps // 2)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            e2e_step()
        e1.record()
        barrier()
        ems = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        e2e = {"value": world * T * B / (float(ems) / k * 1e-3), "unit": "samples/s", "ms_per_step": float(ems) / k,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
               "note": "every iteration's batch (14.8 GB of uint8 frames + the small rollout tensors) is uploaded from "
                       "pinned host memory on a side stream under the previous iteration's compute (BatchPrefetcher)"}
        pre.get()
        del host, pre
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        v, cms, k = cpu_reference_ppo(T, args.cpu_B, A, epochs, 3, 1)
        cpu = {"value": v, "unit": "samples/s", "cores": os.cpu_count(), "cpu_model": cpu_model(), "kind": "port",
               "sample": f"T={T}, B={args.cpu_B} ({T * args.cpu_B} samples/iteration) of the same workload, {k} iterations, "
                         f"{cms:.0f} ms each"}
    hbm, tf_burst, tf_sust, src = peaks()
    flop = ppo_flops(A, epochs)
    if rank == 0:
        tfl = flop * T * B / (ms_per_step * 1e-3) / 1e12
        print(json.dumps({
            "metric": "intrinsic-reward samples/sec (bonus+update)", "value": world * T * B / (ms_per_step * 1e-3),
            "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"ICM + LSTM PPO full iteration (BASELINE configs[3]): bonus + returns + {epochs} epochs x "
                                   f"{minibatches} minibatches of policy+ICM loss/backward/clip/Adam, T={T}, B={B}/GPU, "
                                   f"4x84x84 uint8, A={A}",
                       "global_B": B * world, "l2": "inputs (14.8 GB frames/GPU) larger than L2", "engine": args.engine,
                       "parallelism": f"dp{world} over B"},
            "clocks": clocks, "gpu_launches": launches, "e2e": e2e,
            "roofline": {"bound": "tensor", "kernel": "whole iteration", "achieved": tfl, "peak": tf_sust,
                         "unit": "TFLOP/s", "frac": tfl / tf_sust, "traffic": None,
                         "peak_source": f"{src} bf16 sustained", "mflop_per_sample": flop / 1e6},
            "cpu_baseline": cpu, "with_feature_reuse": reuse, "policy_fwd_loss_bwd_ms": policy_ms,
            "loss": float(infos[-1].loss), "grad_norm": float(infos[-1].gradNorm)}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--T", type=int, default=128)
    ap.add_argument("--B", type=int, default=None, help="env columns per GPU (weak scaling); default 1024 "
                                                       "(2048 for ppo_icm, BASELINE configs[3])")
    ap.add_argument("--epochs", type=int, default=3, help="ppo_icm: PPO epochs per iteration (launcher default 3)")
    ap.add_argument("--minibatches", type=int, default=1, help="ppo_icm: minibatches per epoch (launcher default 1)")
    ap.add_argument("--engine", default="auto")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="ppo_icm: skip the host-buffer end-to-end leg")
    ap.add_argument("--model", default="rnd", choices=["rnd", "icm", "disagreement", "ndigo", "ppo_icm"],
                    help="rnd = BASELINE configs[1] (the headline line); icm / disagreement: the same unit of work "
                         "for the other two conv curiosity models (device-resident inputs only)")
    ap.add_argument("--A", type=int, default=None, help="number of discrete actions (icm: 4, disagreement: 7)")
    ap.add_argument("--cpu-B", type=int, default=8)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.B is None:
        args.B = 2048 if args.model == "ppo_icm" else 1024
    T, B = args.T, args.B
    workload = f"RND bonus+update, T={T}, B={B}/GPU, 4x84x84 uint8 (BASELINE configs[1])"

    if args.impl == "reference":
        if rank != 0:
            return
        if args.model == "ppo_icm":
            v, ms, k = cpu_reference_ppo(T, args.cpu_B, args.A or 4, args.epochs, args.steps, args.warmup, seconds_cap=120.0)
            workload = (f"ICM + LSTM PPO full iteration (BASELINE configs[3]), T={T}, B={B}/GPU, 4x84x84 uint8, "
                        f"A={args.A or 4}, {args.epochs} epochs")
        else:
            v, ms, k = cpu_reference(T, args.cpu_B, args.steps, args.warmup)
        sample = f"T={T}, B={args.cpu_B} ({T * args.cpu_B} samples/step) of the same workload, {k} steps"
        print(json.dumps({
            "impl": "reference", "metric": "intrinsic-reward samples/sec (bonus+update)", "value": v,
            "unit": "samples/s", "n_gpus": args.gpus, "steps": k, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload},
            "cpu_baseline": {"value": v, "unit": "samples/s", "cores": os.cpu_count(), "cpu_model": cpu_model(),
                             "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    if args.model == "ndigo":
        return ndigo_model(args, rank, world, local)
    if args.model == "ppo_icm":
        return ppo_model(args, rank, world, local)
    if args.model != "rnd":
        return other_model(args, rank, world, local)
    import torch.distributed as dist
    from curiosity_baselines_b200 import _lib as L
    from curiosity_baselines_b200 import layers as LY
    from curiosity_baselines_b200.curiosity.rnd import RND
    from curiosity_baselines_b200.trainer import RndStep
    from curiosity_baselines_b200.agents.hooks import newest_frame_to_device

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = None
    if world > 1:
        from curiosity_baselines_b200.dist import bind_to_gpu_numa_node
        numa = bind_to_gpu_numa_node(local)    # host staging buffers local to the GPU's socket
        dist.init_process_group("nccl", device_id=dev)
    LY.set_engine(args.engine)

    torch.manual_seed(0)                       # identical weights on every rank
    model = RND(image_shape=(4, 84, 84), prediction_beta=1.0, drop_probability=0.25, gamma=0.99).to(dev)
    step = RndStep(model, lr=1e-4)
    g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
    # synthetic rollout, device resident (3.7 GB of frames at the default size: larger than L2)
    frames = torch.randint(0, 256, (T, B, 4, 84, 84), dtype=torch.uint8, device=dev, generator=g)
    first = torch.randint(1, T + 1, (B,), device=dev, generator=g)
    has = torch.rand(B, device=dev, generator=g) < 0.01
    first = torch.where(has, first, torch.full_like(first, T))
    done = (torch.arange(T, device=dev).unsqueeze(1) >= first.unsqueeze(0)).float()
    value = torch.randn(T, B, device=dev, generator=g)
    boot = torch.randn(B, device=dev, generator=g)
    mask = (torch.rand(T, B, device=dev, generator=g) > 0.25).float()
    reward0 = torch.zeros(T, B, device=dev)     # no_extrinsic

    def one_step():
        return step(frames, done, reward0.clone(), value, boot, mask)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                      # sampled under load: warm-up + timed region
    # at least 8 untimed steps: a fresh process finds the GPU at idle clocks, and one outlier run measured the
    # first ~0.5 s of work 15 % slow while they ramped under the power cap
    args.warmup = max(args.warmup, 8)
    for _ in range(args.warmup):
        one_step()
    barrier()
    l0 = L.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        loss, r_int, adv, ret = one_step()
    ev1.record()
    barrier()
    launches = L.launch_count() - l0
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    value_sps = world * T * B / (ms_per_step * 1e-3)

    # ---- end to end through the host-facing call: pinned host buffers in, results out
    h_frames = torch.empty((T, B, 4, 84, 84), dtype=torch.uint8).pin_memory()
    h_frames.copy_(frames)
    h_small = [t.cpu().pin_memory() for t in (done, value, boot, mask)]
    h_out = [torch.empty(T, B).pin_memory() for _ in range(3)]
    h_loss = torch.empty(1).pin_memory()

    from curiosity_baselines_b200.agents.hooks import FrameStager
    stager = FrameStager(dev)

    def e2e_step():
        """Host buffers in, host results out.  The strided upload of the NEXT step's frames runs on a
        side stream under this step's compute (double buffered), as a training loop would do it."""
        d_frames = stager.get()
        d_done, d_value, d_boot, d_mask = stager.get_extras()
        loss, r_int, adv, ret = step(d_frames, d_done, torch.zeros(T, B, device=dev), d_value, d_boot, d_mask)
        stager.done()
        h_out[0].copy_(r_int, non_blocking=True); h_out[1].copy_(adv, non_blocking=True)
        h_out[2].copy_(ret, non_blocking=True); h_loss.copy_(loss.reshape(1), non_blocking=True)
        stager.prefetch(h_frames, h_small)

    stager.prefetch(h_frames, h_small)
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    n_e2e = max(3, min(args.steps, 10))
    for _ in range(n_e2e):
        e2e_step()
    barrier()
    e2e_ms = torch.tensor([(time.perf_counter() - t0) * 1e3 / n_e2e], device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    h2d = T * B * 84 * 84 + sum(t.numel() * 4 for t in h_small)
    d2h = 3 * T * B * 4 + 4

    # ---- the same step with RND.reuse_features (NOT the headline value): the reference passes the same
    # next_observation to compute_bonus and compute_loss (base.py:72-73, ppo.py:107-110), so the loss pass can reuse
    # the bonus pass's features after a device-side equality check of the frames.  Reported separately; `value`
    # above recomputes everything.
    reuse = None
    try:
        model.reuse_features = True
        for _ in range(3):
            one_step()
        barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        hits0 = dict(model.reuse_hits)
        r0.record()
        for _ in range(args.steps):
            one_step()
        r1.record()
        barrier()
        rms = torch.tensor([r0.elapsed_time(r1)], device=dev)
        if world > 1:
            dist.all_reduce(rms, op=dist.ReduceOp.MAX)
        rms_step = float(rms) / args.steps
        reuse = {"ms_per_step": rms_step, "value": world * T * B / (rms_step * 1e-3), "unit": "samples/s",
                 "full_hits": model.reuse_hits["full"] - hits0["full"], "steps": args.steps,
                 "note": "loss pass reuses the bonus pass's target+predictor features of the same frames "
                         "(exact; guarded by cb_frames_differ); not the headline value"}
    finally:
        model.reuse_features = False
        model._cache = None

    # ---- roofline of the dominant kernel: predictor conv2 forward (implicit GEMM 512x64 per position)
    hbm, tf_burst, tf_sust, src = peaks()
    layer = model._chains[0].layers[1]
    n = T * B
    a1 = torch.randn(n * 400, 32, device=dev).to(torch.bfloat16)
    for _ in range(3):
        layer.forward(a1, n)
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    k0.record()
    for _ in range(reps):
        layer.forward(a1, n)
    k1.record()
    torch.cuda.synchronize()
    kern_ms = k0.elapsed_time(k1) / reps
    kern_tflops = 2.0 * 512 * 64 * 81 * n / (kern_ms * 1e-3) / 1e12
    # algorithmic bytes of this kernel: a1 read once (400 positions x 32 ch bf16) + a2 written once (81 x 64 bf16)
    alg_bytes = float(n) * (400 * 64 + 81 * 128)
    kern_gbs = alg_bytes / (kern_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r1_roofline_traffic.json")
    if os.path.exists(tpath) and n == 131072:      # dram__bytes_read+write of this kernel at this size (ncu --set full)
        traffic = json.load(open(tpath))["conv2_fwd_window_n131072"]["dram_bytes_per_launch"]
    # 148 FLOP/byte is below the ridge (~206): the binding roofline of this layer is HBM, not the tensor pipe
    roofline = {"bound": "hbm", "kernel": "conv2 forward (32->64, k4 s2) window implicit GEMM (umma_window_conv_kernel<64>)",
                "achieved": kern_gbs, "peak": hbm, "unit": "GB/s", "frac": kern_gbs / hbm,
                "traffic": traffic, "peak_source": f"{src} HBM copy bandwidth", "kernel_ms": kern_ms,
                "algorithmic_bytes": alg_bytes, "tensor_tflops": kern_tflops, "tensor_frac_of_burst": kern_tflops / tf_burst,
                "step_tflops": FLOP_PER_SAMPLE * T * B / (ms_per_step * 1e-3) / 1e12,
                "step_frac_of_sustained": FLOP_PER_SAMPLE * T * B / (ms_per_step * 1e-3) / 1e12 / tf_sust}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if not args.no_cpu_baseline:
        v, ms_cpu, k = cpu_reference(T, args.cpu_B, 6, 1, seconds_cap=25.0)
        cpu = {"value": v, "unit": "samples/s", "cores": os.cpu_count(), "cpu_model": cpu_model(), "kind": "port",
               "sample": f"T={T}, B={args.cpu_B} ({T * args.cpu_B} samples/step), {k} steps, {ms_cpu:.0f} ms/step"}
    print(json.dumps({
        "metric": "intrinsic-reward samples/sec (bonus+update)", "value": value_sps, "unit": "samples/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload, "global_B": B * world, "l2": "inputs (3.7 GB frames/GPU) larger than L2",
                   "engine": args.engine, "parallelism": f"dp{world} over B", "numa_node_rank0": numa},
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": world * T * B / (float(e2e_ms) * 1e-3), "unit": "samples/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": float(e2e_ms)},
        "roofline": roofline, "cpu_baseline": cpu, "with_feature_reuse": reuse, "loss": float(loss)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
