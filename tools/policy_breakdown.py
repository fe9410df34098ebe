"""Where the policy pass spends its time: CUDA-event time of each stage + the CPU time spent issuing it.
    python tools/policy_breakdown.py [--B 2048] [--T 128]"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from curiosity_baselines_b200 import _lib as L                                   # noqa: E402
from curiosity_baselines_b200.layers import pad_side                              # noqa: E402
from curiosity_baselines_b200.pg.atari_lstm_model import AtariLstmModel           # noqa: E402


class Timer:
    def __init__(self):
        self.rows = []

    def section(self, name):
        t = self

        class S:
            def __enter__(self):
                torch.cuda.synchronize()
                self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                self.t0 = time.perf_counter()
                self.l0 = L.launch_count()
                self.e0.record()

            def __exit__(self, *a):
                self.e1.record()
                cpu = (time.perf_counter() - self.t0) * 1e3
                torch.cuda.synchronize()
                t.rows.append((name, self.e0.elapsed_time(self.e1), cpu, L.launch_count() - self.l0))
        return S()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=2048)
    ap.add_argument("--T", type=int, default=128)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    T, B, A, H = a.T, a.B, 4, 512
    n = T * B
    m = AtariLstmModel((4, 84, 84), A, curiosity_kwargs=dict(curiosity_alg="icm", feature_encoding="idf_burda",
                                                             batch_norm=False, prediction_beta=1.0,
                                                             forward_loss_wt=0.2)).to(dev)
    with torch.no_grad():
        for p in m.conv.parameters():
            p.mul_(0.65)
    g = torch.Generator(device=dev); g.manual_seed(1)
    obs = torch.randint(0, 256, (T, B, 4, 84, 84), dtype=torch.uint8, device=dev, generator=g)
    prev = torch.nn.functional.one_hot(torch.randint(0, A, (T, B), device=dev, generator=g), A).float()
    h0 = torch.randn(1, B, H, device=dev, generator=g) * 0.1
    for rep in range(2):
        tm = Timer()
        with tm.section("forward (all)"):
            s = m._run_forward(obs, prev, (h0, h0), save=True)
        dlogits = torch.randn(n, A, device=dev) * 1e-4
        dvalue = torch.randn(n, device=dev) * 1e-4
        with tm.section("backward (all)"):
            m._run_backward(s, dlogits, dvalue)
        ex = m._exec
        # stage by stage
        from curiosity_baselines_b200.curiosity.encoders import frame_input
        keep, x, dtype, strides = frame_input(obs.reshape(n, 4, 84, 84))
        with tm.section("encoder fwd"):
            acts, _ = ex["chain"].forward(x, n, in_dtype=dtype, strides=strides, last_f32=False, last_bf16=True)
        feat = acts[-1].reshape(n, -1)
        side = prev.reshape(n, A).contiguous()
        sp = pad_side(side)
        with tm.section("x W_ih^T (one GEMM)"):
            _, gates = ex["ih"].forward(feat, n, side=side, side_pad=sp, out_bf16=False, out_f32=True)
        gates = gates.view(T, B, 4 * H)
        h_all = torch.zeros(T + 1, B, H, dtype=torch.bfloat16, device=dev)
        c_all = torch.zeros(T + 1, B, H, dtype=torch.float32, device=dev)
        with tm.section("time loop fwd (T x GEMM+cell)"):
            m._lstm_forward(gates, h_all, c_all, None)
        h_out = h_all[1:].reshape(n, H)
        logits, value, prob = torch.empty(n, A, device=dev), torch.empty(n, device=dev), torch.empty(n, A, device=dev)
        with tm.section("heads fwd (fused pi+value+softmax)"):
            L.call("cb_policy_heads_fwd", L.ptr(h_out), L.ptr(m.pi.weight), L.ptr(m.pi.bias), L.ptr(m.value.weight),
                   L.ptr(m.value.bias), L.ptr(logits), L.ptr(value), L.ptr(prob), n, H, A, L.stream())
        with tm.section("heads bwd (fused dh+dW)"):
            dh_out = m._heads_backward(h_out, dlogits, dvalue, n).view(T, B, H)
        dgates = torch.empty(T, B, 4 * H, dtype=torch.bfloat16, device=dev)
        with tm.section("time loop bwd (T x cell+dgrad)"):
            m._lstm_backward(gates, h_all, c_all, dh_out, dgates)
        dg = dgates.view(n, 4 * H)
        with tm.section("W_hh wgrad"):
            ex["hh"].wgrad(h_all[:T].reshape(n, H), dg, n)
        with tm.section("W_ih wgrad"):
            ex["ih"].wgrad(feat, dg, n, side=side)
        with tm.section("W_ih dgrad"):
            dfeat = ex["ih"].dgrad(dg, n, x_saved=feat, x_act="none")
        with tm.section("encoder bwd"):
            ex["chain"].backward(x, acts, dfeat, n, in_dtype=dtype, strides=strides)
        if rep == 1:
            print(f"policy pass breakdown, T={T}, B={B} (n={n})")
            print(f"{'stage':36s} {'gpu ms':>9s} {'cpu issue ms':>13s} {'lib calls':>10s}")
            for name, gpu, cpu, calls in tm.rows:
                print(f"{name:36s} {gpu:9.3f} {cpu:13.3f} {calls:10d}")


if __name__ == "__main__":
    main()
