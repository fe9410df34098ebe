"""Event-timed micro-benchmark of single layers through the library (used for profiling).
    python tools/layer_bench.py [--n 131072] [--reps 5] [--only conv2]"""
import argparse
import os
import sys

import torch
from torch import nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from curiosity_baselines_b200 import layers as LY  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=131072)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--only", default="")
ap.add_argument("--engine", default="auto")
ap.add_argument("--warmup", type=int, default=2)
args = ap.parse_args()
dev = "cuda"
n = args.n
LY.set_engine(args.engine)
torch.manual_seed(0)


def timeit(fn, flops, name):
    for _ in range(args.warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    print(f"{name:28s} {ms:8.3f} ms  {flops / ms / 1e9:8.1f} TFLOP/s", flush=True)


cases = {}
conv2 = LY.conv_layer(nn.Conv2d(32, 64, 4, 2).to(dev), (20, 20, 32), "lrelu")
conv3 = LY.conv_layer(nn.Conv2d(64, 64, 3, 1).to(dev), (9, 9, 64), "lrelu")
fc = LY.linear_layer(nn.Linear(3136, 512).to(dev), "relu", spatial=(7, 64))
lin = LY.linear_layer(nn.Linear(512, 512).to(dev), "relu")
a1 = torch.randn(n * 400, 32, device=dev).to(torch.bfloat16)
a2 = torch.randn(n * 81, 64, device=dev).to(torch.bfloat16)
a3 = torch.randn(n * 49, 64, device=dev).to(torch.bfloat16)
h = torch.randn(n, 512, device=dev).to(torch.bfloat16)
cases["conv2_fwd"] = (lambda: conv2.forward(a1, n), 2.0 * 512 * 64 * 81 * n)
cases["conv3_fwd"] = (lambda: conv3.forward(a2, n), 2.0 * 576 * 64 * 49 * n)
cases["fc_fwd"] = (lambda: fc.forward(a3, n), 2.0 * 3136 * 512 * n)
cases["lin_fwd"] = (lambda: lin.forward(h, n), 2.0 * 512 * 512 * n)
cases["lin_fwd_f32"] = (lambda: lin.forward(h, n, out_bf16=False, out_f32=True), 2.0 * 512 * 512 * n)
cases["lin_dgrad"] = (lambda: lin.dgrad(h, n, x_saved=h, x_act="relu"), 2.0 * 512 * 512 * n)
cases["conv3_dgrad"] = (lambda: conv3.dgrad(a3, n, x_saved=a2, x_act="lrelu"), 2.0 * 576 * 64 * 81 * n)
dy2 = torch.randn(n * 81, 64, device=dev).to(torch.bfloat16)
cases["conv2_dgrad"] = (lambda: conv2.dgrad(dy2, n, x_saved=a1, x_act="lrelu"), 2.0 * 512 * 64 * 81 * n)
cases["fc_dgrad"] = (lambda: fc.dgrad(h, n, x_saved=a3, x_act="lrelu"), 2.0 * 3136 * 512 * n)
cases["lin_wgrad"] = (lambda: lin.wgrad(h, h, n), 2.0 * 512 * 512 * n)
cases["fc_wgrad"] = (lambda: fc.wgrad(a3, h, n), 2.0 * 3136 * 512 * n)
cases["conv3_wgrad"] = (lambda: conv3.wgrad(a2, a3, n), 2.0 * 576 * 64 * 49 * n)
cases["conv2_wgrad"] = (lambda: conv2.wgrad(a1, dy2, n), 2.0 * 512 * 64 * 81 * n)
from curiosity_baselines_b200 import _lib as L
conv1 = LY.conv_layer(nn.Conv2d(1, 32, 8, 4).to(dev), (84, 84, 1), "lrelu")
frames = torch.randint(0, 256, (n, 1, 84, 84), dtype=torch.uint8, device=dev)
nm = (torch.rand(7056, device=dev) * 255, torch.rand(7056, device=dev) / 50 + 0.01)
st1 = (7056, 84, 1, 7056)
dy1 = torch.randn(n * 400, 32, device=dev).to(torch.bfloat16)
cases["conv1_fwd"] = (lambda: conv1.forward(frames.data_ptr(), n, in_dtype=L.U8, strides=st1, norm=nm), 2.0 * 64 * 32 * 400 * n)
conv1b = LY.conv_layer(nn.Conv2d(1, 32, 8, 4).to(dev), (84, 84, 1), "lrelu")
cases["conv1_twin_fwd"] = (lambda: conv1.forward(frames.data_ptr(), n, in_dtype=L.U8, strides=st1, norm=nm, twin=conv1b), 2.0 * 64 * 64 * 400 * n)
cases["conv1_wgrad"] = (lambda: conv1.wgrad(frames.data_ptr(), dy1, n, in_dtype=L.U8, strides=st1, norm=nm), 2.0 * 64 * 32 * 400 * n)
conv1c4 = LY.conv_layer(nn.Conv2d(4, 32, 8, 4).to(dev), (84, 84, 4), "lrelu")
if args.only and "c4" in args.only:      # ICM-shaped first conv: 4-channel stacks, raw pixels (3.7 GB of frames)
    frames4 = torch.randint(0, 256, (n, 4, 84, 84), dtype=torch.uint8, device=dev)
    st4 = (4 * 7056, 84, 1, 7056)
    cases["conv1c4_fwd"] = (lambda: conv1c4.forward(frames4.data_ptr(), n, in_dtype=L.U8, strides=st4), 2.0 * 256 * 32 * 400 * n)
    cases["conv1c4_wgrad"] = (lambda: conv1c4.wgrad(frames4.data_ptr(), dy1, n, in_dtype=L.U8, strides=st4), 2.0 * 256 * 32 * 400 * n)
# first two convolutions in one launch (cb_rnd_conv12_fwd), without / with the 20x20x32 activation written out
conv1.prepare(); conv2.prepare()
a2o = torch.empty(n * 81, 64, dtype=torch.bfloat16, device=dev)
fl12 = 2.0 * (64 * 32 * 400 + 512 * 64 * 81) * n


img12 = torch.empty(n * 14112, dtype=torch.uint8, device=dev)


def normalize12():
    L.call("cb_rnd_normalize_s2d", frames.data_ptr(), 7056, n, L.ptr(nm[0]), L.ptr(nm[1]), L.ptr(img12), L.stream())


def fused12(a1_out):
    L.call("cb_rnd_conv12_fwd", L.ptr(img12), n, L.ptr(conv1.w_native), L.ptr(conv1.bias), L.ptr(conv2.w_fwd),
           L.ptr(conv2.bias), L.ptr(a1_out), L.ptr(a2o), L.stream())


conv1b.prepare()
conv2b = LY.conv_layer(nn.Conv2d(32, 64, 4, 2).to(dev), (20, 20, 32), "lrelu")
conv2b.prepare()
a2o2 = torch.empty(n * 81, 64, dtype=torch.bfloat16, device=dev)


def fused12_twin(a1_out):
    L.call("cb_rnd_conv12_twin_fwd", L.ptr(img12), n, L.ptr(conv1.w_native), L.ptr(conv1.bias), L.ptr(conv2.w_fwd),
           L.ptr(conv2.bias), L.ptr(conv1b.w_native), L.ptr(conv1b.bias), L.ptr(conv2b.w_fwd), L.ptr(conv2b.bias),
           L.ptr(a2o), L.ptr(a1_out), L.ptr(a2o2), L.stream())


normalize12()
cases["fused12_normalize"] = (normalize12, 0.0)
cases["fused12_twin_fwd"] = (lambda: fused12_twin(None), 2 * fl12)
cases["fused12_twin_keep_fwd"] = (lambda: fused12_twin(a1), 2 * fl12)
cases["fused12_fwd"] = (lambda: fused12(None), fl12)
cases["fused12_keep_fwd"] = (lambda: fused12(a1), fl12)
for name, (fn, fl) in cases.items():
    if args.only and args.only not in name:
        continue
    timeit(fn, fl, name)
