"""Repeatability probe (regression check for the raw-frame-slot race): the stand-alone first-conv
kernel and the fused conv1+conv2 kernel are launched many times on the same frames and every result is compared
bit for bit with the first one.      python tools/repeat_conv1.py [samples=4000] [launches=20]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from curiosity_baselines_b200 import _lib as L
from curiosity_baselines_b200.curiosity.rnd import RND

n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
torch.manual_seed(0)
m = RND(image_shape=(4, 84, 84)).cuda(); m._build()
m._norm[0].copy_(torch.rand(7056).cuda() * 200 + 20); m._norm[1].copy_(torch.rand(7056).cuda() * 0.05 + 0.005)
obs = torch.randint(0, 256, (1, n_total, 4, 84, 84), dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
obs_copy = obs.clone()
obs, T, B, base, strides = m._frame_view(obs)
n = T * B
ch = m._chains[0]
img = m._normalized(base, n, strides[0])
ref = None
bad_sep = {}
for r in range(reps):
    a1, _ = ch.layers[0].forward(base, n, in_dtype=L.U8, strides=strides, norm=m._norm)
    torch.cuda.synchronize()
    if ref is None:
        ref = a1
    else:
        d = torch.nonzero((a1.float() - ref.float()).abs().view(n, -1).amax(1) > 0).flatten().tolist()
        if d: bad_sep[r] = d
print("separate conv1: launches differing from launch 0:", bad_sep)
print("frames unchanged:", torch.equal(obs, obs_copy))
for keep in (True, False):
    bad = {}
    first = None
    for r in range(reps):
        a1, a2 = m._conv12(ch, img, n, keep)
        torch.cuda.synchronize()
        if first is None:
            first = a2
        d = torch.nonzero((a2.float() - first.float()).abs().view(n, -1).amax(1) > 0).flatten().tolist()
        if d: bad[r] = d
    print(f"fused keep={keep}: launches differing from launch 0:", bad)
print("frames unchanged:", torch.equal(obs, obs_copy))
