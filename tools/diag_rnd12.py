"""Determinism / error-location probe for cb_rnd_conv12_fwd (development aid)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from curiosity_baselines_b200 import _lib as L
from curiosity_baselines_b200.curiosity.rnd import RND

n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 1043
torch.manual_seed(0)
m = RND(image_shape=(4, 84, 84)).cuda(); m._build()
m._norm[0].copy_(torch.rand(7056).cuda() * 200 + 20); m._norm[1].copy_(torch.rand(7056).cuda() * 0.05 + 0.005)
obs = torch.randint(0, 256, (1, n_total, 4, 84, 84), dtype=torch.uint8, device="cuda")
obs, T, B, base, strides = m._frame_view(obs)
n = T * B
ch = m._chains[0]
a1r, _ = ch.layers[0].forward(base, n, in_dtype=L.U8, strides=strides, norm=m._norm)
a2r, _ = ch.layers[1].forward(a1r, n)
runs = []
for keep in (True, False, True, False, False, True, True, False, True, False):
    a1, a2 = m._conv12(ch, base, n, strides[0], keep)
    torch.cuda.synchronize()
    runs.append((keep, a1, a2))
ref = runs[1][2]
for k, (keep, a1, a2) in enumerate(runs):
    d = (a2.float() - ref.float()).abs().view(n, -1).amax(1)
    bad = torch.nonzero(d > 0).flatten()
    dr = (a2.float() - a2r.float()).abs().view(n, -1).amax(1)
    badr = torch.nonzero(dr > 0.1 * float(a2r.float().abs().max())).flatten()
    msg = f"run {k} keep={keep}: samples differing from run1: {bad.numel()} {bad[:12].tolist()}  | far from separate kernels: {badr.numel()} {badr[:12].tolist()}"
    if a1 is not None:
        d1 = (a1.float() - a1r.float()).abs().view(n, -1).amax(1)
        b1 = torch.nonzero(d1 > 0.05 * float(a1r.float().abs().max())).flatten()
        msg += f" | a1 far: {b1.numel()} {b1[:12].tolist()}"
    print(msg)
a1r2, _ = ch.layers[0].forward(base, n, in_dtype=L.U8, strides=strides, norm=m._norm)
a2r2, _ = ch.layers[1].forward(a1r2, n)
print("separate kernels reproducible:", torch.equal(a1r, a1r2), torch.equal(a2r, a2r2))
for k, (keep, a1, a2) in enumerate(runs):
    A2 = a2.float().view(n, -1); R2 = a2r.float().view(n, -1)
    bad = torch.nonzero((A2 - R2).abs().amax(1) > 0.1 * float(R2.abs().max())).flatten().tolist()
    print("run", k, keep, "bad samples", bad, "(j:", [b // 148 for b in bad], ")")
