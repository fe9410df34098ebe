"""Duration of the newest-frame upload while the step computes (tools only)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from curiosity_baselines_b200 import _lib as L
from curiosity_baselines_b200.curiosity.rnd import RND
from curiosity_baselines_b200.trainer import RndStep
T, B = 128, 1024
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = RND(image_shape=(4, 84, 84), drop_probability=0.25).to(dev)
step = RndStep(model)
h = torch.randint(0, 256, (T, B, 4, 84, 84), dtype=torch.uint8).pin_memory()
h1 = torch.empty((T, B, 1, 84, 84), dtype=torch.uint8).pin_memory()
bufs = [torch.empty((T, B, 1, 84, 84), dtype=torch.uint8, device=dev) for _ in range(2)]
done = torch.zeros(T, B, device=dev); value = torch.randn(T, B, device=dev); boot = torch.randn(B, device=dev)
mask = (torch.rand(T, B, device=dev) > 0.25).float()
side = torch.cuda.Stream()
for kind in ("2d", "1d", "chunks"):
    for it in range(4):
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        with torch.cuda.stream(side):
            a0.record(side)
            if kind == "2d":
                L.call("cb_h2d_2d", bufs[1].data_ptr(), 7056, h.data_ptr() + 3 * 7056, 4 * 7056, 7056, T * B, side.cuda_stream)
            elif kind == "1d":
                bufs[1].copy_(h1, non_blocking=True)
            else:
                for t in range(0, T, 8):
                    L.call("cb_h2d_2d", bufs[1][t].data_ptr(), 7056, h[t].data_ptr() + 3 * 7056, 4 * 7056, 7056, 8 * B, side.cuda_stream)
            a1.record(side)
        c0.record()
        step(bufs[0], done, torch.zeros(T, B, device=dev), value, boot, mask)
        c1.record()
        torch.cuda.synchronize()
    print(f"{kind}: copy under load {a0.elapsed_time(a1):6.2f} ms, compute {c0.elapsed_time(c1):6.2f} ms")
