"""Timeline of upload vs compute in a double-buffered loop (tools only)."""
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
bufs = [torch.empty((T, B, 1, 84, 84), dtype=torch.uint8, device=dev) for _ in range(2)]
done = torch.zeros(T, B, device=dev); value = torch.randn(T, B, device=dev); boot = torch.randn(B, device=dev)
mask = (torch.rand(T, B, device=dev) > 0.25).float()
side = torch.cuda.Stream()
main = torch.cuda.current_stream()
E = lambda: torch.cuda.Event(enable_timing=True)
def upload(i):
    a0, a1 = E(), E()
    a0.record(side)
    L.call("cb_h2d_2d", bufs[i].data_ptr(), 7056, h.data_ptr() + 3 * 7056, 4 * 7056, 7056, T * B, side.cuda_stream)
    a1.record(side)
    return a0, a1
for warm in range(2):
    step(bufs[0], done, torch.zeros(T, B, device=dev), value, boot, mask)
torch.cuda.synchronize()
t0 = E(); t0.record(main)
side.wait_stream(main)
rec = []
up = upload(0)
free = [None, None]
for it in range(5):
    b = it % 2
    main.wait_event(up[1])
    c0, c1 = E(), E()
    c0.record(main)
    step(bufs[b], done, torch.zeros(T, B, device=dev), value, boot, mask)
    c1.record(main)
    free[b] = c1
    if free[1 - b] is not None:
        side.wait_event(free[1 - b])
    nxt = upload(1 - b)
    rec.append((up, c0, c1))
    up = nxt
torch.cuda.synchronize()
for up_, c0, c1 in rec:
    print(f"upload {t0.elapsed_time(up_[0]):7.2f} -> {t0.elapsed_time(up_[1]):7.2f}   compute {t0.elapsed_time(c0):7.2f} -> {t0.elapsed_time(c1):7.2f}")
