"""Where does the end-to-end step spend its time?  (tools only)"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from curiosity_baselines_b200.curiosity.rnd import RND
from curiosity_baselines_b200.trainer import RndStep
from curiosity_baselines_b200.agents.hooks import FrameStager
T, B = 128, 1024
dev = torch.device("cuda")
torch.manual_seed(0)
model = RND(image_shape=(4, 84, 84), drop_probability=0.25).to(dev)
step = RndStep(model)
h_frames = torch.randint(0, 256, (T, B, 4, 84, 84), dtype=torch.uint8).pin_memory()
done = torch.zeros(T, B, device=dev); value = torch.randn(T, B, device=dev); boot = torch.randn(B, device=dev)
mask = (torch.rand(T, B, device=dev) > 0.25).float()
stager = FrameStager(dev)
stager.prefetch(h_frames)
mode = sys.argv[1] if len(sys.argv) > 1 else "overlap"
evs = []
for it in range(8):
    d = stager.get()
    if mode == "serial":
        torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    c0.record()
    step(d, done, torch.zeros(T, B, device=dev), value, boot, mask)
    c1.record()
    t1 = time.perf_counter()
    stager.done()
    if mode == "serial":
        torch.cuda.synchronize()
    stager.prefetch(h_frames)
    evs.append((c0, c1, t1 - t0))
torch.cuda.synchronize()
for c0, c1, cpu in evs:
    print(f"{mode}: compute {c0.elapsed_time(c1):7.2f} ms   cpu launch time {cpu*1e3:6.2f} ms")

# wall clock per step, adding the bench's extra transfers one at a time
h_small = [t.cpu().pin_memory() for t in (done, value, boot, mask)]
h_out = [torch.empty(T, B).pin_memory() for _ in range(3)]
h_loss = torch.empty(1).pin_memory()
def run2(n=6):
    """prefetch queued BEFORE the step's kernels"""
    stager.prefetch(h_frames)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        d = stager.get()
        stager.prefetch(h_frames)
        dd, dv, db, dm = (t.to(dev, non_blocking=True) for t in h_small)
        loss, r_int, adv, ret = step(d, dd, torch.zeros(T, B, device=dev), dv, db, dm)
        stager.done()
        h_out[0].copy_(r_int, non_blocking=True); h_out[1].copy_(adv, non_blocking=True)
        h_out[2].copy_(ret, non_blocking=True); h_loss.copy_(loss.reshape(1), non_blocking=True)
    torch.cuda.synchronize()
    stager.get()
    print(f"prefetch-before-compute: {(time.perf_counter()-t0)*1e3/n:.2f} ms/step")


def run(small, d2h, n=6):
    stager.prefetch(h_frames)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        d = stager.get()
        if small:
            dd, dv, db, dm = (t.to(dev, non_blocking=True) for t in h_small)
        else:
            dd, dv, db, dm = done, value, boot, mask
        loss, r_int, adv, ret = step(d, dd, torch.zeros(T, B, device=dev), dv, db, dm)
        stager.done()
        if d2h:
            h_out[0].copy_(r_int, non_blocking=True); h_out[1].copy_(adv, non_blocking=True)
            h_out[2].copy_(ret, non_blocking=True); h_loss.copy_(loss.reshape(1), non_blocking=True)
        stager.prefetch(h_frames)
    torch.cuda.synchronize()
    stager.get()
    print(f"small_h2d={small} d2h={d2h}: {(time.perf_counter()-t0)*1e3/n:.2f} ms/step")
run(False, False); run(True, False); run(False, True); run(True, True)
