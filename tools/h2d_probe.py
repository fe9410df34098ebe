"""Times the host->device staging options for the newest-frame upload (tools only)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from curiosity_baselines_b200 import _lib as L
T, B = 128, 1024
dev = torch.device("cuda")
h = torch.empty((T, B, 4, 84, 84), dtype=torch.uint8).pin_memory()
h1 = torch.empty((T, B, 1, 84, 84), dtype=torch.uint8).pin_memory()
d = torch.empty((T, B, 1, 84, 84), dtype=torch.uint8, device=dev)
def timeit(fn, name, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f"{name:40s} {dt*1e3:8.2f} ms  {d.numel()/dt/1e9:6.1f} GB/s", flush=True)
timeit(lambda: L.call("cb_h2d_2d", d.data_ptr(), 7056, h.data_ptr() + 3 * 7056, 4 * 7056, 7056, T * B, L.stream()), "cudaMemcpy2DAsync strided (newest frame)")
timeit(lambda: d.copy_(h1, non_blocking=True), "contiguous pinned copy")
t0 = time.perf_counter()
L.call("cb_h2d_2d", d.data_ptr(), 7056, h.data_ptr() + 3 * 7056, 4 * 7056, 7056, T * B, L.stream())
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"cudaMemcpy2DAsync: call returns after {1e3*(t1-t0):.2f} ms, copy done after {1e3*(t2-t0):.2f} ms")
