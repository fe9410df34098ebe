"""Probe for the conv1 weight-gradient formulation (tools only): one MMA per k-step gives all four
2x2 taps from a SINGLE dy grid image and a SINGLE space-to-depth frame image:
  A = dy grid, MN-major SWIZZLE_64B (rows of 32 channels), M = 64 = two chunks 21 rows apart (LBO = 21*64)
  B = s2d frame, MN-major SWIZZLE_32B (rows of 16 pixels), N = 32 = two chunks one row apart (LBO = 32)
  D[(j,n)][(i,e)] = sum_q dyP[q + 21 j][n] * img[q + i][e]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from probe_umma import run, bf16_bits, rel  # noqa: E402  (runs the base probe table on import)

torch.manual_seed(1)
rb = lambda *s: torch.randn(*s).to(torch.bfloat16).float()
swz64 = lambda off: off ^ (((off >> 7) & 3) << 4)
swz32 = lambda off: off ^ (((off >> 7) & 1) << 4)

ROWS_A, ROWS_B, KSTEPS, G = 504, 480, 29, 21
dyP = rb(ROWS_A, 32)
img = rb(ROWS_B, 16)
a_img = np.zeros(ROWS_A * 32 + 4096, dtype=np.int16)
r, c = np.meshgrid(np.arange(ROWS_A), np.arange(32), indexing="ij")
a_img[swz64(r * 64 + c * 2) // 2] = bf16_bits(dyP)
b_img = np.zeros(ROWS_B * 16 + 4096, dtype=np.int16)
r, c = np.meshgrid(np.arange(ROWS_B), np.arange(16), indexing="ij")
b_img[swz32(r * 32 + c * 2) // 2] = bf16_bits(img)
K = 16 * KSTEPS
ref = torch.zeros(64, 32)
for j in range(2):
    for i in range(2):
        ref[j * 32:(j + 1) * 32, i * 16:(i + 1) * 16] = dyP[G * j:G * j + K].t() @ img[i:i + K]
d = run(a_img, b_img, (0, G * 64, 512, 0, 1024, 4), (0, 32, 256, 0, 512, 6), 1, 1, 64, 32, KSTEPS)
got = torch.stack([d[(m // 16) * 32 + m % 16, :32] for m in range(64)])
print(f"W1 SW64 A (LBO=21 rows) x SW32 B (LBO=1 row), M=64 N=32, {KSTEPS} k-steps: err {rel(got, ref):.3e}")
for j in range(2):
    for i in range(2):
        print(f"   block j={j} i={i}: err {rel(got[j*32:(j+1)*32, i*16:(i+1)*16], ref[j*32:(j+1)*32, i*16:(i+1)*16]):.3e}")
