"""Exploratory probe of tcgen05 descriptor semantics on the installed GPU (prints a table)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from curiosity_baselines_b200 import _lib as L  # noqa: E402

dev = "cuda"


def swz(off):
    return off ^ (((off >> 7) & 7) << 4)


def put(img, off, val_bf16_bits):
    img[swz(off) // 2] = val_bf16_bits


def bf16_bits(x):
    return torch.tensor(x, dtype=torch.float32).to(torch.bfloat16).view(torch.int16).numpy()


def run(a_img, b_img, a_desc, b_desc, a_mn, b_mn, M, N, ksteps):
    a = torch.from_numpy(a_img.view(np.int16)).to(dev)
    b = torch.from_numpy(b_img.view(np.int16)).to(dev)
    out = torch.full((128, N), float("nan"), device=dev)
    ad = (C.c_int * 6)(*(tuple(a_desc) + (0,) * (6 - len(a_desc))))
    bd = (C.c_int * 6)(*(tuple(b_desc) + (0,) * (6 - len(b_desc))))
    L.call_probe(L.ptr(a), a.numel() * 2, L.ptr(b), b.numel() * 2, ad, bd, a_mn, b_mn, M, N, ksteps,
           L.ptr(out), L.stream())
    torch.cuda.synchronize()
    return out.cpu()


def kmajor_image(mat, rows_alloc):
    """mat [rows, 64] fp32 -> SW128 K-major image (rows of 128 B, address-based swizzle)."""
    bits = bf16_bits(mat)
    img = np.zeros(rows_alloc * 64, dtype=np.int16)
    r, k = np.meshgrid(np.arange(mat.shape[0]), np.arange(64), indexing="ij")
    off = r * 128 + k * 2
    img[swz(off) // 2] = bits
    return img


def mnmajor_image(mat, krows_alloc, chunk_stride):
    """mat [MN, K] fp32 -> SW128 MN-major image: chunk c = mn//64 at c*chunk_stride; K row k at k*128."""
    bits = bf16_bits(mat)
    nchunks = (mat.shape[0] + 63) // 64
    img = np.zeros(max(nchunks * chunk_stride, krows_alloc * 128) // 2 + 64 * 64, dtype=np.int16)
    m, k = np.meshgrid(np.arange(mat.shape[0]), np.arange(mat.shape[1]), indexing="ij")
    off = (m // 64) * chunk_stride + k * 128 + (m % 64) * 2
    img[swz(off) // 2] = bits
    return img


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


torch.manual_seed(0)
rb = lambda *s: torch.randn(*s).to(torch.bfloat16).float()

# E0: K-major sanity
A = rb(160, 64); B = rb(64, 64)
a_img = kmajor_image(A, 160); b_img = kmajor_image(B, 64)
d = run(a_img, b_img, (0, 16, 1024, 0, 32), (0, 16, 1024, 0, 32), 0, 0, 128, 64, 4)
print("E0 K-major sanity", rel(d, A[:128] @ B.t()))

# E1: K-major A with a row-shifted start
for shift in (1, 3, 8, 9, 20):
    for bo in sorted({0, shift & 7}):
        d = run(a_img, b_img, (shift * 128, 16, 1024, bo, 32), (0, 16, 1024, 0, 32), 0, 0, 128, 64, 4)
        print(f"E1 K-major shift={shift} base_off={bo}: err {rel(d, A[shift:shift + 128] @ B.t()):.3e}")

# E2: MN-major both operands (K = 64 rows)
A2 = rb(128, 64); B2 = rb(64, 64)      # [MN, K]
for (lbo, sbo) in ((8192, 1024), (1024, 8192)):
    a2 = mnmajor_image(A2, 64, 8192); b2 = mnmajor_image(B2, 64, 8192)
    d = run(a2, b2, (0, lbo, sbo, 0, 2048), (0, lbo, sbo, 0, 2048), 1, 1, 128, 64, 4)
    print(f"E2 MN-major lbo={lbo} sbo={sbo}: err {rel(d, A2 @ B2.t()):.3e}")

# E2b: MN-major A, K-major B and vice versa
a2 = mnmajor_image(A2, 64, 8192); bk = kmajor_image(B2, 64)
d = run(a2, bk, (0, 8192, 1024, 0, 2048), (0, 16, 1024, 0, 32), 1, 0, 128, 64, 4)
print(f"E2b MN-major A x K-major B: err {rel(d, A2 @ B2.t()):.3e}")

# E3: MN-major B with a K-row-shifted start (rows = K index), 96 K-rows allocated
B3 = rb(64, 96)
b3 = mnmajor_image(B3, 96, 96 * 128)
for shift in (1, 5, 8, 11):
    for bo in sorted({0, shift & 7}):
        d = run(a2, b3, (0, 8192, 1024, 0, 2048), (shift * 128, 96 * 128, 1024, bo, 2048), 1, 1, 128, 64, 4)
        print(f"E3 MN-major B K-shift={shift} base_off={bo}: err {rel(d, A2 @ B3[:, shift:shift + 64].t()):.3e}")

# E4: M=64 accumulator layout: which TMEM lanes hold row i
A4 = rb(64, 64)
a4 = kmajor_image(A4, 64)
d = run(a4, b_img, (0, 16, 1024, 0, 32), (0, 16, 1024, 0, 32), 0, 0, 64, 64, 4)
ref = A4 @ B.t()
lanes = []
for i in range(64):
    hit = [l for l in range(128) if torch.isfinite(d[l]).all() and rel(d[l:l + 1], ref[i:i + 1]) < 1e-2]
    lanes.append(hit[0] if hit else -1)
print("E4 M=64 row->lane", lanes[:20], "...", lanes[60:])

# E5: MN-major A whose two 64-wide chunks are windows one K-row apart (LBO = 128 B)
X = rb(64, 96)               # [c, rows]
x_img = mnmajor_image(X, 96, 96 * 128)
for bo in (0,):
    d = run(x_img, b2, (0, 128, 1024, bo, 2048), (0, 8192, 1024, 0, 2048), 1, 1, 128, 64, 4)
    ref = torch.cat([X[:, 0:64], X[:, 1:65]], 0) @ B2.t()
    print(f"E5 MN-major A two overlapping chunks (LBO=128): err {rel(d, ref):.3e}  (first half {rel(d[:64], ref[:64]):.3e})")

# E6: K-major SWIZZLE_32B (layout code 6): rows of 32 B (K = 16), 8-row atoms of 256 B, row-shifted starts
def swz32(off):
    return off ^ (((off >> 7) & 1) << 4)


def kmajor32_image(mat, rows_alloc, swz_fn):
    b = bf16_bits(mat)
    img = np.zeros(rows_alloc * 16, dtype=np.int16)
    r, k = np.meshgrid(np.arange(mat.shape[0]), np.arange(16), indexing="ij")
    img[swz_fn(r * 32 + k * 2) // 2] = b
    return img


A6 = rb(200, 16); B6 = rb(64, 16)
for name, fn in (("addr-swizzle", swz32), ("no-swizzle", lambda o: o)):
    a6 = kmajor32_image(A6, 256, fn); b6 = kmajor32_image(B6, 64, fn)
    for shift in (0, 1, 5, 21, 22):
        d = run(a6, b6, (shift * 32, 16, 256, 0, 0, 6), (0, 16, 256, 0, 0, 6), 0, 0, 128, 64, 1)
        print(f"E6 SW32 K-major {name} shift={shift}: err {rel(d, A6[shift:shift + 128] @ B6.t()):.3e}")

# E7: MN-major SWIZZLE_32B B operand with N = 16 (one 32-byte chunk per K row), K-row-shifted start,
#     against an MN-major SWIZZLE_128B A operand with M = 64
def mn32_image(mat, krows_alloc):
    """mat [16, K] -> K rows of 32 B (16 MN elements), 32-byte swizzle."""
    b = bf16_bits(mat)
    img = np.zeros(krows_alloc * 16, dtype=np.int16)
    m, k = np.meshgrid(np.arange(16), np.arange(mat.shape[1]), indexing="ij")
    img[swz32(k * 32 + m * 2) // 2] = b
    return img


A7 = rb(64, 64)                       # [M, K]  (dy channels x positions)
B7 = rb(16, 96)                       # [N, K]  (patch elements x block rows)
a7 = mnmajor_image(A7, 64, 8192)
b7 = mn32_image(B7, 128)
for shift in (0, 1, 21, 22):
    d = run(a7, b7, (0, 8192, 1024, 0, 2048, 0), (shift * 32, 16, 256, 0, 512, 6), 1, 1, 64, 16, 4)
    ref = A7 @ B7[:, shift:shift + 64].t()            # [64, 16]
    got = torch.stack([d[(i // 16) * 32 + i % 16, :16] for i in range(64)])
    print(f"E7 MN-major SW32 N=16 shift={shift}: err {rel(got, ref):.3e}")
