"""Pins the tcgen05 descriptor semantics the window / wgrad kernels rely on, on the installed GPU
(through the diagnostic entry point cb_umma_probe; see tools/probe_umma.py for the exploratory form)."""
import ctypes as C

import numpy as np
import pytest
import torch

from curiosity_baselines_b200 import _lib as L

pytestmark = pytest.mark.gpu
DEV = "cuda"


def swz(off):
    return off ^ (((off >> 7) & 7) << 4)


def bits(x):
    return x.to(torch.bfloat16).view(torch.int16).numpy()


def run(a_img, b_img, a_desc, b_desc, a_mn, b_mn, M, N, ksteps):
    a, b = torch.from_numpy(a_img).to(DEV), torch.from_numpy(b_img).to(DEV)
    out = torch.full((128, N), float("nan"), device=DEV)
    L.call_probe(L.ptr(a), a.numel() * 2, L.ptr(b), b.numel() * 2, (C.c_int * 6)(*(tuple(a_desc) + (0,) * (6 - len(a_desc)))),
           (C.c_int * 6)(*(tuple(b_desc) + (0,) * (6 - len(b_desc)))), a_mn, b_mn, M, N, ksteps, L.ptr(out), L.stream())
    torch.cuda.synchronize()
    return out.cpu()


def kmajor(mat, rows_alloc):
    img = np.zeros(rows_alloc * 64, dtype=np.int16)
    r, k = np.meshgrid(np.arange(mat.shape[0]), np.arange(64), indexing="ij")
    img[swz(r * 128 + k * 2) // 2] = bits(mat)
    return img


def mnmajor(mat, chunk_stride, alloc_bytes):
    img = np.zeros(alloc_bytes // 2, dtype=np.int16)
    m, k = np.meshgrid(np.arange(mat.shape[0]), np.arange(mat.shape[1]), indexing="ij")
    img[swz((m // 64) * chunk_stride + k * 128 + (m % 64) * 2) // 2] = bits(mat)
    return img


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


def rb(*s):
    return torch.randn(*s).to(torch.bfloat16).float()


def test_kmajor_row_shifted_start_needs_no_base_offset():
    torch.manual_seed(0)
    A, B = rb(160, 64), rb(64, 64)
    a, b = kmajor(A, 160), kmajor(B, 64)
    for shift in (0, 1, 3, 9, 20):
        d = run(a, b, (shift * 128, 16, 1024, 0, 32), (0, 16, 1024, 0, 32), 0, 0, 128, 64, 4)
        assert rel(d, A[shift:shift + 128] @ B.t()) < 1e-5, shift


def test_mnmajor_layout_and_shift():
    torch.manual_seed(1)
    A, B = rb(128, 64), rb(64, 96)                    # [MN, K]
    a = mnmajor(A, 8192, 2 * 8192)
    b = mnmajor(B, 96 * 128, 2 * 96 * 128)
    for shift in (0, 5, 11):
        d = run(a, b, (0, 8192, 1024, 0, 2048), (shift * 128, 96 * 128, 1024, 0, 2048), 1, 1, 128, 64, 4)
        assert rel(d, A @ B[:, shift:shift + 64].t()) < 1e-5, shift


def test_overlapping_chunks_and_m64_lanes():
    torch.manual_seed(2)
    X, B = rb(64, 96), rb(64, 64)
    x, b = mnmajor(X, 96 * 128, 2 * 96 * 128), mnmajor(B, 8192, 2 * 8192)
    d = run(x, b, (0, 128, 1024, 0, 2048), (0, 8192, 1024, 0, 2048), 1, 1, 128, 64, 4)
    assert rel(d, torch.cat([X[:, 0:64], X[:, 1:65]], 0) @ B.t()) < 1e-5      # two windows one row apart
    A4, Bk = rb(64, 64), rb(64, 64)
    d = run(kmajor(A4, 64), kmajor(Bk, 64), (0, 16, 1024, 0, 32), (0, 16, 1024, 0, 32), 0, 0, 64, 64, 4)
    ref = A4 @ Bk.t()
    for i in (0, 15, 16, 40, 63):
        lane = (i // 16) * 32 + i % 16
        assert rel(d[lane:lane + 1], ref[i:i + 1]) < 1e-5, i
