"""Shared test helpers: golden loading and comparison metrics."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def rel_err(a, b):
    """max |a-b| / max(|b|, tiny) over the tensor (scale-relative max error)."""
    a = torch.as_tensor(np.asarray(a), dtype=torch.float64)
    b = torch.as_tensor(np.asarray(b), dtype=torch.float64)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def per_elem_rel(a, b, floor=0.0):
    """max over elements of |a-b| / max(|b|, floor)."""
    a = torch.as_tensor(np.asarray(a), dtype=torch.float64)
    b = torch.as_tensor(np.asarray(b), dtype=torch.float64)
    return float(((a - b).abs() / b.abs().clamp_min(max(floor, 1e-30))).max())


def cosine(a, b):
    a = torch.as_tensor(np.asarray(a), dtype=torch.float64).reshape(-1)
    b = torch.as_tensor(np.asarray(b), dtype=torch.float64).reshape(-1)
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-300))


def grad_sample(g):
    g = g.detach().reshape(-1).cpu()
    step = max(1, g.numel() // 4096)
    return g[::step]


def check_grads(named_grads, gold, cos_min, norm_tol, prefix="g."):
    """named_grads: dict name -> grad tensor.  Compares against the golden per-parameter
    norm and strided sample written by make_golden.grad_summary."""
    bad = []
    for name, g in named_grads.items():
        key = prefix + name
        if key + ".norm" not in gold:
            continue
        gn = float(gold[key + ".norm"])
        n = float(g.detach().double().norm())
        c = cosine(grad_sample(g), gold[key + ".sample"])
        if gn > 0 and (abs(n - gn) / gn > norm_tol or c < cos_min):
            bad.append((name, n, gn, c))
        if gn > 0 and os.environ.get("CB_GRAD_REPORT"):
            print(f"GRADREPORT {name} cos={c:.5f} normerr={abs(n - gn) / gn:.4f}")
    return bad
