"""CPU checks of the boundary: the C-ABI library loads and exports every symbol that
include/curiosity_b200.h declares; the ctypes signature table covers them all; the host
layer refuses to run without a GPU instead of falling back."""
import os
import re

import pytest
import torch

from curiosity_baselines_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "curiosity_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    lib = L.load()
    names = _declared()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_ctypes_table_matches_header():
    assert sorted(L.EXPORTS) == _declared()


def test_struct_layout_matches_header():
    # cb_conv_desc: 12 int32, 4 int64, 2 int32 ; cb_epilogue: 4 ptr, int32(+pad), 5 ptr
    import ctypes as C
    assert C.sizeof(L.ConvDesc) == 12 * 4 + 4 * 8 + 2 * 4
    assert C.sizeof(L.Epilogue) == 4 * 8 + 8 + 5 * 8
    assert L.ConvDesc.in_stride_n.offset == 48 and L.Epilogue.norm_mean.offset == 40


def test_version_and_error_string():
    lib = L.load()
    assert lib.cb_version() >= 100
    # argument validation happens before any CUDA call, so it is checkable on CPU
    rc = lib.cb_gae(None, None, None, None, 0.99, 0.95, None, None, 4, 4, None)
    assert rc == -1 and b"null" in lib.cb_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback():
    from curiosity_baselines_b200.algos.utils import generalized_advantage_estimation
    r = torch.zeros(4, 3)
    with pytest.raises(L.CuriosityLibError):
        generalized_advantage_estimation(r, r, r, torch.zeros(3), 0.99, 0.95)
