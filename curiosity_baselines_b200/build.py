"""In-tree build of libcuriosity_b200.so with nvcc for sm_100a.

    python -m curiosity_baselines_b200.build [--force] [--probe]

nvcc cross-compiles without a GPU.  Every .cu is compiled to an object under ``csrc/_obj/`` (in
parallel, only when it or a header changed) and linked into the .so next to this file, which travels
to the GPU box with the repo snapshot.  ``--probe`` additionally builds ``libcuriosity_b200_probe.so``,
the kernel author's descriptor probe used by tests/test_gpu_umma_probe.py -- test infrastructure kept
out of the product library.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libcuriosity_b200.so")
PROBE_LIB = os.path.join(HERE, "libcuriosity_b200_probe.so")
PROBE_SOURCES = ["umma_probe.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu") and f not in PROBE_SOURCES)


def _headers_mtime():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "curiosity_b200.h"))
    return max(os.path.getmtime(d) for d in deps)


def _nvcc():
    return os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return r.stdout + r.stderr


def _compile(srcs, lib, force, verbose, extra=()):
    os.makedirs(OBJ, exist_ok=True)
    hdr = _headers_mtime()
    todo, objs = [], []
    for s in srcs:
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ, s[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr):
            todo.append([_nvcc()] + NVCC_FLAGS + list(extra) + ["-c", src, "-o", obj])
    if todo:
        with ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 4)) as ex:
            list(ex.map(lambda c: _run(c, verbose), todo))
    if todo or not os.path.exists(lib) or any(os.path.getmtime(o) > os.path.getmtime(lib) for o in objs):
        _run([_nvcc()] + ARCH + ["-shared", "-o", lib] + objs, verbose)
    return lib


def build(force=False, verbose=False):
    return _compile(sources(), LIB, force, verbose)


def build_probe(force=False, verbose=False):
    return _compile(PROBE_SOURCES, PROBE_LIB, force, verbose)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
    if "--probe" in sys.argv:
        print(build_probe(force="--force" in sys.argv, verbose=True))
