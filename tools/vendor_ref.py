"""Vendor the pure-Python subset of the reference that the hot path imports into the git-ignored
``baseline/_ref/`` so that the UNMODIFIED reference travels to the GPU box with the repo snapshot
(``/root/reference`` only exists in the build container).  Used by ``bench.py --impl reference`` / ``cpu_baseline``
(kind "reference") and by the reference-driven drop-in tests.  Run in the build container:

    python tools/vendor_ref.py            # also called by __graft_entry__.build() when /root/reference exists

The file list is found by importing the path's modules from the checkout and recording every ``rlpyt`` source file
that was loaded; files are copied byte for byte and their sha256 written to ``baseline/_ref/MANIFEST.txt``.  Nothing
is copied into tracked directories.
"""
import hashlib
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SRC = os.environ.get("CURIOSITY_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")

HOT_PATH_MODULES = [
    "rlpyt.models.curiosity.icm", "rlpyt.models.curiosity.disagreement", "rlpyt.models.curiosity.rnd",
    "rlpyt.models.curiosity.ndigo", "rlpyt.models.curiosity.encoders", "rlpyt.models.pg.atari_lstm_model",
    "rlpyt.algos.utils", "rlpyt.algos.pg.base", "rlpyt.algos.pg.ppo", "rlpyt.algos.pg.a2c",
    "rlpyt.agents.pg.atari", "rlpyt.agents.pg.categorical", "rlpyt.utils.averages", "rlpyt.utils.tensor",
    "rlpyt.samplers.collections", "rlpyt.envs.base", "rlpyt.spaces.int_box", "rlpyt.spaces.float_box", "rlpyt.distributions.categorical",
]


def vendor(verbose=True):
    if not os.path.isfile(os.path.join(SRC, "rlpyt", "algos", "pg", "ppo.py")):
        raise SystemExit(f"no reference checkout at {SRC}")
    os.environ["CURIOSITY_REFERENCE"] = SRC
    from oracle import refshim
    refshim.import_reference()
    import importlib
    for name in HOT_PATH_MODULES:
        importlib.import_module(name)
    files = set()
    src_real = os.path.realpath(SRC)
    for name, mod in list(sys.modules.items()):
        f = getattr(mod, "__file__", None)
        if name.split(".")[0] == "rlpyt" and f and os.path.realpath(f).startswith(src_real + os.sep):
            files.add(os.path.relpath(os.path.realpath(f), src_real))
    # package markers of every directory on the way (namespace packages have none: keep it that way)
    for rel in list(files):
        d = os.path.dirname(rel)
        while d:
            init = os.path.join(d, "__init__.py")
            if os.path.isfile(os.path.join(src_real, init)):
                files.add(init)
            d = os.path.dirname(d)
    for extra in ("LICENSE", "LICENSE.txt", "LICENSE.md"):
        if os.path.isfile(os.path.join(src_real, extra)):
            files.add(extra)
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    lines = []
    for rel in sorted(files):
        out = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(os.path.join(src_real, rel), out)
        with open(out, "rb") as fh:
            lines.append(f"{hashlib.sha256(fh.read()).hexdigest()}  {rel}")
    with open(os.path.join(DST, "MANIFEST.txt"), "w") as fh:
        fh.write(f"# unmodified files of {SRC} (sha256  path); written by tools/vendor_ref.py\n" + "\n".join(lines) + "\n")
    if verbose:
        total = sum(os.path.getsize(os.path.join(DST, r)) for r in files)
        print(f"vendored {len(files)} files ({total / 1024:.0f} KB) into {DST}")
    return sorted(files)


if __name__ == "__main__":
    vendor()
