"""Build libpsf_b200.so in-tree with nvcc for sm_100a (no torch headers: compiles in seconds).

    python -m sparsefactorization_b200.build [--force] [--verbose]

The shared object lands next to this file (``_lib/libpsf_b200.so``); it is git-ignored but
travels with the repo snapshot to the GPU box.  The CUDA runtime is linked statically so the
library does not depend on which libcudart torch happens to load.
"""
from __future__ import annotations

import fcntl
import glob
import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libpsf_b200.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC,-O3", "-shared", "-cudart", "static",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _source_hash() -> str:
    """Content hash of everything the library is built from (mtimes do not survive a copy of the tree)."""
    h = hashlib.sha256()
    deps = sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + sorted(glob.glob(os.path.join(ROOT, "include", "*.h")))
    for d in deps:
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    try:
        with open(LIB_PATH + ".srchash") as f:
            return f.read().strip() != _source_hash()
    except OSError:
        return True


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (needed to build libpsf_b200.so)")


def build(force: bool = False, verbose: bool = False, defines=(), out: str = None) -> str:
    """Build (when the sources changed since the last build, or ``force``) and return the library path.
    ``defines`` / ``out``: tuning builds (e.g. ``-DPSF2_NW_BWD=6``) written next to the default library and
    selected at run time with the PSF_B200_LIB environment variable.
    Safe under torchrun: one process builds (file lock), into a temporary file that is renamed into place."""
    out = out or LIB_PATH
    default = out == LIB_PATH
    if not force and default and not _stale():
        return LIB_PATH
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(os.path.join(os.path.dirname(out), ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and default and not _stale():       # another rank built it while we waited
                return LIB_PATH
            tmp = "{}.tmp.{}".format(out, os.getpid())
            cmd = [find_nvcc(), *NVCC_FLAGS, *defines, "-I", os.path.join(ROOT, "include"), "-I", CSRC]
            if verbose:
                cmd += ["-Xptxas", "-v"]
            cmd += ["-o", tmp, *sources()]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed building libpsf_b200.so")
            os.replace(tmp, out)
            if default:
                with open(LIB_PATH + ".srchash", "w") as f:
                    f.write(_source_hash())
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return out


if __name__ == "__main__":
    _out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    print(build(force="--force" in sys.argv or _out is not None, verbose="--verbose" in sys.argv,
                defines=[a for a in sys.argv[1:] if a.startswith("-D")], out=_out))
