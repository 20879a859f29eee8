"""In-tree build of the CUDA library (sm_100a only).  nvcc cross-compiles without a GPU."""
from __future__ import annotations

import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libcollision_b200.so")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")

SOURCES = ["clsn.cu"]
HEADERS = ["narrow.cuh", "cubic.cuh", "fastpath.cuh", "lbvh.cuh", "reduce.cuh", "rigid.cuh", "strain.cuh", "dist.cuh", "crmath.cuh",
           "crmath_constants.inc"]

# --fmad=false: FP64 expressions must round exactly like the reference's (no FMA contraction);
# the double-double code in crmath.cuh issues its FMAs explicitly.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


HASH_PATH = LIB_PATH + ".srchash"


def _source_hash() -> str:
    """Content hash of everything the library is built from (file times do not survive the copy to the GPU box)."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for d in [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(INCLUDE, "collision_b200.h")]:
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH) or not os.path.exists(HASH_PATH):
        return True
    with open(HASH_PATH) as f:
        return f.read().strip() != _source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile into a temporary file and rename it into place, under a file lock: the ranks of a torchrun job may all
    find the library stale at the same time, and a half-written .so must never be loadable."""
    import fcntl
    with open(LIB_PATH + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():
                return LIB_PATH
            tmp = LIB_PATH + f".tmp{os.getpid()}"
            cmd = [_nvcc()] + NVCC_FLAGS + ["-I", INCLUDE, "-I", CSRC]
            if verbose:
                cmd += ["-Xptxas", "-v"]
            cmd += [os.path.join(CSRC, s) for s in SOURCES] + ["-o", tmp]
            subprocess.check_call(cmd)
            os.replace(tmp, LIB_PATH)
            with open(HASH_PATH, "w") as f:
                f.write(_source_hash())
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
