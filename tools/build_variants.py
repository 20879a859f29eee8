#!/usr/bin/env python
"""Tuning experiments: build alternative libcollision_b200 variants (same sources, different -D knobs, or the
sources of another commit) into collision_b200/variants/.  The .so files are git-ignored but travel to the GPU
box with gpurun; tools/ab_bench.py times them through CLSN_LIB.  Not part of the product."""
from __future__ import annotations

import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from collision_b200 import build as b  # noqa: E402

OUT = os.path.join(ROOT, "collision_b200", "variants")

VARIANTS = {
    "c3": ["-DCULL_MIN_BLOCKS=3"],
    "c5": ["-DCULL_MIN_BLOCKS=5"],
    "c5co": ["-DCULL_MIN_BLOCKS=5", "-DCULL_SMEM_CARVEOUT=100"],   # 5 resident blocks: <= 102 registers + full carve-out
    "g4": ["-DNARROW_GRID_MULT=4"],
    "g8": ["-DNARROW_GRID_MULT=8"],
    "g32": ["-DNARROW_GRID_MULT=32"],
    "rb3": ["-DREFIT_MIN_BLOCKS=3"],
    "rb4": ["-DREFIT_MIN_BLOCKS=4"],
    "t64": ["-DTRAV_THREADS=64"],
    "t256": ["-DTRAV_THREADS=256"],
    "fa3": ["-DFAST_MIN_BLOCKS=3"],
    "fa5": ["-DFAST_MIN_BLOCKS=5"],
    "ex5": ["-DEXACT_MIN_BLOCKS=5"],
    "r6": ["-DROOTS_MIN_BLOCKS=6"],
    "pf": ["-DCULL_PREFILTER=1"],
    "nosat": ["-DCULL_SAT=0"],      # without the normal-axis separating test in k_cull   # FP32 Bernstein pre-filter in k_cull (cubic.cuh: coplanar_prefilter32)
}


def nvcc(csrc, include, flags, out):
    cmd = [b._nvcc()] + b.NVCC_FLAGS + ["-I", include, "-I", csrc] + flags + [os.path.join(csrc, "clsn.cu"), "-o", out]
    subprocess.check_call(cmd)


def main():
    os.makedirs(OUT, exist_ok=True)
    names = sys.argv[1:] or list(VARIANTS)
    procs = []
    for n in names:
        if n.startswith("rev:"):
            # sources of another commit, e.g. rev:a0a4a6e  ->  variants/libclsn_rev_a0a4a6e.so
            rev = n[4:]
            tmp = tempfile.mkdtemp()
            subprocess.check_call(f"git -C {ROOT} archive {rev} collision_b200/csrc include | tar -x -C {tmp}", shell=True)
            nvcc(os.path.join(tmp, "collision_b200", "csrc"), os.path.join(tmp, "include"), [],
                 os.path.join(OUT, f"libclsn_rev_{rev}.so"))
        else:
            nvcc(b.CSRC, b.INCLUDE, VARIANTS[n], os.path.join(OUT, f"libclsn_{n}.so"))
        print("built", n, flush=True)


if __name__ == "__main__":
    main()
