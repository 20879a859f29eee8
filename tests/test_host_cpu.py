"""CPU tests of the host side: scene generators, the C ABI surface, failing loudly without a GPU,
the correctly rounded math of the kernels (host build of crmath.cuh) and the multi-rank exchange."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_config4_is_one_million_triangles():
    from collision_b200 import scenes
    sc = scenes.layered_cloth(8, 251)
    assert sc.T == 1_000_000 and sc.V == 504_008 and sc.n_surf == 8
    sc2 = scenes.layered_cloth(8, 251)
    assert np.array_equal(sc.x, sc2.x) and np.array_equal(sc.vel, sc2.vel)
    assert (np.diff(sc.tri_surf) >= 0).all()


def test_reference_decks_restated():
    from collision_b200 import scenes
    s = scenes.string_string()
    assert s.B == 2 * 192 and s.V == 2 * 193 and s.T == 0          # cdinit.cpp:157-159
    b = scenes.ball_plane()
    assert (b.vflags[b.vhs == 0] == 2).all() and (b.vflags[b.vhs == 1] == 0).all()
    assert np.allclose(b.lo, [0, 0, 0.25]) and np.allclose(b.hi, [0.5, 0.5, 0.75])
    bb = scenes.box_boundary()
    assert (bb.vflags == 2).all()


def _lib_path():
    from collision_b200 import build
    return build.build()


def test_c_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "collision_b200.h")).read()
    names = sorted(set(re.findall(r"\b(clsn_[a-z_0-9]+)\s*\(", hdr)))
    assert len(names) >= 30
    lib = ctypes.CDLL(_lib_path())
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/collision_b200.h but not exported: {missing}"


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to work instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from collision_b200.solver import CollisionError, CollisionSolver3d
    with pytest.raises(CollisionError):
        CollisionSolver3d()


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "collision_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f
                assert "collision_oracle" not in src and "libcollision_ref" not in src, f


def test_static_parameter_api():
    from collision_b200.solver import CollisionSolver3d as S
    old = S.getFrictionConstant()
    S.setFrictionConstant(0.0)
    assert S.getFrictionConstant() == 0.0
    S.setFrictionConstant(old)
    assert (S.getRoundingTolerance(), S.getFabricThickness(), S.getSpringConstant(), S.getPointMass()) == (1e-6, 1e-4, 1000.0, 0.01)


def test_crmath_correctly_rounded_on_host():
    """crmath.cuh (the kernels' acos/cos/pow) compiled for the host and compared with binary128."""
    exe = "/tmp/clsn_crmath_check"
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-mfma", "-x", "c++", "-I", os.path.join(ROOT, "collision_b200", "csrc"),
                           os.path.join(ROOT, "tests", "crmath_check.cpp"), "-o", exe, "-lquadmath", "-lm"])
    out = subprocess.check_output([exe, "300000"], text=True)
    for line in out.strip().splitlines():
        name, n, bad = line.split()
        assert int(bad) == 0, f"{name}: {bad} of {n} results are not correctly rounded"


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
from collision_b200.dist import gather_varlen
rank = int(os.environ["RANK"])
dist.init_process_group("gloo")
n = [5 * 64, 0, 3 * 64][rank]
local = (torch.arange(n, dtype=torch.int64) %% 251 + rank).to(torch.uint8)
out = gather_varlen(local)
exp = torch.cat([(torch.arange(m, dtype=torch.int64) %% 251 + r).to(torch.uint8) for r, m in enumerate([5 * 64, 0, 3 * 64])])
assert torch.equal(out, exp), (rank, out.shape)
empty = gather_varlen(torch.empty(0, dtype=torch.uint8))
assert empty.numel() == 0
# owner-computes all-to-all: rank r sends (r+1)*(d+1)*64 bytes of value 10*r+d to rank d
from collision_b200.dist import exchange_by_owner
counts = [(rank + 1) * (d + 1) * 64 for d in range(3)]
send = torch.cat([torch.full((n,), 10 * rank + d, dtype=torch.uint8) for d, n in enumerate(counts)])
recv, rc = exchange_by_owner(send, counts)
assert rc == [(s + 1) * (rank + 1) * 64 for s in range(3)], rc
exp = torch.cat([torch.full(((s + 1) * (rank + 1) * 64,), 10 * s + rank, dtype=torch.uint8) for s in range(3)])
assert torch.equal(recv, exp)
dist.destroy_process_group()
print("ok", rank)
"""


def test_record_exchange_gloo_world3():
    """the N > 1 exchange (variable-length all-gather of record buffers) on CPU tensors with gloo"""
    code = _WORKER % ROOT
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=3", "--master-addr", "127.0.0.1",
           "--master-port", "29631", "--no-python", sys.executable, "-c", code]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 3


def test_cubic_path_matches_oracle_on_host():
    """collision_b200/csrc/cubic.cuh -- the CUDA path's isCoplanar (coefficients, trig-free classifier,
    selective correctly rounded solve) -- compiled for the host and fuzzed against the oracle's
    is_coplanar (binary128 libm flavour): identical return value and root bits on every case, and no
    cubic that the classifier rejects has a root the oracle keeps.  (48 M cases were run once by hand.)"""
    from oracle import port
    so = port.build()
    exe = "/tmp/clsn_cubic_check"
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-mfma", "-x", "c++", "-I", os.path.join(ROOT, "collision_b200", "csrc"),
                           os.path.join(ROOT, "tests", "cubic_check.cpp"), "-o", exe, "-ldl", "-lm"])
    n, bad, rejected, rejected_wrong, oracle_true, pre, pre_wrong = map(
        int, subprocess.check_output([exe, so, "2000000", "7"], text=True).split())
    assert n == 2000000 and bad == 0 and rejected_wrong == 0
    assert pre > n // 8 and pre_wrong == 0   # the (opt-in) FP32 pre-filter only rejects what the oracle rejects
    assert rejected > n // 4 and oracle_true > n // 4   # both outcomes are well represented


def test_normal_axis_separation_test_is_conservative_on_host():
    """collision_b200/csrc/cubic.cuh: sat_normal_far() -- the FP32 separating-axis test along the feature's own normal that
    k_cull puts in front of the proximity narrow phase -- fuzzed on the host against the oracle's PointToTri / EdgeToEdge
    (static) and MovingPointToTri / MovingEdgeToEdge (moving): it may only reject what the oracle does not report.
    Half of the cases sit within a few contact distances of the boundary, slivers and rescaled elements included.
    (30 M cases were run once by hand: 0 wrong.)"""
    from oracle import port
    so = port.build()
    exe = "/tmp/clsn_sat_check"
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-mfma", "-x", "c++", "-I", os.path.join(ROOT, "collision_b200", "csrc"),
                           os.path.join(ROOT, "tests", "sat_check.cpp"), "-o", exe, "-ldl", "-lm"])
    n, rej_s, wrong_s, hit_s, rej_m, wrong_m, hit_m = map(int, subprocess.check_output([exe, so, "1500000", "21"], text=True).split())
    assert n == 1500000 and wrong_s == 0 and wrong_m == 0
    assert rej_s > n // 3 and rej_m > n // 3 and hit_s > n // 20 and hit_m > n // 20   # both outcomes well represented


def _build_fastpath_check():
    from oracle import port
    so = port.build()
    exe, lib = "/tmp/clsn_fastpath_check", "/tmp/libclsn_fastpath_check.so"
    base = ["g++", "-O2", "-ffp-contract=off", "-mfma", "-x", "c++", "-I", os.path.join(ROOT, "collision_b200", "csrc"),
            os.path.join(ROOT, "tests", "fastpath_check.cpp")]
    subprocess.check_call(base + ["-o", exe, "-ldl", "-lm"])
    subprocess.check_call(base + ["-fPIC", "-shared", "-o", lib, "-ldl", "-lm"])
    return so, exe, lib


def test_ccd_fast_path_is_conservative_on_host():
    """collision_b200/csrc/fastpath.cuh -- the plain-FP64 fast path of the fused CCD feature kernel -- compiled
    for the host with a libm perturbed by up to +-4 ulp per call, fuzzed against the oracle's
    MovingPointToTri / MovingEdgeToEdge (correctly rounded libm): FAST_MISS only where the oracle's isCoplanar is
    false, FAST_DT_ONLY only where it is true and nothing fires before t = dt.  Half of the cases sit on the
    accept/reject boundary of the static tests.  (190 M cases were run once by hand.)"""
    so, exe, _ = _build_fastpath_check()
    n, wrong, n_miss, n_dt, n_unc, hits, hits_root = map(int, subprocess.check_output([exe, so, "1500000", "3"], text=True).split())
    assert n == 1500000 and wrong == 0
    assert n_miss > n // 10 and n_dt > n // 5 and hits_root > n // 5   # every outcome is well represented
    assert n_unc < hits_root + n // 5                                  # ... and the fast path settles most non-hits


def test_ccd_fast_path_on_oracle_scenes():
    """The same check on real states: every CCD feature test of every candidate pair of every pass of a few
    oracle-driven steps (cloth stacks, fixed and movable rigid bodies, strings)."""
    import ctypes as C
    from collision_b200 import scenes
    from oracle import port
    so, _, libp = _build_fastpath_check()
    L = C.CDLL(libp)
    L.fastpath_scene_check.restype = C.c_long
    L.fastpath_scene_check.argtypes = [C.c_char_p, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                       C.c_double, C.c_void_p]
    port.set_libm(port.LIBM_CR)
    try:
        for sc, steps in ((scenes.layered_cloth(3, 16, seed=3), 2), (scenes.two_sheets(), 2), (scenes.mixed(), 2),
                          (scenes.ball_plane(level=2), 2), (scenes.string_string(), 20), (scenes.drape(n=24, level=2), 2)):
            elem = np.full((sc.T + sc.B, 4), -1, np.int32)
            elem[:sc.T, :3] = sc.tri_idx
            elem[sc.T:, :2] = sc.bond_idx
            orc = port.OracleSolver(sc, impact_zones=False, strain_limiting=False)
            x, vel = sc.x.copy(), sc.vel.copy()
            tot = np.zeros(6, np.int64)
            for _ in range(steps):
                orc.set_state(x, x + sc.dt * vel)
                orc.avg_velocity()
                orc.detect(port.PROXIMITY)
                orc.apply()
                for _it in range(5):
                    xo = np.ascontiguousarray(orc.get(port.F_X_OLD))
                    av = np.ascontiguousarray(orc.get(port.F_AVGVEL))
                    n_true = orc.detect(port.COLLISION)
                    cand = np.ascontiguousarray(orc.candidates())
                    counts = np.zeros(6, np.int64)   # miss, dt_only, uncertain, oracle hits, classifier / pre-filter rejections
                    wrong = L.fastpath_scene_check(so.encode(), len(cand), cand.ctypes.data, elem.ctypes.data, xo.ctypes.data,
                                                   av.ctypes.data, sc.dt, sc.params.eps, counts.ctypes.data)
                    assert wrong == 0, (sc.name, wrong)
                    tot += counts
                    orc.apply()
                    if n_true == 0:
                        break
                orc.boundary()
                orc.final_position()
                v = vel.copy()
                orc.final_velocity(v)
                x, vel = orc.get(port.F_X).copy(), v
            # uncertain = features that do fire at a root + a thin boundary layer
            assert tot[2] <= 2 * tot[3] + 50, (sc.name, tot)
            assert tot[5] <= tot[4]   # the FP32 pre-filter is weaker than the FP64 classifier
            orc.close()
    finally:
        port.set_libm(port.LIBM_NATIVE)


def test_feature_slot_arithmetic_matches_the_tables():
    """k_cull computes the local point slots of a feature test arithmetically (feature_slots) instead of indexing
    the __constant__ tables that document the reference's loop order (MovingTriToTri :286-323, TriToTri :592-625,
    TriToBond / MovingTriToBond): both are extracted from clsn.cu and compared on the host."""
    src = open(os.path.join(ROOT, "collision_b200", "csrc", "clsn.cu")).read()

    def table(name):
        m = re.search(name + r"\[\d+\]\[4\] = \{(.*?)\};", src, re.S)
        return [list(map(int, re.findall(r"\d+", row))) for row in re.findall(r"\{([^{}]*)\}", m.group(1))]

    tm, ts, tb = table("c_feat_tt_moving"), table("c_feat_tt_static"), table("c_feat_tb")
    assert (len(tm), len(ts), len(tb)) == (15, 15, 5)
    a = src.index("template <bool MOVING>\n__host__ __device__ __forceinline__ void feature_slots")
    fn = src[a:src.index("#define CULL_THREADS 128")].replace("__host__ __device__ __forceinline__", "static inline")
    calls = []
    for mv, typ, n in (("true", 0, 15), ("false", 0, 15), ("true", 1, 5), ("false", 1, 5), ("true", 2, 1)):
        calls += [f'feature_slots<{mv}>({typ},{f},sl); printf("%d %d %d %d\\n",sl[0],sl[1],sl[2],sl[3]);' for f in range(n)]
    prog = "#include <cstdio>\n" + fn + "\nint main(){int sl[4];\n" + "\n".join(calls) + "\nreturn 0;}\n"
    with open("/tmp/clsn_slots.cpp", "w") as f:
        f.write(prog)
    subprocess.check_call(["g++", "-O1", "/tmp/clsn_slots.cpp", "-o", "/tmp/clsn_slots"])
    out = [list(map(int, ln.split())) for ln in subprocess.check_output(["/tmp/clsn_slots"], text=True).strip().splitlines()]
    assert out == tm + ts + tb + tb + [[0, 1, 3, 4]]


def test_vtk_dump_mirrors_the_reference_layout(tmp_path):
    """collision_b200/vtkdump.py (vtk.cpp:14-93): points numbered by first appearance in
    hseList, CollsnImpulse point vectors, cells red where a point collected an impulse -- parsed back from the file."""
    import xml.etree.ElementTree as ET
    from collision_b200 import scenes, vtkdump
    sc = scenes.mixed()
    rng = np.random.default_rng(5)
    imp = np.zeros((sc.V, 3))
    cnt = np.zeros(sc.V, np.int32)
    hit = rng.choice(sc.V, 40, replace=False)
    imp[hit] = rng.normal(size=(40, 3))
    cnt[hit] = rng.integers(1, 5, 40)
    f = str(tmp_path / "collsn.vtp")
    info = vtkdump.vtkplotVectorSurface(f, sc.x, sc.tri_idx, sc.bond_idx, imp, cnt)
    order = vtkdump.point_order(sc.tri_idx, sc.bond_idx)
    assert order[:3].tolist() == sc.tri_idx[0].tolist() and len(set(order.tolist())) == order.size
    assert info == {"points": order.size, "cells": sc.T + sc.B}
    piece = ET.parse(f).getroot().find("PolyData/Piece")
    assert int(piece.get("NumberOfPolys")) == sc.T and int(piece.get("NumberOfLines")) == sc.B
    arr = lambda e: np.array(e.text.split(), dtype=np.float64)
    pts = arr(piece.find("Points/DataArray")).reshape(-1, 3)
    assert np.allclose(pts, sc.x[order], rtol=1e-7)
    vec = arr(piece.find("PointData/DataArray")).reshape(-1, 3)
    assert np.allclose(vec, imp[order], rtol=1e-7)
    conn = arr(piece.find("Polys/DataArray[@Name='connectivity']")).astype(int).reshape(-1, 3)
    assert np.array_equal(order[conn], sc.tri_idx)
    col = arr(piece.find("CellData/DataArray")).astype(int).reshape(-1, 3)
    red_tris = (cnt[sc.tri_idx] > 0).any(axis=1)
    assert np.array_equal(col[sc.B:, 0] == 255, red_tris) and np.array_equal(col[sc.B:, 1] == 255, ~red_tris)


def test_bench_reference_arm_line(tmp_path):
    """`bench.py --impl reference` (here on a bounded sample, so that the CPU suite stays short): one JSON line on stdout with
    the contract's keys, the reference's own library as the thing timed, nothing from the CUDA path loaded."""
    import json
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built here")
    env = dict(os.environ, TMPDIR=str(tmp_path))   # the per-box cache of the arm goes to the temp dir
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--sample-reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "ccd_pairs_per_sec" and d["unit"] == "pairs/s" and d["higher_is_better"] is True
    assert d["steps"] == 1 and d["warmup"] == 0 and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] == 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    # a second run on the same "box" reuses the measurement and says so
    r2 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--sample-reference", "--steps", "1",
                         "--warmup", "0", "--gpus", "2"], capture_output=True, text=True, timeout=600, env=env)
    d2 = json.loads([ln for ln in r2.stdout.splitlines() if ln.strip()][0])
    assert d2["value"] == d["value"] and "reused" in d2["cpu_baseline"]["sample"] and d2["n_gpus"] == 2


def test_cpp_mirror_links_against_the_c_abi_and_has_no_cpu_path(tmp_path):
    """collision_b200/host (the C++ mirror of collid.h + its test.cpp-like driver) compiles against include/collision_b200.h and
    links against the library; without a CUDA device it stops with an error instead of computing anything on the host."""
    import subprocess
    import torch
    from collision_b200 import scenes
    from test_gpu_host_cpp import HOST, write_scene
    _lib_path()
    subprocess.check_call(["make", "-s", "-C", HOST])
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present: the GPU suite runs the driver for real")
    sc = scenes.two_sheets(n=6)
    inp, out = str(tmp_path / "scene.bin"), str(tmp_path / "out.bin")
    write_scene(sc, inp)
    r = subprocess.run([os.path.join(HOST, "host_check"), inp, out, "1"], capture_output=True, text=True)
    assert r.returncode == 3 and "no usable CUDA device" in r.stderr
    assert not os.path.exists(out)
