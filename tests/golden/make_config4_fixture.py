"""Full-size pin of config 4 (layered_cloth 8x251^2, 1 000 000 triangles) against THE REFERENCE ITSELF:
oracle/_ref/libcollision_ref.so (= /root/reference/{AABB,dcollid,dcollid3d}.cpp compiled unmodified) is run
once on the benchmark scene and the per-pass integer results are committed as a small fixture
(tests/golden/config4_reference.npz, checked by tests/test_gpu_fullsize.py on the B200).

    python tests/golden/make_config4_fixture.py [layers n]     # ~10 min and ~2 GB at the default size

Two kinds of records:

* "pinned" passes -- run from inputs that are a pure function of the scene arrays, so that the CUDA path can be
  started from IDENTICAL inputs without shipping 12 MB of state:
      P   proximity pass   from avgVel0 = (x_new - x_old) / dt
      C0  CCD pass         from avgVel0            (the proximity impulses are discarded: avgVel is reset)
      C1  CCD pass         from 0.8 * avgVel0      (a second, different state)
  for each: number of callbacks (candidates), tree count (true pairs), sha256 of the sorted candidate set and of
  the sorted true-pair set, sha256 of the per-point contribution counts (collsn_num) and of has_collsn after
  updateAverageVelocity.  All integer work: the CUDA path must reproduce them exactly.
* "natural" passes -- the reference's own resolveCollision sequence (proximity, apply, CCD x <= 5): per-pass
  callbacks and tree counts.  From the second CCD pass on the reference's state depends on its own summation
  order, so these are compared with a stated tolerance, not bit for bit.
"""
import hashlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from collision_b200 import scenes  # noqa: E402
from oracle import ref  # noqa: E402


def sorted_pairs(p):
    q = np.sort(np.asarray(p, dtype=np.int32).reshape(-1, 2), axis=1)
    return np.ascontiguousarray(q[np.lexsort((q[:, 1], q[:, 0]))])


def digest(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def pinned_pass(r, sc, phase, av_in, out, tag):
    """one detection pass of the reference from avgVel = av_in; accumulators start from zero"""
    V = sc.V
    r.put(ref.F_AVGVEL, av_in)
    r.put(ref.F_IMP, np.zeros((V, 3)))
    r.put(ref.F_FRIC, np.zeros((V, 3)))
    r.puti(ref.I_CNT, np.zeros(V, np.int32))
    r.puti(ref.I_HAS_COLLSN, np.zeros(V, np.int32))
    r.record(True)
    t0 = time.perf_counter()
    n_true = r.phase(phase)
    t1 = time.perf_counter()
    pr = r.pairs()
    cand = sorted_pairs(pr[:, :2])
    true = sorted_pairs(pr[pr[:, 2] != 0, :2])
    cnt = r.geti(ref.I_CNT)
    r.phase(ref.PH_APPLY)
    has = r.geti(ref.I_HAS_COLLSN)
    r.record(False)
    out[tag + "_candidates"] = np.int64(len(cand))
    out[tag + "_true_pairs"] = np.int64(n_true)
    out[tag + "_cand_sha"] = digest(cand)
    out[tag + "_true_sha"] = digest(true)
    out[tag + "_cnt_sha"] = digest(cnt.astype(np.int32))
    out[tag + "_cnt_total"] = np.int64(cnt.sum())
    # the arrays themselves (compressed: a few hundred KB), so that a mismatch can be located and counted -- a verdict of a
    # borderline edge-edge test may depend on the order in which the reference's tree hands the pair over (ref_compare.py)
    assert cnt.max() < 65536
    out[tag + "_cnt"] = cnt.astype(np.uint16)
    out[tag + "_has"] = np.packbits(has != 0)
    out[tag + "_has_sha"] = digest((has != 0).astype(np.uint8))
    out[tag + "_has_total"] = np.int64((has != 0).sum())
    out[tag + "_seconds"] = np.float64(t1 - t0)
    assert len(true) == n_true
    print(f"{tag}: {len(cand)} candidates, {n_true} true pairs, {int(cnt.sum())} contributions, "
          f"{int((has != 0).sum())} points hit, {t1 - t0:.1f} s", flush=True)


def main():
    layers, n = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (8, 251)
    sc = scenes.layered_cloth(layers, n)
    name = "config4_reference.npz" if (layers, n) == (8, 251) else f"config4_reference_{layers}x{n}.npz"
    r = ref.RefSolver(sc)
    x = sc.x.copy()
    xn = sc.x_new()
    out = {"layers": np.int32(layers), "n": np.int32(n), "T": np.int64(sc.T), "V": np.int64(sc.V)}
    # ---- natural sequence (the reference's own resolveCollision order, dcollid.cpp:317-362 without the tail stages)
    r.set_state(x, xn, sc.vel)
    r.assemble(sc.dt)
    r.record(False)
    t0 = time.perf_counter()
    r.phase(ref.PH_AVG_VELOCITY)
    av0 = r.get(ref.F_AVGVEL)
    n0 = r.num_callbacks()
    nat_true = [r.phase(ref.PH_PROXIMITY_DETECT)]
    r.phase(ref.PH_APPLY)
    nat_cand = [r.num_callbacks() - n0]
    for _ in range(5):
        n0 = r.num_callbacks()
        nat_true.append(r.phase(ref.PH_COLLISION_DETECT))
        r.phase(ref.PH_APPLY)
        nat_cand.append(r.num_callbacks() - n0)
        print(f"natural pass {len(nat_true) - 1}: {nat_cand[-1]} candidates, {nat_true[-1]} true pairs", flush=True)
        if nat_true[-1] == 0:
            break
    r.phase(ref.PH_BOUNDARY)
    r.phase(ref.PH_FINAL_POSITION)
    r.phase(ref.PH_FINAL_VELOCITY)
    t1 = time.perf_counter()
    has = r.geti(ref.I_HAS_COLLSN)
    out["natural_candidates"] = np.array(nat_cand, np.int64)
    out["natural_true_pairs"] = np.array(nat_true, np.int64)
    out["natural_has_total"] = np.int64((has != 0).sum())
    out["natural_seconds"] = np.float64(t1 - t0)
    print(f"natural step: {t1 - t0:.1f} s on one core", flush=True)
    # ---- pinned passes from inputs that are pure functions of the scene
    r.set_state(x, xn, sc.vel)
    r.assemble(sc.dt)
    r.phase(ref.PH_AVG_VELOCITY)
    assert np.array_equal(r.get(ref.F_AVGVEL).view(np.uint64), av0.view(np.uint64))
    pinned_pass(r, sc, ref.PH_PROXIMITY_DETECT, av0, out, "P")
    pinned_pass(r, sc, ref.PH_COLLISION_DETECT, av0, out, "C0")
    pinned_pass(r, sc, ref.PH_COLLISION_DETECT, 0.8 * av0, out, "C1")
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    if not ref.available():
        raise SystemExit("oracle/_ref/libcollision_ref.so missing: run `make -C oracle ref` (needs /root/reference)")
    main()
