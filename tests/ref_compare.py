"""Direct comparison against THE REFERENCE ITSELF (oracle/_ref/libcollision_ref*.so = the unmodified
/root/reference/{AABB,dcollid,dcollid3d}.cpp), pass by pass from IDENTICAL inputs.

The implementation under test ("impl") is either the CUDA path (tests/test_gpu_vs_reference.py, -m gpu) or the C
restatement in its canonical-order / correctly-rounded flavour (tests/test_oracle_vs_reference.py, CPU): the same
checks, so the CPU run validates the checker and shows the numbers the GPU run must reproduce.

Two builds of the reference are used (oracle/Makefile): the native-libm one (exactly what the reference is on this
host) and one with a correctly rounded acos/cos/sin/pow bound in (oracle/cr_libm.c) -- "the reference on a platform
with a correctly rounded libm", which is the contract of the CUDA path.

After every detection pass (reference: AABB.cpp:254-343 -> dcollid.cpp:753-836 -> dcollid3d.cpp:203-369):
  A. candidate set  = the reference tree's narrow-phase callbacks as sorted (a<b) pairs            -> bit-equal
  B. canonical-order verdicts: EVERY candidate pair is re-run feature by feature through the reference's own
     MovingPointToTri / MovingEdgeToEdge / PointToTri / EdgeToEdge with the pair taken as (min, max) -- the order the
     implementation is defined for (clsn_ref_feature calls the file-static functions of dcollid3d.cpp):
        contact set (ea, eb, feature)                                                              -> equal
        time of impact of every CCD contact          CR build: bit-equal;  native build: <= 1e-8 relative and >= 99 % within 1e-12
  C. the reference's own run (its tree hands pairs over in either order):
        true-pair set (callbacks that returned true) -> equal up to ORDER_FLIP_FRACTION of the pairs: an edge-edge test
            called as (b, a) instead of (a, b) is a different floating-point evaluation of the same geometry, and a
            borderline verdict can flip -- the reference disagrees with ITSELF there (measured: column `order_flips`)
        collsn_num per point                         -> equal (outside flipped pairs)
        per-point impulse sums, point-triangle only  -> <= 1e-12 relative (summation order)
        per-point impulse sums with edge-edge        -> stated bound, observed max reported (the edge-edge normal at a
            coplanarity root is a difference of nearly equal points, dcollid3d.cpp:729-744: O(1) sensitive to the last
            bit of the root, hence to the call order)
After the pass the reference's avgVel (and body accumulators) are copied into the implementation, so the next pass starts
from identical inputs again.
"""
from __future__ import annotations

import numpy as np

from oracle import ref

TOI_RTOL = 1e-12            # north star: times of impact within 1e-12 relative
TOI_RTOL_NATIVE_LIBM = 1e-8  # against the native-libm build: glibc's <= 1 ulp acos/cos/pow on the cancelling trig root form
PT_SUM_RTOL = 1e-12         # per-point impulse sums, points touched by point-triangle contacts only (summation order)
ORDER_FLIP_FRACTION = 0.02  # share of true pairs whose verdict may depend on the reference's own call order (edge-edge)


def sort_pairs(p):
    p = np.asarray(p).reshape(-1, 2)
    if len(p) == 0:
        return p.astype(np.int32)
    q = np.sort(p, axis=1)
    return q[np.lexsort((q[:, 1], q[:, 0]))].astype(np.int32)


def elem_points(sc, e):
    return sc.tri_idx[e] if e < sc.T else sc.bond_idx[e - sc.T]


def reference_features(sc, a, b, moving):
    """Feature tests of one callback isCollision/isProximity(a, b) in the reference's loop order, as
    (canonical feature index for the pair (min, max), is_edge, points[4], swapped).  dcollid3d.cpp:203-325, 485-627."""
    T = sc.T
    A, B = elem_points(sc, a), elem_points(sc, b)
    out = []
    swapped = a > b
    if a < T and b < T:
        # CCD (MovingTriToTri :286-303): k=0 tri a + vertex of b, k=1 tri b + vertex of a
        # proximity (TriToTri :592-610): k=0 tri b + vertex of a, k=1 tri a + vertex of b
        order = ((A, B), (B, A)) if moving else ((B, A), (A, B))
        for k, (tri, oth) in enumerate(order):
            for i in range(3):
                f_here = 3 * k + i                       # index in the call as made
                f = f_here if not swapped else (f_here + 3) % 6   # same feature seen from the canonical pair
                out.append((f, False, [tri[0], tri[1], tri[2], oth[i]], swapped))
        for i in range(3):
            for j in range(3):
                f = 6 + 3 * i + j if not swapped else 6 + 3 * j + i
                out.append((f, True, [A[i], A[(i + 1) % 3], B[j], B[(j + 1) % 3]], swapped))
    elif a < T or b < T:
        # the dispatcher always calls (Moving)TriToBond(tri, bond) (dcollid.cpp:775-786, 818-829)
        tri, bond = (A, B) if a < T else (B, A)
        for i in range(2):
            out.append((i, False, [tri[0], tri[1], tri[2], bond[i]], False))
        for i in range(3):
            out.append((2 + i, True, [tri[i], tri[(i + 1) % 3], bond[0], bond[1]], False))
    else:
        out.append((0, True, [A[0], A[1], B[0], B[1]], swapped))
    return out


def shares_vertex(A, B):
    return bool(set(int(v) for v in A) & set(int(v) for v in B))


def canonical_verdicts(sc, cand, x_old, avgvel, moving):
    """The reference's own primitives on every candidate pair taken as (min, max): dict (ea, eb, f) -> dict(toi, edge, pts).
    Pair-level early returns restated from the drivers: shared vertex (dcollid3d.cpp:209-214, 257-264, 279-284, 491-496,
    546-553, 574-579) and same-surface pairs with a rigid first element (dcollid.cpp:762, 805)."""
    p = sc.params
    params = np.array([p.eps, p.thickness, p.k, p.m, p.friction, p.cr])
    h = p.eps if moving else p.thickness
    T = sc.T
    keys, kinds, quads = [], [], []
    for a, b in cand:
        a, b = int(a), int(b)
        A, B = elem_points(sc, a), elem_points(sc, b)
        if shares_vertex(A, B):
            continue
        if a < T and b < T and sc.tri_surf[a] == sc.tri_surf[b] and (sc.vflags[A] & 3).any():
            continue
        for f, edge, pts, _ in reference_features(sc, a, b, moving):
            keys.append((a, b, f, edge))
            kinds.append((ref.K_MOVING_EDGE_TO_EDGE if edge else ref.K_MOVING_POINT_TO_TRI) if moving else
                         (ref.K_EDGE_TO_EDGE if edge else ref.K_POINT_TO_TRI))
            quads.append(pts)
    ret, hit = ref.feature_batch(kinds, np.asarray(quads, np.int32).reshape(-1, 4), x_old, avgvel, sc.vflags,
                                 sc.hs_mass[sc.vhs], h, sc.dt, params)
    out = {}
    for i in np.nonzero(ret)[0]:
        a, b, f, edge = keys[i]
        out[(a, b, f)] = dict(toi=float(hit[i]) if moving else 0.0, edge=edge, pts=np.asarray(quads[i]))
    return out


class Impl:
    """what the comparison needs from the implementation under test"""
    def upload(self, x_old, x_new): ...
    def avg_velocity(self): ...
    def avgvel(self): ...
    def set_avgvel(self, av): ...
    def set_body(self, imp_rg_per_body, cnt_rg_per_body): ...
    def detect(self, moving): ...          # -> dict(candidates=(n,2) sorted, contacts=structured array, cnt, imp, fric)
    def apply(self): ...
    def has_collsn(self): ...


def _bump(report, key, v):
    report[key] = max(report.get(key, 0.0), float(v))


def compare_pass(sc, r, impl, moving, report, cr_libm):
    """one detection pass on both sides from their current, identical state; then sync the implementation to the
    reference's post-apply state.  `report` collects counts and observed maxima."""
    x_old = r.get(ref.F_X_OLD)
    av_in = r.get(ref.F_AVGVEL)
    assert np.array_equal(impl.avgvel().view(np.uint64), av_in.view(np.uint64)), "inputs of the pass are not identical"
    r.record(True)
    n_true = r.phase(ref.PH_COLLISION_DETECT if moving else ref.PH_PROXIMITY_DETECT)
    pairs = r.pairs()
    r.record(False)
    got = impl.detect(moving)
    con = got["contacts"]
    # ---- A. candidate set
    cand_ref = sort_pairs(pairs[:, :2])
    assert np.array_equal(got["candidates"], cand_ref), f"candidate sets differ ({len(got['candidates'])} vs {len(cand_ref)})"
    report["candidates"] = report.get("candidates", 0) + len(cand_ref)
    # ---- B. canonical-order verdicts of the reference's own primitives
    rc = canonical_verdicts(sc, cand_ref, x_old, av_in, moving)
    keys_got = {(int(c["ea"]), int(c["eb"]), int(c["feature"])) for c in con}
    assert keys_got == set(rc), f"contact sets differ: {len(keys_got)} vs {len(rc)}, e.g. {sorted(keys_got ^ set(rc))[:4]}"
    ee_points = set()
    for c in con:
        k = (int(c["ea"]), int(c["eb"]), int(c["feature"]))
        q = rc[k]
        assert bool(c["kind"]) == q["edge"] and np.array_equal(c["p"], q["pts"])
        if q["edge"]:
            ee_points.update(int(v) for v in q["pts"])
        if moving:
            rel = abs(float(c["root"]) - q["toi"]) / max(abs(q["toi"]), 1e-300)
            _bump(report, "toi_rel_max", rel)
            report["toi"] = report.get("toi", 0) + 1
            report["toi_bit_equal"] = report.get("toi_bit_equal", 0) + (1 if float(c["root"]) == q["toi"] else 0)
            report["toi_le_1e-12"] = report.get("toi_le_1e-12", 0) + (1 if rel <= TOI_RTOL else 0)
            if cr_libm:
                assert float(c["root"]) == q["toi"], f"time of impact of {k}: {c['root']!r} vs {q['toi']!r} (rel {rel:.2e})"
            else:
                assert rel <= TOI_RTOL_NATIVE_LIBM, f"time of impact of {k}: {c['root']!r} vs {q['toi']!r} (rel {rel:.2e})"
    report["contacts"] = report.get("contacts", 0) + len(con)
    report["ee_contacts"] = report.get("ee_contacts", 0) + sum(1 for q in rc.values() if q["edge"])
    # ---- C. the reference's own run, pairs in the order its tree produced them
    true_ref = {(min(int(a), int(b)), max(int(a), int(b))) for a, b, res in pairs if res}
    assert len(true_ref) == n_true
    true_got = {(k[0], k[1]) for k in keys_got}
    flips = true_ref ^ true_got
    swapped_calls = {(min(int(a), int(b)), max(int(a), int(b))) for a, b, _ in pairs if a > b}
    assert flips <= swapped_calls, f"true-pair sets differ on pairs the reference evaluated in canonical order: {sorted(flips - swapped_calls)[:4]}"
    report["true_pairs"] = report.get("true_pairs", 0) + len(true_ref)
    report["order_flips"] = report.get("order_flips", 0) + len(flips)
    assert len(flips) <= max(1, int(ORDER_FLIP_FRACTION * len(true_ref))), f"{len(flips)} of {len(true_ref)} true pairs flip with the call order"
    flip_pts = np.zeros(sc.V, bool)
    for a, b in flips:
        flip_pts[elem_points(sc, a)] = True
        flip_pts[elem_points(sc, b)] = True
    # the verdict of a swapped edge-edge feature can flip inside a pair that stays true: then collsn_num moves as well
    cnt_ref = r.geti(ref.I_CNT)
    cnt_bad = (got["cnt"] != cnt_ref) & ~flip_pts
    report["collsn_num_mismatch_points"] = report.get("collsn_num_mismatch_points", 0) + int(cnt_bad.sum())
    ee_mask = np.zeros(sc.V, bool)
    if ee_points:
        ee_mask[list(ee_points)] = True
    assert not (cnt_bad & ~ee_mask).any(), "per-point collsn_num differs on points without an edge-edge contact"
    assert cnt_bad.sum() <= max(2, int(ORDER_FLIP_FRACTION * (cnt_ref > 0).sum())), "too many collsn_num differences"
    # ---- per-point sums against the reference's accumulators
    tot_ref = r.get(ref.F_IMP) + r.get(ref.F_FRIC)
    tot_got = got["imp"] + got["fric"]
    scale = max(np.abs(tot_ref).max(), 1e-300)
    d = np.abs(tot_got - tot_ref).max(axis=1)
    pt_only = ~ee_mask & ~flip_pts
    pt_err = d[pt_only].max() / scale if pt_only.any() else 0.0
    ee_err = d[ee_mask].max() / scale if ee_mask.any() else 0.0
    _bump(report, "pt_sum_rel_max", pt_err)
    _bump(report, "ee_sum_rel_max", ee_err)
    assert pt_err <= PT_SUM_RTOL, f"impulse sums of point-triangle-only points differ by {pt_err:.2e} (relative to {scale:.3e})"
    # ---- updateAverageVelocity on both sides
    r.phase(ref.PH_APPLY)
    impl.apply()
    has_ref = r.geti(ref.I_HAS_COLLSN) != 0
    has_bad = ((impl.has_collsn() != 0) != has_ref)
    excl = report.setdefault("_has_excluded", np.zeros(sc.V, bool))   # has_collsn is cumulative over the step
    excl |= flip_pts | cnt_bad
    assert not (has_bad & ~excl).any(), "has_collsn differs"
    av_ref = r.get(ref.F_AVGVEL)
    dv = np.abs(impl.avgvel() - av_ref).max(axis=1)
    vs = max(np.abs(av_ref).max(), 1e-300)
    _bump(report, "avgvel_rel_max_pt", dv[pt_only].max() / vs if pt_only.any() else 0.0)
    _bump(report, "avgvel_rel_max_ee", dv[ee_mask].max() / vs if ee_mask.any() else 0.0)
    # ---- next pass from identical inputs again
    impl.set_avgvel(av_ref)
    irg = r.get(ref.F_IMP_RG)
    crg = r.geti(ref.I_CNT_RG)
    nb = len(sc.hs_mass)
    b_imp, b_cnt = np.zeros((nb, 3)), np.zeros(nb, np.int32)
    first = np.full(nb, -1)
    for v in range(sc.V - 1, -1, -1):
        first[sc.vhs[v]] = v
    for b in range(nb):
        if first[b] >= 0:
            b_imp[b], b_cnt[b] = irg[first[b]], crg[first[b]]
    impl.set_body(b_imp, b_cnt)
    return n_true


def run_steps(sc, impl, n_steps=2, max_passes=5, cr_libm=True):
    """resolveCollision's pass sequence (dcollid.cpp:317-362) on the reference, the implementation following from
    identical inputs at every pass.  Returns the report of observed maxima."""
    ref.set_libm(ref.LIBM_CR if cr_libm else ref.LIBM_NATIVE)
    r = ref.RefSolver(sc)
    x, vel = sc.x.copy(), sc.vel.copy()
    report = {}
    for step in range(n_steps):
        xn = x + sc.dt * vel
        r.set_state(x, xn, vel)
        r.assemble(sc.dt)
        r.puti(ref.I_HAS_COLLSN, np.zeros(sc.V, np.int32))   # recordOriginPosition (dcollid.cpp:100)
        impl.upload(x, xn)
        report.pop("_has_excluded", None)
        r.phase(ref.PH_AVG_VELOCITY)
        impl.avg_velocity()
        compare_pass(sc, r, impl, False, report, cr_libm)
        for _ in range(max_passes):
            if compare_pass(sc, r, impl, True, report, cr_libm) == 0:
                break
        r.phase(ref.PH_BOUNDARY)
        r.phase(ref.PH_FINAL_POSITION)
        r.phase(ref.PH_FINAL_VELOCITY)
        x, vel = r.get(ref.F_COORDS), r.get(ref.F_VEL)
    r.close()
    ref.set_libm(ref.LIBM_NATIVE)
    report.pop("_has_excluded", None)
    return report
