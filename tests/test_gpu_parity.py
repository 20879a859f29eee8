"""GPU parity tests proper: the CUDA path, called through the C ABI, against the C oracle
(correctly rounded libm flavour) on the same seeded inputs -- bit for bit.

Integer/index work (candidate sets, contact sets, counts, has_collsn) must be identical; FP64
results (times of impact, normals, weights, per-point impulse sums, avgVel, final positions) are
compared BITWISE, which is stronger than the 1e-12 relative tolerance the north star allows.
"""
import os

import numpy as np
import pytest

from collision_b200 import scenes
from collision_b200.solver import CollisionSolver3d
from oracle import port
from parity_util import run_step_by_phases, same_bits

pytestmark = pytest.mark.gpu

SCENES = {
    "string_string": lambda: scenes.string_string(dt=0.01, gap=0.003),
    "two_sheets": lambda: scenes.two_sheets(n=20),
    "two_sheets_nofric": lambda: scenes.two_sheets(n=16, friction=0.0, seed=5),
    "mixed": lambda: scenes.mixed(),
    "ball_plane": lambda: scenes.ball_plane(gap=2e-4),
    "box_boundary": lambda: scenes.box_boundary(),
    "sheet_wall": lambda: scenes.sheet_wall(),
    "drape_small": lambda: scenes.drape(n=40, level=3),
    "layered_4x24": lambda: scenes.layered_cloth(4, 24, seed=99),
    # config-5 family: cloth stack + fast movable rigid spheres (per-body accumulators, rigidification, updateFinalForRG)
    "cloth_spheres": lambda: scenes.cloth_spheres(n_layers=2, n=17, n_side=2, level=1, seed=31),
}


@pytest.fixture(scope="module", autouse=True)
def _cr_libm():
    port.set_libm(port.LIBM_CR)
    yield
    port.set_libm(port.LIBM_NATIVE)


def make_pair(sc):
    gpu = CollisionSolver3d()
    CollisionSolver3d.set_params_from(sc.params)
    gpu.assembleFromInterface(sc, sc.dt)
    gpu.set_debug(True, True)
    orc = port.OracleSolver(sc)
    return gpu, orc


@pytest.mark.parametrize("name", list(SCENES))
def test_phase_parity(name):
    sc = SCENES[name]()
    gpu, orc = make_pair(sc)
    x, vel = sc.x.copy(), sc.vel.copy()
    total_contacts = 0
    for step in range(3):
        x, vel, stats = run_step_by_phases(gpu, orc, sc, x, vel)
        total_contacts += sum(s["contacts"] for s in stats)
    if name not in ("box_boundary", "sheet_wall"):
        assert total_contacts > 0, "scene never collided: the test would be vacuous"
    gpu.close()


@pytest.mark.parametrize("name", ["two_sheets", "mixed", "ball_plane", "layered_4x24", "sheet_wall", "cloth_spheres"])
def test_whole_step_parity(name):
    """clsn_step_host (the drop-in call) against orc_resolve over several steps."""
    sc = SCENES[name]()
    gpu, orc = make_pair(sc)
    gpu.set_debug(False, False)
    x, vel = sc.x.copy(), sc.vel.copy()
    # updateFinalForRG (dcollid.cpp:626-675): the caller's centre-of-mass data of every hyper-surface, moved between
    # the steps the way the application's propagation would
    nhs = len(sc.hs_mass)
    com_g = np.stack([sc.x[sc.vhs == b].mean(axis=0) if (sc.vhs == b).any() else np.zeros(3) for b in range(nhs)])
    velo_g = np.zeros((nhs, 3))
    com_o, velo_o = com_g.copy(), velo_g.copy()
    for step in range(4):
        gpu.set_exact_stats(step % 2 == 0)   # pruned traversal on odd steps: same results, fewer candidates counted
        xn = x + sc.dt * vel
        orc.set_state(x, xn)
        vo = vel.copy()
        st_o = orc.resolve(vo)
        orc.update_final_for_rg(com_o, velo_o)
        xg = xn.copy()
        vg = vel.copy()
        has = gpu.resolveCollision(x, xg, vg, bodies=(com_g, velo_g))
        assert same_bits(com_g, com_o) and same_bits(velo_g, velo_o), "updateFinalForRG: centre of mass / velocity differ"
        com_g += sc.dt * velo_g
        com_o += sc.dt * velo_o
        st = gpu.last_stats
        assert st["proximity"]["true_pairs"] == st_o[0]
        assert st["n_ccd_passes"] == st_o[1]
        assert [p["true_pairs"] for p in st["ccd"]] == st_o[2:2 + st_o[1]]
        assert st["proximity"]["candidates"] == st_o[8]
        if step % 2 == 0:
            assert [p["candidates"] for p in st["ccd"]] == st_o[9:9 + st_o[1]]
        else:
            assert all(a <= b for a, b in zip([p["candidates"] for p in st["ccd"]], st_o[9:9 + st_o[1]]))
        assert int(st["still_colliding"]) == st_o[7]
        assert st["zone_iterations"] == st_o[14] and st["zones"] == st_o[15]   # the impact-zone fail-safe, when entered
        assert (st["strain_sweeps"], st["strain_edges"]) == (st_o[16], st_o[17])   # reduceSuperelast
        assert same_bits(xg, orc.get(port.F_X))
        assert same_bits(vg, vo)
        assert np.array_equal(has, orc.geti(port.I_HAS_COLLSN))
        x, vel = xg, vg
    gpu.close()


@pytest.mark.parametrize("name", ["two_sheets", "mixed", "ball_plane", "layered_4x24", "drape_small", "string_string"])
def test_fast_path_and_staged_pipelines_agree(name):
    """The fast-path pipeline (k_fast + k_exact: plain-FP64 fast path, correctly rounded solve of the undecided
    features only) against the staged pipeline that solves every feature correctly rounded: identical contact
    sets, counters and state bits, and the fast path must actually save solves."""
    sc = SCENES[name]()
    outs = []
    # CLSN_TEST_PIPELINE2=1 adds the experimental segment-emission pipeline to the comparison
    pipelines = (0, 1, 2) if os.environ.get("CLSN_TEST_PIPELINE2") == "1" else (0, 1)
    for pipeline in pipelines:
        gpu = CollisionSolver3d()
        CollisionSolver3d.set_params_from(sc.params)
        gpu.assembleFromInterface(sc, sc.dt)
        gpu.set_pipeline(pipeline)
        gpu.set_exact_stats(True)
        x, vel = sc.x.copy(), sc.vel.copy()
        log = []
        for step in range(3):
            xg = x + sc.dt * vel
            vg = vel.copy()
            has = gpu.resolveCollision(x, xg, vg)
            st = gpu.last_stats
            log.append((xg.copy(), vg.copy(), has.copy(),
                        [(p["candidates"], p["true_pairs"], p["contacts"], p["contributions"], p["features"], p["coplanar"])
                         for p in st["ccd"]], [p["exact_solves"] for p in st["ccd"]]))
            x, vel = xg, vg
        outs.append(log)
        gpu.close()
    solves = [0] * len(pipelines)
    for per_step in zip(*outs):
        a = per_step[0]
        for b in per_step[1:]:
            assert same_bits(a[0], b[0]) and same_bits(a[1], b[1]) and np.array_equal(a[2], b[2])
            assert a[3] == b[3]
        for i, o in enumerate(per_step):
            solves[i] += sum(o[4])
    if solves[0] > 1000:
        assert all(s < solves[0] for s in solves[1:]), solves


def test_determinism_and_rerun():
    """same input twice -> identical bits (no floating-point atomics anywhere)"""
    sc = scenes.layered_cloth(4, 24, seed=99)
    gpu, _ = make_pair(sc)
    gpu.set_debug(False, False)
    outs = []
    for rep in range(3):
        x = sc.x.copy()
        vel = sc.vel.copy()
        xg = x + sc.dt * vel
        gpu.resolveCollision(x, xg, vel)
        outs.append((xg.copy(), vel.copy()))
    for o in outs[1:]:
        assert same_bits(o[0], outs[0][0]) and same_bits(o[1], outs[0][1])
    gpu.close()


def test_empty_and_tiny_inputs():
    """one element (no internal node), two far-apart elements, and a degenerate (zero-area) triangle"""
    x = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [5, 5, 5], [6, 5, 5], [5, 6, 5], [2, 2, 2], [2, 2, 2], [2, 2, 2]], float)
    for tris in ([[0, 1, 2]], [[0, 1, 2], [3, 4, 5]], [[0, 1, 2], [3, 4, 5], [6, 7, 8]]):
        tri = np.array(tris, dtype=np.int32)
        sc = scenes.Scene(name="tiny", x=x.copy(), vel=np.zeros_like(x), tri_idx=tri,
                          tri_surf=np.arange(len(tri), dtype=np.int32), bond_idx=np.zeros((0, 2), np.int32),
                          bond_curve=np.zeros(0, np.int32), hs_kind=np.zeros(len(tri), np.int32),
                          hs_mass=np.ones(len(tri)), vflags=np.zeros(len(x), np.uint8),
                          vhs=np.minimum(np.arange(len(x)) // 3, len(tri) - 1).astype(np.int32), dt=1e-3)
        gpu, orc = make_pair(sc)
        run_step_by_phases(gpu, orc, sc, sc.x.copy(), sc.vel.copy())
        gpu.close()


def test_sliced_equals_whole():
    """Multi-GPU decomposition on one device: two contexts each traverse half of the query leaves, their
    record buffers are concatenated (what the NCCL all-gather does) and reduced -> bit-identical to the
    unsliced run, pass after pass."""
    import ctypes as C
    import torch
    from collision_b200.dist import BODY_RECORD_BYTES, POINT_RECORD_BYTES, _DevPtr

    sc = scenes.mixed()
    whole, _ = make_pair(sc)
    parts = [make_pair(sc)[0] for _ in range(2)]
    for g in [whole] + parts:
        g.set_debug(False, False)
    for r, g in enumerate(parts):
        g.ctx.check(g.ctx.L.clsn_set_slice(g.ctx.h, r, 2))
    dev = torch.device("cuda", 0)
    x, vel = sc.x.copy(), sc.vel.copy()
    n_rec = 0
    for step in range(3):
        xn = x + sc.dt * vel
        for g in [whole] + parts:
            g.upload(x, xn)
            g.avg_velocity()
        for ps in range(4):
            mode = 0 if ps == 0 else 1
            sw = whole.detect(mode)
            bufs_p, bufs_b, n_true = [], [], 0
            for g in parts:
                g.detect(mode)
                pp, pb = C.c_void_p(), C.c_void_p()
                npr, nbr, nt = C.c_int64(), C.c_int64(), C.c_int64()
                g.ctx.check(g.ctx.L.clsn_export_records(g.ctx.h, C.byref(pp), C.byref(npr), C.byref(pb), C.byref(nbr), C.byref(nt)))
                nb_p, nb_b = npr.value * POINT_RECORD_BYTES, nbr.value * BODY_RECORD_BYTES
                bufs_p.append(torch.as_tensor(_DevPtr(pp.value, nb_p), device=dev)[:nb_p].clone())
                bufs_b.append(torch.as_tensor(_DevPtr(pb.value, nb_b), device=dev)[:nb_b].clone())
                n_true += nt.value
            allp, allb = torch.cat(bufs_p), torch.cat(bufs_b)
            torch.cuda.synchronize()
            assert n_true == sw["true_pairs"]
            assert allp.numel() // POINT_RECORD_BYTES + allb.numel() // BODY_RECORD_BYTES == sw["contributions"]
            n_rec += sw["contributions"]
            for g in parts:
                g.ctx.check(g.ctx.L.clsn_import_records(g.ctx.h, allp.data_ptr() if allp.numel() else None,
                                                        allp.numel() // POINT_RECORD_BYTES,
                                                        allb.data_ptr() if allb.numel() else None,
                                                        allb.numel() // BODY_RECORD_BYTES))
                g.apply(True)
            whole.apply(True)
            _, avw, hasw = whole.download()
            for g in parts:
                _, av, has = g.download()
                assert same_bits(av, avw) and np.array_equal(has, hasw)
        for g in [whole] + parts:
            g.boundary()
            g.final_position()
        xw, avw, hasw = whole.download()
        for g in parts:
            xg, _, _ = g.download()
            assert same_bits(xg, xw)
        vel[hasw != 0] = avw[hasw != 0]
        x = xw
    assert n_rec > 0
    for g in [whole] + parts:
        g.close()


def test_config5_family_small():
    """cloth stack + fast movable rigid spheres (config 5 generator, small): rigid-fabric and rigid-rigid
    CCD with displacement per step >> edge length, re-rigidification every pass"""
    sc = scenes.cloth_spheres(n_layers=2, n=17, n_side=2, level=1, seed=31)
    gpu, orc = make_pair(sc)
    x, vel = sc.x.copy(), sc.vel.copy()
    total = 0
    for step in range(3):
        x, vel, stats = run_step_by_phases(gpu, orc, sc, x, vel)
        total += sum(s["contacts"] for s in stats)
    assert total > 0
    gpu.close()


def test_nan_input_is_reported_not_propagated():
    """the reference prints and clean_up(ERROR)s on NaN/Inf average velocities (dcollid.cpp:178-184);
    the library returns CLSN_E_NUMERIC and the host mirror raises"""
    from collision_b200.solver import CollisionError
    sc = scenes.two_sheets(n=8)
    gpu, _ = make_pair(sc)
    x = sc.x.copy()
    xn = x + sc.dt * sc.vel
    xn[3, 1] = np.nan
    vel = sc.vel.copy()
    with pytest.raises(CollisionError):
        gpu.resolveCollision(x, xn, vel)
    # the context stays usable
    xn = x + sc.dt * sc.vel
    gpu.resolveCollision(x, xn, vel)
    assert np.isfinite(xn).all()
    gpu.close()


def test_zero_time_step():
    """dt <= ROUND_EPS: avgVel = 0 (dcollid.cpp:174-177), nothing moves"""
    sc = scenes.two_sheets(n=8)
    sc.dt = 0.0
    gpu, orc = make_pair(sc)
    x, vel = sc.x.copy(), sc.vel.copy()
    xn = x + 1e-3 * vel
    orc.set_dt(0.0)
    orc.set_state(x, xn)
    vo = vel.copy()
    orc.resolve(vo)
    xg, vg = xn.copy(), vel.copy()
    gpu.resolveCollision(x, xg, vg)
    assert same_bits(xg, orc.get(port.F_X)) and same_bits(vg, vo)
    gpu.close()


ZONE_SCENES = {
    "layered_fast": lambda: scenes.layered_cloth(4, 13, speed=3.0),
    "sheets_fast": lambda: scenes.two_sheets(n=10, speed=10.0),
    "mixed": lambda: scenes.mixed(),
}


@pytest.mark.parametrize("name", list(ZONE_SCENES))
def test_impact_zone_failsafe_matches_oracle(name):
    """Scenes that still collide after the 5 CCD passes: the drop-in call enters computeImpactZone
    (dcollid.cpp:227-265) like the reference; zone iterations, zone counts and the final state equal
    the oracle's bit for bit, and the fail-safe really ran."""
    sc = ZONE_SCENES[name]()
    gpu, orc = make_pair(sc)
    gpu.set_debug(False, False)
    x, vel = sc.x.copy(), sc.vel.copy()
    iters = 0
    for step in range(5 if name == "mixed" else 2):
        xn = x + sc.dt * vel
        orc.set_state(x, xn)
        vo = vel.copy()
        st_o = orc.resolve(vo)
        xg, vg = xn.copy(), vel.copy()
        has = gpu.resolveCollision(x, xg, vg)
        st = gpu.last_stats
        assert [p["true_pairs"] for p in st["ccd"]] == st_o[2:2 + st_o[1]]
        assert int(st["still_colliding"]) == st_o[7]
        assert (st["zone_iterations"], st["zones"]) == (st_o[14], st_o[15]), f"step {step}"
        assert same_bits(xg, orc.get(port.F_X)) and same_bits(vg, vo), f"step {step}"
        assert np.array_equal(has, orc.geti(port.I_HAS_COLLSN))
        iters += st["zone_iterations"]
        x, vel = xg, vg
    assert iters >= 2, "the fail-safe never ran: the test would be vacuous"
    gpu.close()


def test_impact_zone_loop_alone_and_disabled():
    """clsn_compute_impact_zone on a caller-driven phase sequence == the oracle's loop (avgVel bit for
    bit, the state ends collision free); with the fail-safe off the step stops after the CCD passes."""
    sc = scenes.layered_cloth(4, 13, speed=3.0)
    gpu, orc = make_pair(sc)
    gpu.set_debug(False, False)
    x, vel = sc.x.copy(), sc.vel.copy()
    xn = x + sc.dt * vel
    orc.set_state(x, xn)
    orc.avg_velocity()
    gpu.upload(x, xn)
    gpu.avg_velocity()
    for ps in range(6):
        mode = port.PROXIMITY if ps == 0 else port.COLLISION
        n_o = orc.detect(mode)
        assert gpu.detect(mode)["true_pairs"] == n_o
        orc.apply(True)
        gpu.apply(True)
    assert n_o > 0
    z_o = orc.impact_zone()
    z = gpu.computeImpactZone()
    assert (z["iterations"], z["zones"], z["true_pairs"]) == tuple(z_o) and z["converged"] == 1 and z["merges"] > 0
    assert same_bits(gpu.download()[1], orc.get(port.F_AVGVEL))
    assert gpu.detect(port.COLLISION)["true_pairs"] == 0
    gpu.close()
    # fail-safe off: identical to an oracle that stops after MAX_ITER passes, and still colliding
    off = CollisionSolver3d(impact_zones=False, strain_limiting=False)
    off.assembleFromInterface(sc, sc.dt)
    o2 = port.OracleSolver(sc, impact_zones=False, strain_limiting=False)
    o2.set_state(x, xn)
    vo, xg, vg = vel.copy(), xn.copy(), vel.copy()
    st_o = o2.resolve(vo)
    off.resolveCollision(x, xg, vg)
    assert off.last_stats["still_colliding"] and off.last_stats["zone_iterations"] == 0 and st_o[14] == 0
    assert same_bits(xg, o2.get(port.F_X)) and same_bits(vg, vo)
    off.close()


@pytest.mark.parametrize("name", ["two_sheets", "mixed", "layered", "string_string"])
def test_strain_limiting_matches_oracle(name):
    """reduceSuperelast (dcollid.cpp:485-596) alone on kicked velocity fields: the wavefront-scheduled
    kernel reproduces the sequential Gauss-Seidel sweeps bit for bit -- same sweep count (1..10), same
    number of edges averaged in the last sweep, same avgVel."""
    from parity_util import STRAIN_SCENES, strain_inputs
    sc = STRAIN_SCENES[name]()
    gpu, orc = make_pair(sc)
    gpu.set_debug(False, False)
    xn = sc.x + sc.dt * sc.vel
    orc.set_state(sc.x, xn)
    gpu.upload(sc.x, xn)
    sweeps = []
    for case in range(3):
        av = strain_inputs(sc, case)
        orc.set_avgvel(av)
        gpu.set_avgvel(av)
        r_o = orc.strain_limit()
        r_g = gpu.reduceSuperelast()
        assert r_g == r_o, (case, r_g, r_o)
        assert same_bits(gpu.download()[1], orc.get(port.F_AVGVEL)), case
        sweeps.append(r_o[0])
    assert sweeps[0] == 1 and max(sweeps) > 1
    gpu.close()
