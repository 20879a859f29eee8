"""CPU tests: the C restatement (oracle/collision_oracle.c) against golden vectors produced by the
reference's own sources (tests/golden/make_golden.py, oracle/_ref).  Bit for bit, native libm."""
import os

import numpy as np
import pytest

from collision_b200 import scenes
from oracle import port
from parity_util import same_bits, sort_pairs

HERE = os.path.dirname(os.path.abspath(__file__))
G = os.path.join(HERE, "golden")


@pytest.fixture(autouse=True)
def _native_libm():
    port.set_libm(port.LIBM_NATIVE)
    yield
    port.set_libm(port.LIBM_NATIVE)


def test_feature_known_answers():
    d = np.load(os.path.join(G, "features.npz"))
    n = len(d["kind"])
    assert n >= 500
    hits = 0
    for i in range(n):
        o = port.feature(int(d["kind"][i]), d["x_old"][i], d["coords"][i], d["avgvel"][i], d["flags"][i], d["mass"][i],
                         float(d["h"][i]), float(d["dt"][i]), d["params"])
        assert o["ret"] == int(d["ret"][i]), f"case {i} kind {d['kind'][i]}: ret {o['ret']} vs {d['ret'][i]}"
        if d["kind"][i] in (0, 3, 4):
            assert same_bits(o["roots"], d["roots"][i]), f"case {i}: roots"
        if d["kind"][i] in (3, 4):
            assert same_bits(o["hit_root"], d["hit_root"][i]), f"case {i}: time of impact"
        assert same_bits(o["acc"], d["acc"][i]), f"case {i}: accumulators"
        hits += int(d["ret"][i] != 0)
    assert hits > 100


SCENE_MAKERS = {
    "string_string": lambda: scenes.string_string(dt=0.01, gap=0.003),
    "two_sheets": lambda: scenes.two_sheets(n=10),
    "mixed": lambda: scenes.mixed(),
    "ball_plane": lambda: scenes.ball_plane(level=2, gap=2e-4),
    "sheet_wall": lambda: scenes.sheet_wall(n=10),
}


@pytest.mark.parametrize("name", list(SCENE_MAKERS))
def test_scene_replay_bit_exact(name):
    """Replay the reference's own callback sequence through the restatement: accumulators, avgVel after
    every pass and the final state must equal the reference's bit for bit; the restatement's own broad
    phase must find exactly the reference's candidate set."""
    sc = SCENE_MAKERS[name]()
    d = np.load(os.path.join(G, f"scene_{name}.npz"))
    o = port.OracleSolver(sc)
    vel = sc.vel.copy()
    total_true = 0
    for step in range(int(d["n_steps"])):
        x = d[f"s{step}_x_old"]
        assert same_bits(x, sc.x) if step == 0 else True
        vel = sc.vel.copy() if step == 0 else d[f"s{step - 1}_vel"].copy()  # the reference's own end-of-step vel
        o.set_state(x, x + sc.dt * vel)
        o.avg_velocity()
        assert same_bits(o.get(port.F_AVGVEL), d[f"s{step}_avgvel0"])
        for ps in range(int(d[f"s{step}_npass"])):
            k = f"s{step}_p{ps}_"
            mode = port.PROXIMITY if ps == 0 else port.COLLISION
            pairs = d[k + "pairs"]
            # (1) candidate SET of the restatement's own broad phase == the reference tree's callbacks
            body = o.get_body()
            av = o.get(port.F_AVGVEL)
            o.detect(mode)
            assert np.array_equal(o.candidates(), sort_pairs(pairs[:, :2])), f"{k}: candidate sets differ"
            # (2) narrow phase + accumulation, replaying the reference's order
            o.set_state(x, x)          # clears the accumulators of the canonical-order run above
            o.set_avgvel(av)
            o.set_body(*body)
            n_true = o.detect_ordered(mode, pairs)
            assert n_true == int(d[k + "count"]) == int(pairs[:, 2].sum())
            assert np.array_equal(sort_pairs(o.true_pairs()), sort_pairs(pairs[pairs[:, 2] == 1][:, :2]))
            assert np.array_equal(o.geti(port.I_CNT), d[k + "cnt"])
            assert same_bits(o.get(port.F_IMP), d[k + "imp"]), f"{k}: collsnImpulse"
            assert same_bits(o.get(port.F_FRIC), d[k + "fric"]), f"{k}: friction"
            irg, crg = o.get_body()
            assert same_bits(irg[sc.vhs], d[k + "imp_rg"]), f"{k}: collsnImpulse_RG"
            assert np.array_equal(crg[sc.vhs], d[k + "cnt_rg"])
            o.apply(True)
            assert same_bits(o.get(port.F_AVGVEL), d[k + "avgvel"]), f"{k}: avgVel after updateAverageVelocity"
            total_true += n_true
        o.boundary()
        o.final_position()
        # has_collsn was cleared by the set_state() calls used for the replay, so updateFinalVelocity is
        # checked through the avgVel it copies: vel == avgVel wherever the reference flagged a collision
        assert same_bits(o.get(port.F_X), d[f"s{step}_x"]), "final positions"
        has = d[f"s{step}_has"] != 0
        assert same_bits(o.get(port.F_AVGVEL)[has], d[f"s{step}_vel"][has]), "final velocities"
    if name not in ("sheet_wall",):
        assert total_true > 0


ZONE_SCENES = {
    "mixed_zone": lambda: scenes.mixed(),
    "layered_zone": lambda: scenes.layered_cloth(4, 13, speed=3.0),
    "sheets_zone": lambda: scenes.two_sheets(n=10, speed=10.0),
}


@pytest.mark.parametrize("name", list(ZONE_SCENES))
def test_impact_zone_replay_bit_exact(name):
    """Scenes that still collide after the 5 CCD passes: the restatement of computeImpactZone
    (createImpZone merges inside the narrow phase, updateImpactZoneVelocity) replays the reference's
    callback sequence and must reproduce avgVel after every iteration, the zone counts and the final
    state bit for bit."""
    sc = ZONE_SCENES[name]()
    d = np.load(os.path.join(G, f"scene_{name}.npz"))
    o = port.OracleSolver(sc)
    vel = sc.vel.copy()
    zone_iters = 0
    for step in range(int(d["n_steps"])):
        x = d[f"s{step}_x_old"]
        o.set_state(x, x + sc.dt * vel)
        o.avg_velocity()
        assert same_bits(o.get(port.F_AVGVEL), d[f"s{step}_avgvel0"])
        for ps in range(int(d[f"s{step}_npass"])):
            k = f"s{step}_p{ps}_"
            n_true = o.detect_ordered(port.PROXIMITY if ps == 0 else port.COLLISION, d[k + "pairs"])
            assert n_true == int(d[k + "count"])
            assert same_bits(o.get(port.F_IMP), d[k + "imp"]) and same_bits(o.get(port.F_FRIC), d[k + "fric"])
            assert np.array_equal(o.geti(port.I_CNT), d[k + "cnt"])
            o.apply(True)
            assert same_bits(o.get(port.F_AVGVEL), d[k + "avgvel"]), f"{k}: avgVel"
        nz = int(d[f"s{step}_nzone"])
        if nz:
            o.set_imp_zone(True)
            for it in range(nz):
                k = f"s{step}_z{it}_"
                assert o.detect_ordered(port.COLLISION, d[k + "pairs"]) == int(d[k + "count"])
                o.apply(True)
                assert same_bits(o.get(port.F_AVGVEL), d[k + "avgvel"]), f"{k}: avgVel before the zones"
                assert o.zone_velocity() == int(d[k + "zones"])
                assert same_bits(o.get(port.F_AVGVEL), d[k + "zvel"]), f"{k}: avgVel after updateImpactZoneVelocity"
            assert int(d[f"s{step}_z{nz - 1}_count"]) == 0      # the loop ends on a collision-free pass
            o.set_imp_zone(False)
            zone_iters += nz
        o.boundary()
        o.final_position()
        o.final_velocity(vel)
        assert same_bits(o.get(port.F_X), d[f"s{step}_x"])
        assert np.array_equal(o.geti(port.I_HAS_COLLSN), d[f"s{step}_has"])
        assert same_bits(vel, d[f"s{step}_vel"])
    assert zone_iters >= 2


def test_impact_zone_loop_in_resolve():
    """orc_resolve with the fail-safe enabled == the phase-by-phase loop, and it ends collision free."""
    sc = scenes.layered_cloth(4, 13, speed=3.0)
    o1, o2 = port.OracleSolver(sc), port.OracleSolver(sc)
    o1.enable_impact_zones(True)
    v1, v2 = sc.vel.copy(), sc.vel.copy()
    o1.set_state(sc.x, sc.x + sc.dt * sc.vel)
    st = o1.resolve(v1)
    assert st[7] == 1 and st[14] >= 2 and st[15] >= 1
    o2.set_state(sc.x, sc.x + sc.dt * sc.vel)
    o2.avg_velocity()
    o2.detect(port.PROXIMITY)
    o2.apply(True)
    for _ in range(5):
        n = o2.detect(port.COLLISION)
        o2.apply(True)
    assert n > 0
    out = o2.impact_zone()
    assert out[0] == st[14] and out[1] == st[15]
    o2.boundary(); o2.final_position(); o2.strain_limit(); o2.final_velocity(v2)
    assert same_bits(o1.get(port.F_X), o2.get(port.F_X)) and same_bits(v1, v2)
    o2.set_state(o2.get(port.F_X_OLD), o2.get(port.F_X))
    o2.avg_velocity()
    assert o2.detect(port.COLLISION) == 0


def test_strain_limiting_matches_reference_golden():
    """reduceSuperelast restated (sequential Gauss-Seidel sweeps over the edges in hseList order) against
    the compiled reference's output on kicked velocity fields, bit for bit; sweeps run 1..10 times."""
    from parity_util import STRAIN_SCENES, strain_inputs
    d = np.load(os.path.join(G, "strain.npz"))
    sweeps = set()
    for name, mk in STRAIN_SCENES.items():
        sc = mk()
        o = port.OracleSolver(sc)
        o.set_state(sc.x, sc.x + sc.dt * sc.vel)
        for case in range(3):
            av = strain_inputs(sc, case)
            o.set_avgvel(av)
            it, _ = o.strain_limit()
            sweeps.add(it)
            assert same_bits(o.get(port.F_AVGVEL), d[f"{name}_{case}_out"]), (name, case)
            assert (case == 0) == (it == 1)
    assert 10 in sweeps and len(sweeps) >= 3


def test_canonical_order_is_order_independent_for_sets():
    """canonical (a<b sorted) evaluation finds the same candidate set and, for point-triangle-only
    contacts (static sphere), the same per-point sums up to summation order (<= 1e-12 relative)."""
    sc = scenes.drape(n=24, level=2)
    o = port.OracleSolver(sc)
    x, vel = sc.x.copy(), sc.vel.copy()
    o.set_state(x, x + sc.dt * vel)
    o.avg_velocity()
    o.detect(port.COLLISION)
    c1 = o.candidates().copy()
    imp1 = o.get(port.F_IMP).copy()
    rev = c1[::-1][:, ::-1].copy()   # reversed order, swapped roles
    o.set_state(x, x)
    o.detect_ordered(port.COLLISION, rev)
    imp2 = o.get(port.F_IMP)
    scale = max(np.abs(imp1).max(), 1e-300)
    assert np.abs(imp1 - imp2).max() <= 1e-12 * scale


def test_libm_flavours_agree_almost_everywhere():
    """correctly rounded vs native libm: identical contact sets on a small CCD-heavy scene"""
    sc = scenes.two_sheets(n=10)
    res = []
    for mode in (port.LIBM_NATIVE, port.LIBM_CR):
        port.set_libm(mode)
        o = port.OracleSolver(sc)
        x, vel = sc.x.copy(), sc.vel.copy()
        for step in range(2):
            o.set_state(x, x + sc.dt * vel)
            st = o.resolve(vel)
            x = o.get(port.F_X)
        res.append((st, o.true_pairs().copy()))
    assert res[0][0][1] == res[1][0][1]  # same number of CCD passes


def test_whole_step_matches_reference_when_available():
    """When oracle/_ref is present (this container), whole steps through resolveCollision agree on the
    integer outputs and agree numerically within summation-order noise on a point-triangle scene."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built here")
    sc = scenes.drape(n=24, level=2)
    r = ref.RefSolver(sc)
    r.set_rest_lengths(sc.x)
    o = port.OracleSolver(sc, impact_zones=True, strain_limiting=True)
    x, vel = sc.x.copy(), sc.vel.copy()
    for step in range(3):
        xn = x + sc.dt * vel
        r.set_state(x, xn, vel)
        r.assemble(sc.dt)
        r.resolve(True)      # resolveCollision() verbatim: impact zones and strain limiting included
        o.set_state(x, xn)
        vo = vel.copy()
        o.resolve(vo)
        xr, vr = r.get(ref.F_COORDS), r.get(ref.F_VEL)
        assert np.array_equal(r.geti(ref.I_HAS_COLLSN), o.geti(port.I_HAS_COLLSN))
        # canonical (a<b) order vs the reference's tree order: the pair order changes which role the
        # two edges of an edge-edge test play, hence the last bits of the root (DESIGN.md "order
        # sensitivity"); bit-exactness is what test_scene_replay_bit_exact establishes.  Here: sanity.
        assert np.abs(xr - o.get(port.F_X)).max() <= 1e-8 * np.abs(xr).max()
        assert np.abs(vr - vo).max() <= 1e-5 * max(np.abs(vr).max(), 1.0)
        x, vel = xr, vr


def test_update_final_for_rg_matches_reference():
    """updateFinalForRG (dcollid.cpp:626-675: centre of mass / its velocity of every hit movable body, mrg_com bookkeeping)
    restated, against the compiled reference over several steps of the ball_plane deck with the application moving the
    centres of mass between steps like FronTier's propagation does.  The restatement is fed the reference's own avgVel and
    has_collsn, so the comparison is bit for bit.  Variants: natural flags (the body's first point in hseList order has no
    collision, a later one has), every point flagged (first point collides), nothing flagged."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built here")
    sc = scenes.ball_plane(level=2, gap=2e-4)
    nhs = len(sc.hs_mass)
    movable = (sc.vflags & 2) != 0
    assert movable.any()
    r = ref.RefSolver(sc)
    o = port.OracleSolver(sc, impact_zones=False, strain_limiting=False)
    com = np.zeros((nhs, 3))
    for b in range(nhs):
        com[b] = sc.x[sc.vhs == b].mean(axis=0)
    velo = np.zeros((nhs, 3))
    com_o, velo_o = com.copy(), velo.copy()
    x, vel = sc.x.copy(), sc.vel.copy()
    hit_steps = 0
    for step in range(6):
        xn = x + sc.dt * vel
        r.set_bodies(com, velo)
        r.set_state(x, xn, vel)
        r.assemble(sc.dt)
        r.puti(ref.I_HAS_COLLSN, np.zeros(sc.V, np.int32))
        # resolveCollision phase by phase (dcollid.cpp:317-362), so that the flags can be overridden before the last one
        r.phase(ref.PH_AVG_VELOCITY)
        r.phase(ref.PH_PROXIMITY_DETECT)
        r.phase(ref.PH_APPLY)
        for _ in range(5):
            n = r.phase(ref.PH_COLLISION_DETECT)
            r.phase(ref.PH_APPLY)
            if n == 0:
                break
        r.phase(ref.PH_BOUNDARY)
        r.phase(ref.PH_FINAL_POSITION)
        variant = step % 3
        if variant == 1:      # first point of every body collides
            r.puti(ref.I_HAS_COLLSN, np.ones(sc.V, np.int32))
        has = r.geti(ref.I_HAS_COLLSN) != 0
        av = r.get(ref.F_AVGVEL)
        r.phase(ref.PH_FINAL_VELOCITY)     # updateFinalVelocity + updateFinalForRG
        com_r, velo_r = r.get_bodies(nhs)
        o.set_state(x, xn)
        o.set_dt(sc.dt)
        o.set_avgvel(av)
        o.set_has_collsn(has)
        o.update_final_for_rg(com_o, velo_o)
        assert same_bits(com_o, com_r), f"step {step}: centre of mass"
        assert same_bits(velo_o, velo_r), f"step {step}: centre-of-mass velocity"
        hit_steps += int((has & movable).any())
        # the application's propagation of the bodies between two collision steps
        com = com_r + sc.dt * velo_r
        velo = velo_r
        com_o, velo_o = com.copy(), velo.copy()
        x, vel = r.get(ref.F_COORDS), r.get(ref.F_VEL)
    assert hit_steps >= 2
