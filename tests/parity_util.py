"""Shared helpers of the parity tests: drive the CUDA solver and the C oracle through the same
phases from identical inputs and compare bit for bit."""
import numpy as np

from collision_b200 import scenes
from oracle import port


def sort_pairs(p):
    p = np.asarray(p).reshape(-1, 2)
    if len(p) == 0:
        return p.astype(np.int32)
    q = np.sort(p, axis=1)
    return q[np.lexsort((q[:, 1], q[:, 0]))].astype(np.int32)


def sort_contacts(c):
    if len(c) == 0:
        return c
    return c[np.lexsort((c["feature"], c["eb"], c["ea"]))]


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def same_bits(a, b):
    """bitwise equality of float arrays, except that +0 == -0"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and bool(np.all((bits(a) == bits(b)) | ((a == 0) & (b == 0))))


def assert_contacts_equal(cg, co):
    cg, co = sort_contacts(cg), sort_contacts(co)
    assert len(cg) == len(co), f"contact count {len(cg)} vs {len(co)}"
    for f in ("ea", "eb", "feature", "kind", "p"):
        assert np.array_equal(cg[f], co[f]), f"contact field {f} differs"
    for f in ("root", "dist", "nor", "w"):
        assert same_bits(cg[f], co[f]), f"contact field {f} differs (max abs {np.abs(cg[f] - co[f]).max():.3e})"


def compare_pass(gpu, orc, mode, scene, check_contacts=True):
    """one detection pass on both sides from their current (identical) state"""
    sg = gpu.detect(mode)
    n_true = orc.detect(mode)
    cand_o = orc.candidates()
    cand_g = sort_pairs(gpu.candidates())
    assert np.array_equal(cand_g, cand_o), f"candidate sets differ: {len(cand_g)} vs {len(cand_o)}"
    assert sg["candidates"] == len(cand_o)
    assert sg["true_pairs"] == n_true, f"true pairs {sg['true_pairs']} vs {n_true}"
    co = orc.contacts()
    assert sg["contacts"] == len(co)
    if check_contacts:
        assert_contacts_equal(gpu.contacts(), co)
    imp, fric, cnt, irg, crg = gpu.accumulators()
    assert np.array_equal(cnt, orc.geti(port.I_CNT)), "per-point contribution counts differ"
    assert same_bits(imp, orc.get(port.F_IMP)), "collsnImpulse sums differ"
    assert same_bits(fric, orc.get(port.F_FRIC)), "friction sums differ"
    oirg, ocrg = orc.get_body()
    assert np.array_equal(crg, ocrg), "collsn_num_RG differs"
    assert same_bits(irg, oirg), "collsnImpulse_RG differs"
    return sg, n_true


def run_step_by_phases(gpu, orc, scene, x, vel, max_passes=5):
    """resolveCollision phase by phase on both sides, asserting parity after every phase.
    Returns (x_final, vel_final, stats)."""
    xn = x + scene.dt * vel
    gpu.upload(x, xn)
    orc.set_state(x, xn)
    orc.set_dt(scene.dt)
    gpu.avg_velocity()
    orc.avg_velocity()
    _, av, _ = gpu.download()
    assert same_bits(av, orc.get(port.F_AVGVEL))
    stats = []
    sg, _ = compare_pass(gpu, orc, port.PROXIMITY, scene)
    stats.append(sg)
    gpu.apply(True)
    orc.apply(True)
    _, av, has = gpu.download()
    assert same_bits(av, orc.get(port.F_AVGVEL)), "avgVel after proximity apply differs"
    for it in range(max_passes):
        sg, n_true = compare_pass(gpu, orc, port.COLLISION, scene)
        stats.append(sg)
        gpu.apply(True)
        orc.apply(True)
        _, av, has = gpu.download()
        assert same_bits(av, orc.get(port.F_AVGVEL)), f"avgVel after CCD pass {it} differs"
        assert np.array_equal(has, orc.geti(port.I_HAS_COLLSN))
        if n_true == 0:
            break
    gpu.boundary()
    orc.boundary()
    gpu.final_position()
    orc.final_position()
    xg, av, has = gpu.download()
    assert same_bits(av, orc.get(port.F_AVGVEL)), "avgVel after boundary differs"
    assert same_bits(xg, orc.get(port.F_X)), "final positions differ"
    assert np.array_equal(has, orc.geti(port.I_HAS_COLLSN))
    vo = vel.copy()
    orc.final_velocity(vo)
    vg = vel.copy()
    vg[has != 0] = av[has != 0]
    assert same_bits(vg, vo)
    return xg, vg, stats


# ---- strain limiting (reduceSuperelast): shared by the golden generator, the oracle test and the GPU test
STRAIN_SCENES = {
    "two_sheets": lambda: scenes.two_sheets(n=10),
    "mixed": lambda: scenes.mixed(),
    "layered": lambda: scenes.layered_cloth(4, 13, speed=3.0),
    "string_string": lambda: scenes.string_string(dt=0.01, gap=0.003),
}


def strain_inputs(sc, case):
    """avgVel fed to reduceSuperelast: the step's own average velocity, kicked on 5 % of the points
    (case 0: no kick, nothing over the limit)."""
    rng = np.random.default_rng(100 + case)
    av = np.repeat(sc.vel[None], 1, 0)[0].copy()
    if case:
        av = av + rng.normal(0, 30.0 * case, (sc.V, 3)) * (rng.random((sc.V, 1)) < 0.05)
    return av
