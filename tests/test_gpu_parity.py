"""GPU parity tests proper: the CUDA path, called through the C ABI, against the C oracle
(correctly rounded libm flavour) on the same seeded inputs -- bit for bit.

Integer/index work (candidate sets, contact sets, counts, has_collsn) must be identical; FP64
results (times of impact, normals, weights, per-point impulse sums, avgVel, final positions) are
compared BITWISE, which is stronger than the 1e-12 relative tolerance the north star allows.
"""
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


@pytest.mark.parametrize("name", ["two_sheets", "mixed", "ball_plane", "layered_4x24", "sheet_wall"])
def test_whole_step_parity(name):
    """clsn_step_host (the drop-in call) against orc_resolve over several steps."""
    sc = SCENES[name]()
    gpu, orc = make_pair(sc)
    gpu.set_debug(False, False)
    x, vel = sc.x.copy(), sc.vel.copy()
    for step in range(4):
        xn = x + sc.dt * vel
        orc.set_state(x, xn)
        vo = vel.copy()
        st_o = orc.resolve(vo)
        xg = xn.copy()
        vg = vel.copy()
        has = gpu.resolveCollision(x, xg, vg)
        st = gpu.last_stats
        assert st["proximity"]["true_pairs"] == st_o[0]
        assert st["n_ccd_passes"] == st_o[1]
        assert [p["true_pairs"] for p in st["ccd"]] == st_o[2:2 + st_o[1]]
        assert st["proximity"]["candidates"] == st_o[8]
        assert [p["candidates"] for p in st["ccd"]] == st_o[9:9 + st_o[1]]
        assert int(st["still_colliding"]) == st_o[7]
        assert same_bits(xg, orc.get(port.F_X))
        assert same_bits(vg, vo)
        assert np.array_equal(has, orc.geti(port.I_HAS_COLLSN))
        x, vel = xg, vg
    gpu.close()


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
