"""GPU parity against THE REFERENCE ITSELF: the CUDA path (through the C ABI) and oracle/_ref (the unmodified
/root/reference/{AABB,dcollid,dcollid3d}.cpp, prebuilt; see oracle/Makefile) are fed the same arrays and compared pass
by pass from identical inputs -- no C restatement in between.  Checker and tolerances: tests/ref_compare.py
(candidate sets bit-equal; per-feature contact sets equal to the reference's own primitives in canonical order; times
of impact bit-equal against the correctly-rounded-libm build of the reference and <= 1e-8 / 99 % <= 1e-12 against the
native-libm build; point-triangle impulse sums <= 1e-12; edge-edge sums within the stated bound, observed max printed).
Scenes: BASELINE configs 1-4 at sizes the reference finishes in seconds (config 4 at full size:
tests/test_gpu_fullsize.py against the committed fixture of the reference's 1 M-triangle run)."""
import numpy as np
import pytest

from collision_b200 import scenes
from collision_b200.solver import COLLISION, PROXIMITY, CollisionSolver3d
from oracle import ref
import ref_compare

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libcollision_ref.so did not travel")]

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
EE_SUM_BOUND = 2.0   # see tests/test_oracle_vs_reference.py


class GpuImpl:
    def __init__(self, sc):
        self.sc = sc
        self.g = CollisionSolver3d(impact_zones=False, strain_limiting=False)
        CollisionSolver3d.set_params_from(sc.params)
        self.g.assembleFromInterface(sc, sc.dt)
        self.g.set_debug(True, True)

    def upload(self, x_old, x_new):
        self.g.upload(x_old, x_new)

    def avg_velocity(self):
        self.g.avg_velocity()

    def avgvel(self):
        return self.g.download()[1]

    def set_avgvel(self, av):
        self.g.set_avgvel(av)

    def set_body(self, imp, cnt):
        self.g.set_body_accumulators(imp, cnt)

    def detect(self, moving):
        self.g.detect(COLLISION if moving else PROXIMITY)
        imp, fric, cnt, _, _ = self.g.accumulators()
        return dict(candidates=ref_compare.sort_pairs(self.g.candidates()), contacts=self.g.contacts(), cnt=cnt, imp=imp, fric=fric)

    def apply(self):
        self.g.apply(True)

    def has_collsn(self):
        return self.g.download()[2]


@pytest.mark.parametrize("flavour", ["cr", "native"])
@pytest.mark.parametrize("name", list(SCENES))
def test_cuda_path_matches_the_reference(name, flavour):
    if not ref.available(flavour):
        pytest.skip("this build of the reference did not travel")
    sc = SCENES[name]()
    impl = GpuImpl(sc)
    rep = ref_compare.run_steps(sc, impl, n_steps=2, cr_libm=flavour == "cr")
    impl.g.close()
    print(name, flavour, {k: (f"{v:.3e}" if isinstance(v, float) else v) for k, v in rep.items()})
    assert rep.get("ee_sum_rel_max", 0.0) <= EE_SUM_BOUND
    if rep.get("toi"):
        assert rep["toi_le_1e-12"] >= 0.99 * rep["toi"]
        assert flavour != "cr" or rep["toi_bit_equal"] == rep["toi"]
    if name not in ("box_boundary", "sheet_wall"):
        assert rep["contacts"] > 0
