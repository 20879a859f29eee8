"""The C++ host mirror (collision_b200/host/collid_b200.h: CollisionSolver3d / CD_HSE adapters over the
C ABI) driven like the reference's test.cpp, against the Python mirror on the same scene: bit-identical."""
import os
import struct
import subprocess

import numpy as np
import pytest

from collision_b200 import scenes
from collision_b200.solver import CollisionSolver3d
from parity_util import same_bits

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "collision_b200", "host")


def write_scene(sc, path):
    p = sc.params
    with open(path, "wb") as f:
        f.write(struct.pack("6i", sc.V, sc.T, sc.B, sc.n_surf, sc.n_curve, len(sc.hs_kind)))
        f.write(np.array([p.eps, p.thickness, p.k, p.m, p.friction, p.cr, sc.dt, *sc.lo, *sc.hi], dtype=np.float64).tobytes())
        for a, dt in ((sc.x, np.float64), (sc.vel, np.float64), (sc.tri_idx, np.int32), (sc.tri_surf, np.int32),
                      (sc.bond_idx, np.int32), (sc.bond_curve, np.int32), (sc.hs_kind, np.int32), (sc.hs_mass, np.float64),
                      (sc.vflags, np.uint8), (sc.vhs, np.int32)):
            f.write(np.ascontiguousarray(a, dtype=dt).tobytes())


@pytest.mark.parametrize("name", ["two_sheets", "mixed"])
def test_cpp_host_matches_python_host(name, tmp_path):
    exe = os.path.join(HOST, "host_check")
    subprocess.check_call(["make", "-s", "-C", HOST])   # always: a binary built against an older header would be stale
    sc = scenes.two_sheets(n=16) if name == "two_sheets" else scenes.mixed()
    steps = 3
    inp, out = str(tmp_path / "scene.bin"), str(tmp_path / "out.bin")
    write_scene(sc, inp)
    subprocess.check_call([exe, inp, out, str(steps)])
    res = np.fromfile(out, dtype=np.float64).reshape(2, sc.V, 3)
    gpu = CollisionSolver3d()
    CollisionSolver3d.set_params_from(sc.params)
    gpu.assembleFromInterface(sc, sc.dt)
    x, vel = sc.x.copy(), sc.vel.copy()
    for _ in range(steps):
        xg = x + sc.dt * vel
        gpu.resolveCollision(x, xg, vel)
        x = xg
    assert same_bits(res[0], x) and same_bits(res[1], vel)
    gpu.close()


def test_cpp_single_pair_entry_points(tmp_path):
    """isProximity / isCollision of the C++ mirror (collid.h:199-200) on triangle pairs of a stepped two-sheet mesh: the
    first call of a solver creates its small pair context, later calls reuse it -- verdicts and accumulators of the first
    and of the repeated evaluation agree bit for bit (host_check.cpp::check_single_pairs)."""
    exe = os.path.join(HOST, "host_check")
    subprocess.check_call(["make", "-s", "-C", HOST])
    sc = scenes.two_sheets(n=16)
    inp, out = str(tmp_path / "scene.bin"), str(tmp_path / "out.bin")
    write_scene(sc, inp)
    r = subprocess.run([exe, inp, out, "2", "1", "pairs"], capture_output=True, text=True)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 mismatches" in r.stdout
