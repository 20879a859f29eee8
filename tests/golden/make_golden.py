"""Generates the golden fixtures of tests/golden/ by RUNNING THE REFERENCE ITSELF in this container:
oracle/_ref/libcollision_ref.so = /root/reference/{AABB,dcollid,dcollid3d}.cpp compiled unmodified
(oracle/Makefile).  The reference ships no golden vectors of its own, so these are the
pins of the C restatement (tests/test_oracle_golden.py).

    python tests/golden/make_golden.py          # needs /root/reference (build: make -C oracle ref)

Outputs (committed):
    features.npz  known-answer vectors of the file-static primitives of dcollid3d.cpp
                  (isCoplanar, PointToTri, EdgeToEdge, MovingPointToTri, MovingEdgeToEdge)
    scene_<name>.npz  per step and per pass: the reference's ordered callback list (a, b, result),
                  accumulators after the tree query, avgVel after updateAverageVelocity, final state
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from collision_b200 import scenes  # noqa: E402
from oracle import ref  # noqa: E402

sys.path.insert(0, os.path.dirname(HERE))
from parity_util import STRAIN_SCENES, strain_inputs  # noqa: E402

PARAMS = np.array([1e-6, 1e-4, 1000.0, 0.01, 0.02, 0.0])


def feature_cases(rng, n_each=120):
    """Random quads around real contact configurations (plus flag variants and degenerate shapes)."""
    cases = []
    dt = 1e-3
    for kind in range(5):
        for i in range(n_each):
            edge = kind in (2, 4)
            L = 4e-3
            tri = rng.uniform(0, 1, 3)[None, :] + L * rng.uniform(-1, 1, (3, 3))
            if edge:
                p0, p1 = tri[0], tri[1]
                mid = p0 + rng.uniform(-0.2, 1.2) * (p1 - p0)
                d = rng.normal(size=3)
                d -= d.dot(p1 - p0) / (p1 - p0).dot(p1 - p0) * (p1 - p0) * rng.uniform(0.0, 1.0)
                d /= np.linalg.norm(d)
                off = rng.normal(size=3)
                off /= np.linalg.norm(off)
                gap = rng.choice([3e-7, 5e-5, 2e-4, 6e-4])
                c = mid + gap * off
                half = L * rng.uniform(0.3, 1.0)
                x = np.stack([p0, p1, c - half * d, c + half * d])
                approach = -off
            else:
                nrm = np.cross(tri[0] - tri[2], tri[1] - tri[2])
                nrm /= np.linalg.norm(nrm)
                w = rng.dirichlet(np.ones(3)) * rng.uniform(0.8, 1.3) - rng.uniform(0, 0.1)
                if i % 9 == 0:
                    w = np.array([1.0, 0.0, 0.0]) + 1e-12 * rng.normal(size=3)  # corner case |w_i| < 1e-10
                foot = w @ tri + (1 - w.sum()) * tri[2]
                gap = rng.choice([3e-7, 5e-5, 2e-4, 6e-4]) * rng.choice([-1, 1])
                x = np.vstack([tri, (foot + gap * nrm)[None, :]])
                approach = -np.sign(gap) * nrm
            speed = rng.choice([0.05, 0.3, 0.8])
            v = 1e-3 * rng.normal(size=(4, 3))
            if edge:
                v[2:] += speed * approach
                v[:2] -= 0.3 * speed * approach
            else:
                v[3] += speed * approach
                v[:3] -= 0.3 * speed * approach
            if i % 17 == 0:
                v[:] = v[0]          # rigid translation: cubic degenerates to lower order
            if i % 23 == 0:
                x[1] = x[0] + 1e-13  # degenerate triangle / parallel edges
            flags = np.zeros(4, np.uint8)
            r = i % 8
            if r == 1:
                flags[3 if not edge else 2:] = 1                 # static point / static second edge
            elif r == 2:
                flags[:3 if not edge else 2] = 1                 # static triangle / first edge
            elif r == 3:
                flags[:] = 2                                     # movable-movable (rigid-rigid)
            elif r == 4:
                flags[:3 if not edge else 2] = 2                 # movable triangle vs fabric
            elif r == 5:
                flags[3 if not edge else 2:] = 2                 # fabric vs movable point/edge
            elif r == 6:
                flags[0] = 1                                     # one static vertex
            mass = np.array([2.0, 2.0, 2.0, 5.0]) if not edge else np.array([2.0, 2.0, 5.0, 5.0])
            h = 1e-4 if kind in (1, 2) else 1e-6
            coords = x.copy()
            if kind in (1, 2) and i % 2 == 1:
                coords = x + dt * rng.uniform(0, 1) * v      # static test away from x_old (as CCD calls it)
            cases.append((kind, x, coords, v, flags, mass, h, dt))
    return cases


def make_features(path):
    rng = np.random.default_rng(20241017)
    cases = feature_cases(rng)
    n = len(cases)
    out = dict(kind=np.zeros(n, np.int32), x_old=np.zeros((n, 4, 3)), coords=np.zeros((n, 4, 3)),
               avgvel=np.zeros((n, 4, 3)), flags=np.zeros((n, 4), np.uint8), mass=np.zeros((n, 4)), h=np.zeros(n),
               dt=np.zeros(n), ret=np.zeros(n, np.int32), roots=np.zeros((n, 4)), acc=np.zeros((n, 4, 10)),
               hit_root=np.zeros(n))
    for i, (kind, x, coords, v, flags, mass, h, dt) in enumerate(cases):
        r = ref.feature(kind, x, coords, v, flags, mass, h, dt, PARAMS)
        out["kind"][i], out["x_old"][i], out["coords"][i], out["avgvel"][i] = kind, x, coords, v
        out["flags"][i], out["mass"][i], out["h"][i], out["dt"][i] = flags, mass, h, dt
        out["ret"][i], out["roots"][i], out["acc"][i], out["hit_root"][i] = r["ret"], r["roots"], r["acc"], r["hit_root"]
    out["params"] = PARAMS
    np.savez_compressed(path, **out)
    hits = {k: int(out["ret"][out["kind"] == k].sum()) for k in range(5)}
    print(f"features.npz: {n} cases, hits per kind {hits}")


SCENES = {
    "string_string": (lambda: scenes.string_string(dt=0.01, gap=0.003), 2),
    "two_sheets": (lambda: scenes.two_sheets(n=10), 3),
    "mixed": (lambda: scenes.mixed(), 3),
    "ball_plane": (lambda: scenes.ball_plane(level=2, gap=2e-4), 2),
    "sheet_wall": (lambda: scenes.sheet_wall(n=10), 2),
    # scenes that leave collisions after the 5 CCD passes and enter computeImpactZone
    "mixed_zone": (lambda: scenes.mixed(), 4),
    "layered_zone": (lambda: scenes.layered_cloth(4, 13, speed=3.0), 1),
    "sheets_zone": (lambda: scenes.two_sheets(n=10, speed=10.0), 1),
}


def make_scene(name, path):
    mk, nsteps = SCENES[name]
    sc = mk()
    r = ref.RefSolver(sc)
    x, vel = sc.x.copy(), sc.vel.copy()
    out = {"n_steps": np.int32(nsteps)}
    for step in range(nsteps):
        xn = x + sc.dt * vel
        r.set_state(x, xn, vel)
        r.assemble(sc.dt)
        if name.endswith("_zone"):  # recordOriginPosition clears has_collsn every step (dcollid.cpp:100)
            r.puti(ref.I_HAS_COLLSN, np.zeros(sc.V, np.int32))
        r.phase(ref.PH_AVG_VELOCITY)
        out[f"s{step}_x_old"] = x.copy()
        out[f"s{step}_avgvel0"] = r.get(ref.F_AVGVEL)
        npass = 0
        for ps in range(6):
            r.record(True)
            n = r.phase(ref.PH_PROXIMITY_DETECT if ps == 0 else ref.PH_COLLISION_DETECT)
            k = f"s{step}_p{ps}_"
            out[k + "pairs"] = r.pairs()
            out[k + "count"] = np.int64(n)
            out[k + "imp"] = r.get(ref.F_IMP)
            out[k + "fric"] = r.get(ref.F_FRIC)
            out[k + "cnt"] = r.geti(ref.I_CNT)
            out[k + "imp_rg"] = r.get(ref.F_IMP_RG)
            out[k + "cnt_rg"] = r.geti(ref.I_CNT_RG)
            r.phase(ref.PH_APPLY)
            out[k + "avgvel"] = r.get(ref.F_AVGVEL)
            npass += 1
            if ps > 0 and n == 0:
                break
        out[f"s{step}_npass"] = np.int32(npass)
        nzone = 0
        if name.endswith("_zone") and n > 0:  # computeImpactZone, dcollid.cpp:227-265, phase by phase
            r.phase(ref.PH_IMPZONE_ON)
            while True:
                r.record(True)
                n = r.phase(ref.PH_COLLISION_DETECT)
                k = f"s{step}_z{nzone}_"
                out[k + "pairs"] = r.pairs()
                out[k + "count"] = np.int64(n)
                r.phase(ref.PH_APPLY)
                out[k + "avgvel"] = r.get(ref.F_AVGVEL)
                out[k + "zones"] = np.int32(r.phase(ref.PH_ZONE_VELOCITY))
                out[k + "zvel"] = r.get(ref.F_AVGVEL)
                nzone += 1
                if n == 0:
                    break
            r.phase(ref.PH_IMPZONE_OFF)
        out[f"s{step}_nzone"] = np.int32(nzone)
        r.phase(ref.PH_BOUNDARY)
        r.phase(ref.PH_FINAL_POSITION)
        r.phase(ref.PH_FINAL_VELOCITY)
        x, vel = r.get(ref.F_COORDS), r.get(ref.F_VEL)
        out[f"s{step}_x"] = x.copy()
        out[f"s{step}_vel"] = vel.copy()
        out[f"s{step}_has"] = r.geti(ref.I_HAS_COLLSN)
    np.savez_compressed(path, **out)
    print(f"scene_{name}.npz: {nsteps} steps, {os.path.getsize(path) / 1024:.0f} KiB")


def make_strain(path):
    """reduceSuperelast (dcollid.cpp:485-596) of the compiled reference on kicked velocity fields."""
    out = {}
    for name, mk in STRAIN_SCENES.items():
        sc = mk()
        r = ref.RefSolver(sc)
        r.set_rest_lengths(sc.x)
        r.set_state(sc.x, sc.x + sc.dt * sc.vel, sc.vel)
        r.assemble(sc.dt)
        for case in range(3):
            av = strain_inputs(sc, case)
            r.put(ref.F_AVGVEL, av)
            r.phase(ref.PH_STRAIN_LIMIT)
            out[f"{name}_{case}_out"] = r.get(ref.F_AVGVEL)
    np.savez_compressed(path, **out)
    print(f"strain.npz: {len(out)} cases, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    if not ref.available():
        raise SystemExit("oracle/_ref/libcollision_ref.so missing: run `make -C oracle ref` (needs /root/reference)")
    if not sys.argv[1:]:
        make_features(os.path.join(HERE, "features.npz"))
    if not sys.argv[1:] or "strain" in sys.argv[1:]:
        make_strain(os.path.join(HERE, "strain.npz"))
    only = sys.argv[1:]
    for name in SCENES:
        if not only or name in only:
            make_scene(name, os.path.join(HERE, f"scene_{name}.npz"))
