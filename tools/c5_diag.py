"""Diagnostic: one step of a config-5 family member with the per-phase trace (CLSN_TRACE=1)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from collision_b200 import scenes
from collision_b200.solver import CollisionSolver3d
L, n, side, lev = [int(v) for v in sys.argv[1:5]]
sc = scenes.cloth_spheres(L, n, side, lev)
s = CollisionSolver3d(impact_zones=False, strain_limiting=False)
CollisionSolver3d.set_params_from(sc.params)
s.assembleFromInterface(sc, sc.dt)
print("tris", sc.T, file=sys.stderr, flush=True)
x, xn = sc.x.copy(), sc.x_new()
for it in range(3):
    t = time.time()
    s.upload(x, xn)
    st = s.resolve_device()
    print(f"step {it}: {time.time() - t:.3f} s host, {st['ms_total']:.2f} ms device;", "prox", st["proximity"]["candidates"], st["proximity"]["contacts"],
          "ccd", [(p["candidates"], p["pairs_tested"], p["features"], p["contacts"], p["contributions"]) for p in st["ccd"]], file=sys.stderr, flush=True)
