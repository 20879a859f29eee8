import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo")); sys.path.insert(0, os.path.join(os.environ.get("GRAFT_REPO_ROOT", "/root/repo"), "tests"))
import numpy as np
from collision_b200 import scenes
from collision_b200.solver import CollisionSolver3d
for sc in (scenes.mixed(), scenes.two_sheets(n=14), scenes.cloth_spheres(n_layers=2, n=13, n_side=2, level=1, seed=31), scenes.string_string(dt=0.01, gap=0.003)):
    CollisionSolver3d.set_params_from(sc.params)
    g = CollisionSolver3d(); g.assembleFromInterface(sc, sc.dt)
    x, vel = sc.x.copy(), sc.vel.copy()
    for step in range(3):
        xg = x + sc.dt * vel
        g.resolveCollision(x, xg, vel)
        x = xg
    print(sc.name, g.last_stats["n_ccd_passes"], float(np.abs(x).sum()))
    g.close()
print("done")
