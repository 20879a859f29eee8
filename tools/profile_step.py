"""Profiling target: N steps of config 4 through clsn_resolve (inputs resident in HBM), nothing else.
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \\
        --log-file gpurun_out/launches.csv python tools/profile_step.py 3
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from collision_b200 import scenes
from collision_b200.solver import CollisionSolver3d
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
sc = scenes.layered_cloth(8, 251)
s = CollisionSolver3d(impact_zones=False, strain_limiting=False)
CollisionSolver3d.set_params_from(sc.params)
s.assembleFromInterface(sc, sc.dt)
dev = torch.device("cuda", 0)
d_xo = torch.from_numpy(sc.x.copy()).to(dev)
d_xn = torch.from_numpy(sc.x_new()).to(dev)
for it in range(n):
    s.upload_device(d_xo.data_ptr(), d_xn.data_ptr())
    st = s.resolve_device()
print("ms", st["ms_total"], file=sys.stderr)
