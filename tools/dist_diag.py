"""Diagnostic: the in-library multi-GPU step under torchrun with progress markers (stderr)."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from collision_b200 import scenes
from collision_b200.solver import CollisionSolver3d
from collision_b200.dist import enable_library_exchange
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
def log(*a): print(f"[r{dist.get_rank()} {time.time() % 1000:.2f}]", *a, file=sys.stderr, flush=True)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 251
sc = scenes.layered_cloth(8, n)
s = CollisionSolver3d(device=local, impact_zones=False, strain_limiting=False)
CollisionSolver3d.set_params_from(sc.params)
s.assembleFromInterface(sc, sc.dt); log("assembled", sc.T)
enable_library_exchange(s); log("dist init done")
x, xn = sc.x.copy(), sc.x_new()
for it in range(4):
    s.upload(x, xn); log("uploaded", it)
    st = s.resolve_device(); log("step", it, "ms", round(st["ms_total"], 2), [p["true_pairs"] for p in st["ccd"]])
dist.barrier(); log("done")
