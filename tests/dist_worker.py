"""torchrun worker of tests/test_gpu_multi.py: every rank runs the distributed step (both exchange
flavours) and an ordinary single-GPU solver on its own device, and compares them bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from collision_b200 import scenes  # noqa: E402
from collision_b200.dist import DistributedSolver  # noqa: E402
from collision_b200.solver import CollisionSolver3d  # noqa: E402
from parity_util import same_bits  # noqa: E402


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    for name, sc in (("mixed", scenes.mixed()), ("layered", scenes.layered_cloth(4, 24, seed=99)),
                     ("spheres", scenes.cloth_spheres(n_layers=2, n=17, n_side=2, level=1, seed=31))):
        for mode in ("owner", "gather"):
            CollisionSolver3d.set_params_from(sc.params)
            one = CollisionSolver3d(device=local)   # impact zones and strain limiting on, on both sides
            one.assembleFromInterface(sc, sc.dt)
            many = CollisionSolver3d(device=local)
            many.assembleFromInterface(sc, sc.dt)
            stepper = DistributedSolver(many, mode=mode)
            x, vel = sc.x.copy(), sc.vel.copy()
            contacts = 0
            zone_iters = 0
            for step in range(3):
                xn = x + sc.dt * vel
                xg, vg = xn.copy(), vel.copy()
                one.resolveCollision(x, xg, vg)
                many.upload(x, xn)
                st = stepper.resolve_device()
                xm, avm, hasm = many.download()
                vm = vel.copy()
                vm[hasm != 0] = avm[hasm != 0]
                assert same_bits(xm, xg), (name, mode, step, "positions")
                assert same_bits(vm, vg), (name, mode, step, "velocities")
                assert st["n_ccd_passes"] == one.last_stats["n_ccd_passes"]
                assert [p["true_pairs"] for p in st["ccd"]] == [p["true_pairs"] for p in one.last_stats["ccd"]]
                for k in ("zone_iterations", "zones", "strain_sweeps", "strain_edges"):
                    assert st[k] == one.last_stats[k], (name, mode, step, k)
                zone_iters += st["zone_iterations"]
                contacts += sum(p["true_pairs"] for p in st["ccd"])
                x, vel = xg, vg
            assert contacts > 0
            assert name != "mixed" or zone_iters > 0   # the mixed scene enters the impact-zone fail-safe in step 2
            one.close()
            many.close()
    # ---- the in-library multi-GPU step (csrc/dist.cuh): records pushed to their owners through NVLink peer memory from
    # inside the kernels, NCCL issued by the library, one read-back per step -- against the single-GPU step, bit for bit
    from collision_b200.dist import enable_library_exchange
    for name, sc in (("layered", scenes.layered_cloth(4, 24, seed=99)), ("two_sheets", scenes.two_sheets(n=20)),
                     ("spheres", scenes.cloth_spheres(n_layers=2, n=17, n_side=2, level=1, seed=31)),
                     ("ball_plane", scenes.ball_plane(gap=2e-4))):
        CollisionSolver3d.set_params_from(sc.params)
        one = CollisionSolver3d(device=local, impact_zones=False)
        one.assembleFromInterface(sc, sc.dt)
        many = CollisionSolver3d(device=local, impact_zones=False)
        many.assembleFromInterface(sc, sc.dt)
        enable_library_exchange(many)
        assert many.dist_nranks() == dist.get_world_size()
        x, vel = sc.x.copy(), sc.vel.copy()
        contacts = 0
        for step in range(4):
            xn = x + sc.dt * vel
            xg, vg = xn.copy(), vel.copy()
            has1 = one.resolveCollision(x, xg, vg)
            xm, vm = xn.copy(), vel.copy()
            hasm = many.resolveCollision(x, xm, vm)
            assert same_bits(xm, xg), (name, "library", step, "positions")
            assert same_bits(vm, vg), (name, "library", step, "velocities")
            assert np.array_equal(has1, hasm)
            a, b = many.last_stats, one.last_stats
            assert a["n_ccd_passes"] == b["n_ccd_passes"] and a["still_colliding"] == b["still_colliding"]
            for pa, pb in zip([a["proximity"]] + a["ccd"], [b["proximity"]] + b["ccd"]):
                for k in ("true_pairs", "contacts", "contributions"):
                    assert pa[k] == pb[k], (name, "library", step, k, pa[k], pb[k])
            assert (a["strain_sweeps"], a["strain_edges"]) == (b["strain_sweeps"], b["strain_edges"])
            contacts += sum(p["true_pairs"] for p in a["ccd"])
            x, vel = xg, vg
        assert contacts > 0, name
        one.close()
        many.close()
    dist.barrier()
    print("dist ok", dist.get_rank(), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
