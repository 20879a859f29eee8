"""Multi-GPU step: one process per GPU, mesh + BVH replicated, query leaves sliced, impulse records
all-gathered with torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests of the
exchange logic), every rank reduces the union in canonical key order.

Because the per-point reduction sorts by the canonical key (ea, eb, feature) before summing, the
union of the ranks' record sets gives bit-identical avgVel on every rank and for every world size
-- the N-GPU result equals the 1-GPU result bit for bit (tests/test_gpu_parity.py::test_sliced_equals_whole,
tests/test_dist_gloo.py for the exchange itself).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from .solver import COLLISION, MAX_CCD_PASSES, PROXIMITY, CollisionSolver3d

POINT_RECORD_BYTES = 64
BODY_RECORD_BYTES = 48


class _DevPtr:
    """Expose a raw device pointer of the library to torch (no copy)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (max(nbytes, 1),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2}


def gather_varlen(local: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather byte buffers of different lengths: returns rank0 | rank1 | ... concatenated.
    `local` is a 1-D uint8 tensor (CUDA for NCCL, CPU for gloo).  One small all-gather of the lengths,
    one padded all-gather of the payload."""
    world = dist.get_world_size(group)
    n = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes)
    if mx == 0:
        return local.new_empty(0)
    padded = local.new_zeros(mx)
    padded[: local.numel()] = local
    out = [local.new_empty(mx) for _ in range(world)]
    dist.all_gather(out, padded, group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)])


class DistributedSolver:
    """resolveCollision across the ranks of a process group (dcollid.cpp:317-362 restated as a host
    loop over the per-phase C ABI, with one record exchange per pass)."""

    def __init__(self, solver: CollisionSolver3d, group=None):
        self.s = solver
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        c = solver.ctx
        c.check(c.L.clsn_set_slice(c.h, self.rank, self.world))
        self.device = torch.device("cuda", torch.cuda.current_device())
        self._keep = None

    def _exchange(self):
        c = self.s.ctx
        pp, pb = C.c_void_p(), C.c_void_p()
        npr, nbr, ntrue = C.c_int64(), C.c_int64(), C.c_int64()
        c.check(c.L.clsn_export_records(c.h, C.byref(pp), C.byref(npr), C.byref(pb), C.byref(nbr), C.byref(ntrue)))
        nb_p = npr.value * POINT_RECORD_BYTES
        nb_b = nbr.value * BODY_RECORD_BYTES
        lp = torch.as_tensor(_DevPtr(pp.value, nb_p), device=self.device)[:nb_p]
        lb = torch.as_tensor(_DevPtr(pb.value, nb_b), device=self.device)[:nb_b]
        allp = gather_varlen(lp, self.group)
        allb = gather_varlen(lb, self.group)
        t = torch.tensor([ntrue.value], dtype=torch.int64, device=self.device)
        dist.all_reduce(t, group=self.group)
        torch.cuda.synchronize()
        self._keep = (allp, allb)  # must outlive clsn_apply
        c.check(c.L.clsn_import_records(c.h, allp.data_ptr() if allp.numel() else None,
                                        allp.numel() // POINT_RECORD_BYTES,
                                        allb.data_ptr() if allb.numel() else None, allb.numel() // BODY_RECORD_BYTES))
        return int(t.item())

    def resolve_device(self):
        """Same contract as CollisionSolver3d.resolve_device on state already uploaded."""
        s = self.s
        stats = dict(proximity=None, ccd=[], n_ccd_passes=0, has_collision=False, still_colliding=False)
        s.avg_velocity()
        st = s.detect(PROXIMITY)
        st["true_pairs"] = self._exchange()
        stats["proximity"] = st
        s.apply(True)
        is_collision, niter, cd = True, 1, 0
        while is_collision:
            st = s.detect(COLLISION)
            st["true_pairs"] = self._exchange()
            is_collision = st["true_pairs"] > 0
            if cd == 0 and is_collision:
                stats["has_collision"] = True
            stats["ccd"].append(st)
            cd += 1
            s.apply(True)
            niter += 1
            if niter > MAX_CCD_PASSES:
                break
        stats["n_ccd_passes"] = cd
        stats["still_colliding"] = is_collision
        s.boundary()
        s.final_position()
        s.synchronize()
        return stats
