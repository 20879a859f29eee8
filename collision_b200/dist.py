"""Multi-GPU step: one process per GPU, mesh + BVH replicated, query leaves sliced.

One exchange per pass, two flavours (both over torch.distributed: NCCL/NVLink on GPUs, gloo in the CPU
tests of the exchange logic):

* "owner" (default): vertex v belongs to rank v // ceil(V/G).  Each rank buckets its impulse records by
  owner (clsn_bucket_records), an all-to-all delivers them, every rank reduces the records of ITS OWN
  vertices in canonical key order, and the ranks all-gather their avgVel / has_collsn / touched slices.
  Traffic per rank ~ records/G instead of all records, and the reduction itself is split G ways.
* "gather": all-gather of every rank's records, every rank reduces the union.

Because the per-point reduction sorts by the canonical key (ea, eb, feature) before summing, the result
is bit-identical on every rank and for every world size -- the N-GPU result equals the 1-GPU result bit
for bit (tests/test_gpu_parity.py::test_sliced_equals_whole, tests/test_host_cpu.py for the exchange).
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


def exchange_by_owner(send: torch.Tensor, send_counts, group=None):
    """All-to-all of byte buffers bucketed by destination rank.  `send` holds, contiguously, send_counts[r]
    BYTES for every rank r.  Returns (received bytes concatenated by source rank, per-source byte counts)."""
    world = dist.get_world_size(group)
    sc = torch.tensor(list(send_counts), dtype=torch.int64, device=send.device)
    allc = [torch.zeros_like(sc) for _ in range(world)]
    dist.all_gather(allc, sc, group=group)
    rank = dist.get_rank(group)
    recv_counts = [int(allc[src][rank].item()) for src in range(world)]
    recv = send.new_empty(sum(recv_counts))
    if send.is_cuda:
        dist.all_to_all_single(recv, send[: int(sc.sum().item())], output_split_sizes=recv_counts,
                               input_split_sizes=[int(v) for v in send_counts], group=group)
    else:  # gloo has no all_to_all_single: the CPU test of the protocol goes through point-to-point lists
        ins = list(torch.split(send[: int(sc.sum().item())], [int(v) for v in send_counts]))
        outs = [send.new_empty(n) for n in recv_counts]
        reqs = []
        for peer in range(world):
            if peer == rank:
                outs[peer].copy_(ins[peer])
                continue
            reqs.append(dist.isend(ins[peer].contiguous(), peer, group=group))
            reqs.append(dist.irecv(outs[peer], peer, group=group))
        for r in reqs:
            r.wait()
        recv = torch.cat(outs) if outs else recv
    return recv, recv_counts


def enable_library_exchange(solver: CollisionSolver3d, group=None):
    """Switch `solver` (already assembled) to the in-library multi-GPU step (csrc/dist.cuh): torch.distributed is only
    used to hand rank 0's NCCL unique id to the other ranks; records then travel through NVLink peer memory from inside
    the narrow-phase kernels and NCCL collectives issued by the library itself."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [CollisionSolver3d.dist_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    solver.dist_init(rank, world, box[0])
    return solver


class DistributedSolver:
    """resolveCollision across the ranks of a process group (dcollid.cpp:317-362 restated as a host
    loop over the per-phase C ABI, with one record exchange per pass)."""

    def __init__(self, solver: CollisionSolver3d, group=None, mode: str = "owner"):
        self.s = solver
        self.group = group
        self.mode = mode
        self.small_pass_records = 65536   # passes with fewer records (all ranks together) skip the all-to-all
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        assert self.world <= 64, "clsn_bucket_records keeps 64 owner buckets"
        c = solver.ctx
        c.check(c.L.clsn_set_slice(c.h, self.rank, self.world))
        self.device = torch.device("cuda", torch.cuda.current_device())
        self._keep = None
        # share torch's stream: the NCCL exchange is then ordered with the library's kernels without host syncs
        # (torch's default stream is the legacy stream, handle 0: pass cudaStreamLegacy = 0x1 for it, NULL means 'own')
        c.check(c.L.clsn_set_stream(c.h, C.c_void_p(torch.cuda.current_stream().cuda_stream or 1)))

    def _state_views(self):
        c = self.s.ctx
        pa, ph, pd = C.c_void_p(), C.c_void_p(), C.c_void_p()
        c.check(c.L.clsn_state_device_ptrs(c.h, C.byref(pa), C.byref(ph), C.byref(pd)))
        V = c.V
        return (torch.as_tensor(_DevPtr(pa.value, 32 * V), device=self.device)[: 32 * V],
                torch.as_tensor(_DevPtr(ph.value, V), device=self.device)[:V],
                torch.as_tensor(_DevPtr(pd.value, V), device=self.device)[:V])

    def _exchange_owner(self, rigidify=True):
        """bucket -> all-to-all -> reduce own vertices -> all-gather state slices -> bodies / rigid bodies.
        Three collectives per pass (header all-gather, record all-to-all, state all-gather) and one host
        read of the header; everything else is stream-ordered with the library's kernels."""
        c = self.s.ctx
        G, V = self.world, c.V
        per = (V + G - 1) // G
        counts = (C.c_int64 * G)()
        ps = C.c_void_p()
        c.check(c.L.clsn_bucket_records(c.h, G, counts, C.byref(ps)))
        pp, pb = C.c_void_p(), C.c_void_p()
        npr, nbr, ntrue = C.c_int64(), C.c_int64(), C.c_int64()
        c.check(c.L.clsn_export_records(c.h, C.byref(pp), C.byref(npr), C.byref(pb), C.byref(nbr), C.byref(ntrue)))
        # header: my per-owner record counts, my body-record count, my true-pair count
        hdr = torch.tensor([int(counts[r]) for r in range(G)] + [nbr.value, ntrue.value], dtype=torch.int64, device=self.device)
        allh = torch.empty((G, G + 2), dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(allh, hdr, group=self.group)
        allh = allh.cpu()
        n_true = int(allh[:, G + 1].sum())
        per_rank_total = [int(allh[src, :G].sum()) for src in range(G)]
        if sum(per_rank_total) <= self.small_pass_records and int(allh[:, G].sum()) == 0:
            # small pass (late CCD passes, contact-free proximity): cheaper to all-gather the few records and let
            # every rank reduce all of them -- one collective instead of two, no state exchange
            mx = max(per_rank_total) * POINT_RECORD_BYTES
            allp = torch.empty(0, dtype=torch.uint8, device=self.device)
            if mx > 0:
                nb = npr.value * POINT_RECORD_BYTES
                mine = torch.zeros(mx, dtype=torch.uint8, device=self.device)
                mine[:nb] = torch.as_tensor(_DevPtr(pp.value, nb), device=self.device)[:nb]
                full = torch.empty((G, mx), dtype=torch.uint8, device=self.device)
                dist.all_gather_into_tensor(full, mine, group=self.group)
                allp = torch.cat([full[r, : per_rank_total[r] * POINT_RECORD_BYTES] for r in range(G)])
            self._keep = (allp,)
            c.check(c.L.clsn_import_records(c.h, allp.data_ptr() if allp.numel() else None, allp.numel() // POINT_RECORD_BYTES,
                                            None, 0))
            c.check(c.L.clsn_apply_stage(c.h, 1, 0))
            c.check(c.L.clsn_apply_stage(c.h, 2, 1 if rigidify else 0))
            return n_true
        send_b = [int(counts[r]) * POINT_RECORD_BYTES for r in range(G)]
        recv_b = [int(allh[src, self.rank]) * POINT_RECORD_BYTES for src in range(G)]
        send = torch.as_tensor(_DevPtr(ps.value, sum(send_b)), device=self.device)[: sum(send_b)]
        recv = torch.empty(sum(recv_b), dtype=torch.uint8, device=self.device)
        dist.all_to_all_single(recv, send, output_split_sizes=recv_b, input_split_sizes=send_b, group=self.group)
        allb = torch.empty(0, dtype=torch.uint8, device=self.device)
        if int(allh[:, G].sum()) > 0:
            # body records (rigid-rigid contacts) are few: all-gather, every rank reduces them identically
            nb_b = nbr.value * BODY_RECORD_BYTES
            allb = gather_varlen(torch.as_tensor(_DevPtr(pb.value, nb_b), device=self.device)[:nb_b], self.group)
        self._keep = (recv, allb)
        c.check(c.L.clsn_import_records(c.h, recv.data_ptr() if recv.numel() else None, recv.numel() // POINT_RECORD_BYTES,
                                        allb.data_ptr() if allb.numel() else None, allb.numel() // BODY_RECORD_BYTES))
        c.check(c.L.clsn_apply_stage(c.h, 1, 0))
        # every rank now holds the final avgVel / flags of its own vertex range: one all-gather of packed slices
        av, has, dirty = self._state_views()
        lo = min(V, self.rank * per)      # V < per * (G - 1): the last ranks own nothing
        hi = min(V, (self.rank + 1) * per)
        mine = torch.zeros(per * 34, dtype=torch.uint8, device=self.device)
        n = max(0, hi - lo)
        mine[: n * 32] = av[lo * 32: hi * 32]
        mine[per * 32: per * 32 + n] = has[lo:hi]
        mine[per * 33: per * 33 + n] = dirty[lo:hi]
        full = torch.empty((G, per * 34), dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(full, mine, group=self.group)
        av.copy_(full[:, : per * 32].reshape(-1)[: V * 32])
        has.copy_(full[:, per * 32: per * 33].reshape(-1)[:V])
        dirty.copy_(full[:, per * 33: per * 34].reshape(-1)[:V])
        self._keep = (recv, allb, full)
        c.check(c.L.clsn_apply_stage(c.h, 2, 1 if rigidify else 0))
        return n_true

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
        self._keep = (allp, allb)  # must outlive clsn_apply
        c.check(c.L.clsn_import_records(c.h, allp.data_ptr() if allp.numel() else None,
                                        allp.numel() // POINT_RECORD_BYTES,
                                        allb.data_ptr() if allb.numel() else None, allb.numel() // BODY_RECORD_BYTES))
        return int(t.item())

    def resolve_device(self):
        """Same contract as CollisionSolver3d.resolve_device on state already uploaded."""
        s = self.s
        stats = dict(proximity=None, ccd=[], n_ccd_passes=0, has_collision=False, still_colliding=False)
        owner = self.mode == "owner"
        s.avg_velocity()
        st = s.detect(PROXIMITY)
        st["true_pairs"] = self._exchange_owner() if owner else self._exchange()
        stats["proximity"] = st
        if not owner:
            s.apply(True)
        is_collision, niter, cd = True, 1, 0
        while is_collision:
            st = s.detect(COLLISION)
            st["true_pairs"] = self._exchange_owner() if owner else self._exchange()
            is_collision = st["true_pairs"] > 0
            if cd == 0 and is_collision:
                stats["has_collision"] = True
            stats["ccd"].append(st)
            cd += 1
            if not owner:
                s.apply(True)
            niter += 1
            if niter > MAX_CCD_PASSES:
                break
        stats["n_ccd_passes"] = cd
        stats["still_colliding"] = is_collision
        stats["zone_iterations"] = stats["zones"] = stats["strain_sweeps"] = stats["strain_edges"] = 0
        if is_collision and s.impact_zones:
            # computeImpactZone (dcollid.cpp:227-265) needs every pair's first hit in canonical order: the state is
            # replicated, so each rank runs the fail-safe on the whole mesh (identical results, no exchange)
            c = s.ctx
            c.check(c.L.clsn_set_slice(c.h, 0, 1))
            try:
                zs = s.computeImpactZone(s.max_zone_iterations)
            finally:
                c.check(c.L.clsn_set_slice(c.h, self.rank, self.world))
            stats["zone_iterations"], stats["zones"] = zs["iterations"], zs["zones"]
        s.boundary()
        s.final_position()
        if s.strain_limiting:   # reduceSuperelast (dcollid.cpp:355), replicated like the state it works on
            stats["strain_sweeps"], stats["strain_edges"] = s.reduceSuperelast()
        s.synchronize()
        return stats
