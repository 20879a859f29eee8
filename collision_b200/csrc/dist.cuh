// Multi-GPU layer: one context (process or thread) per B200, mesh + tree replicated, query leaves sliced.
//
// The reference has no multi-process path at all; what has to hold is that N GPUs produce the bits of one GPU.  The
// per-point reduction sums records in canonical key order, so it does not matter WHERE a record was produced -- only that
// every record of a point reaches the rank that reduces that point ("owner computes": vertex v belongs to rank
// v / ceil(V / G)).  Data path of one pass:
//
//   k_emit / k_contact   every impulse record goes to a per-owner region (slot from a per-owner cursor, one atomic per
//                        distinct owner and warp): the rank's own records straight into its receive buffer, the others
//                        into a local staging buffer;
//   k_push_regions       the staged regions are written into the OWNERS' receive buffers through NVLink peer mappings
//                        (cudaIpc between processes): plain coalesced 16-byte stores, 512 B per warp instruction, all
//                        ranks pushing to all owners at once -- an all-to-all without a library call, sizes known only
//                        on the device;
//   k_publish            the per-owner counts go to the owners' headers (8-byte peer stores), overflow flags are folded;
//   ncclAllReduce        of the pass's counter block: the global contact / true-pair counts gate the next pass on the
//                        device, and the collective doubles as the barrier after which every peer store has landed;
//   owner reduce         k_count_regions -> scan -> k_scatter_regions -> k_reduce_points on the owner's vertex range;
//   ncclAllGather        (in place) of avgVel, has_collsn and the touched flags: the state is whole again everywhere.
//
// No host read-back anywhere: counts stay on the device, buffers have fixed capacity, and an overflow is seen by all ranks
// in the all-reduced block at the end of the step (clsn.cu: resolve_impl), which then grow and repeat together.
// NCCL is loaded with dlopen (libnccl.so.2: the copy the host process already uses, e.g. torch's), so a single-GPU user
// never needs it.
#pragma once
#include <dlfcn.h>
#include <nccl.h>   // types only; the entry points are resolved at run time
#include "narrow.cuh"

namespace clsn {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;

    bool load(std::string& err)
    {
        if (lib) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names)
            if ((lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
        if (!lib) { err = std::string("NCCL not found: ") + dlerror(); return false; }
#define CLSN_NCCL_SYM(field, name)                                                     \
    field = reinterpret_cast<decltype(field)>(dlsym(lib, name));                      \
    if (!field) { err = std::string("NCCL symbol missing: ") + name; return false; }
        CLSN_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        CLSN_NCCL_SYM(CommInitRank, "ncclCommInitRank")
        CLSN_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        CLSN_NCCL_SYM(AllReduce, "ncclAllReduce")
        CLSN_NCCL_SYM(AllGather, "ncclAllGather")
        CLSN_NCCL_SYM(GroupStart, "ncclGroupStart")
        CLSN_NCCL_SYM(GroupEnd, "ncclGroupEnd")
        CLSN_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef CLSN_NCCL_SYM
        return true;
    }
};

#define CLSN_MAX_RANKS 64

// after the emitting kernels: counts to the owners' headers, totals and overflow flags into the pass's counter block.
// caps[]: capacities of the local work lists in the order of ovf_ctr[] (device-side overflow detection, so that all ranks
// learn about a too-small list from the all-reduced block without a host exchange).
struct PublishCaps {
    long long pairs, feats, unc, hits, brec, region;
};
__global__ void k_publish(int nranks, int me, const unsigned long long* __restrict__ send_cnt, unsigned long long* const* peer_hdr,
                          unsigned long long* ctr, unsigned long long* maxblk, PublishCaps caps)
{
    const int r = threadIdx.x;
    unsigned long long n = 0;
    if (r < nranks) {
        n = send_cnt[r];
        peer_hdr[r][me] = n < (unsigned long long)caps.region ? n : (unsigned long long)caps.region;   // what was really stored
    }
    unsigned long long tot = n, mx = n;
    for (int o = 16; o > 0; o >>= 1) {
        tot += __shfl_xor_sync(0xffffffffu, tot, o);
        const unsigned long long t = __shfl_xor_sync(0xffffffffu, mx, o);
        mx = t > mx ? t : mx;
    }
    __shared__ unsigned long long s_tot[2], s_mx[2];
    if ((r & 31) == 0) { s_tot[r >> 5] = tot; s_mx[r >> 5] = mx; }
    __syncthreads();
    if (r == 0) {
        tot = s_tot[0] + s_tot[1];
        mx = s_mx[0] > s_mx[1] ? s_mx[0] : s_mx[1];
        ctr[CTR_PREC] = tot;
        bool ovf = mx > (unsigned long long)caps.region;
        ovf = ovf || ctr[CTR_PAIRS] > (unsigned long long)caps.pairs;
        ovf = ovf || ctr[CTR_MAX_FEATS] > (unsigned long long)caps.feats;   // per chunk of the pair list (k_fold_chunk)
        ovf = ovf || ctr[CTR_MAX_UNC] > (unsigned long long)caps.unc;
        ovf = ovf || ctr[CTR_MAX_HITS] > (unsigned long long)caps.hits;
        ovf = ovf || ctr[CTR_BREC] > (unsigned long long)caps.brec;
        ctr[CTR_OVF] = ovf ? 1ull : 0ull;
        maxblk[0] = mx;              // all-reduced with MAX: the region capacity every rank needs
        maxblk[1] = ctr[CTR_BREC];   // ... and the body-record capacity
    }
}

// Records staged per owner in local memory -> the owners' receive regions, as long contiguous runs: every warp
// instruction writes 512 consecutive bytes, which is what NVLink wants (measured: 64-byte records stored one by one from
// the emitting kernels reached ~140 GB/s per GPU, a fraction of the link).  blockIdx.y = owner; the own region was
// written in place by the emitting kernels.
__global__ void k_push_regions(int me, long long cap_region, const PointRec* __restrict__ stage, PointRec* const* peer_region,
                               const unsigned long long* __restrict__ send_cnt)
{
    const int r = blockIdx.y;
    if (r == me) return;
    long long n = (long long)send_cnt[r];
    if (n > cap_region) n = cap_region;
    const uint4* src = reinterpret_cast<const uint4*>(stage + (size_t)r * cap_region);
    uint4* dst = reinterpret_cast<uint4*>(peer_region[r]);
    const long long n16 = n * 4;   // 16-byte chunks
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x) dst[i] = src[i];
}

// ---- owner side: the received records live in nranks regions of cap_region slots, region s holding hdr[s] records
__global__ void k_count_regions(const PointRec* __restrict__ recv, long long cap_region, const unsigned long long* __restrict__ hdr,
                                int* cnt)
{
    const long long n = (long long)hdr[blockIdx.y];
    const PointRec* rec = recv + (size_t)blockIdx.y * cap_region;
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x)
        atomicAdd(cnt + rec[r].point, 1);
}
__global__ void k_scatter_regions(const PointRec* __restrict__ recv, long long cap_region, const unsigned long long* __restrict__ hdr,
                                  const int* __restrict__ offs, int* fill, int* __restrict__ perm, unsigned long long* __restrict__ skey)
{
    const long long n = (long long)hdr[blockIdx.y];
    const long long base = (long long)blockIdx.y * cap_region;
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
        const ulonglong2 h = *reinterpret_cast<const ulonglong2*>(recv + base + r);
        const int p = (int)(unsigned)h.y;
        const int slot = offs[p] + atomicAdd(fill + p, 1);
        perm[slot] = (int)(base + r);
        skey[slot] = h.x;
    }
}

// all-gathered body records (fixed capacity per rank) -> one dense list + per-body counts, identical on every rank
__global__ void k_compact_bodies(int nranks, long long cap_brec, const BodyRec* __restrict__ all, const unsigned long long* __restrict__ counts,
                                 BodyRec* __restrict__ out, unsigned long long* n_out, int* cnt_rg)
{
    // ranks in order, so that the list (and k_reduce_bodies' key-ordered sums) do not depend on timing
    __shared__ unsigned long long s_base;
    unsigned long long base = 0;
    for (int r = 0; r < nranks; ++r) {
        const unsigned long long n = counts[(size_t)r * 4 + 1] < (unsigned long long)cap_brec ? counts[(size_t)r * 4 + 1] : (unsigned long long)cap_brec;
        for (unsigned long long i = threadIdx.x; i < n; i += blockDim.x) {
            const BodyRec b = all[(size_t)r * cap_brec + i];
            out[base + i] = b;
            atomicAdd(cnt_rg + b.body, 1);
        }
        base += n;
    }
    if (threadIdx.x == 0) { s_base = base; *n_out = s_base; }
}

} // namespace clsn
