// Deterministic impulse reduction + per-vertex passes (sm_100a).
//
// The reference accumulates "collsnImpulse += ..., friction += ..., collsn_num += 1" through
// pointers while it walks its tree (dcollid3d.cpp:1053-1079, 1235-1284) and then applies
// avgVel += (imp + fric)/num once per point (updateAverageVelocity, dcollid.cpp:677-751).
// Here the narrow phase has emitted 64-byte records; they are grouped per point with a
// counting sort (count -> exclusive scan -> scatter) and each point's records are summed
// sequentially in canonical key order (ea, eb, feature) -- no floating-point atomics, bit-identical
// from run to run and for any number of ranks.
#pragma once
#include "narrow.cuh"
#include "lbvh.cuh"

namespace clsn {

// records -> per-point slots.  fill[] must be zero.
// hdr: the (key, point) headers of the records as a dense 16-byte array when the emitting kernels wrote one (the
// context's own list), else null: then the header is read out of the 64-byte record, which costs the whole record in
// DRAM traffic (measured: k_scatter read 1.9 GB per step for 0.4 GB of headers).
__global__ void k_scatter(const PointRec* __restrict__ rec, const ulonglong2* __restrict__ hdr, const unsigned long long* __restrict__ n_rec_ptr,
                          long long cap, const int* __restrict__ offs, int* fill, int* __restrict__ perm,
                          unsigned long long* __restrict__ skey)
{
    const long long n = (long long)*n_rec_ptr;
    if (n > cap) return;   // the record list overflowed: offs[] describes records that were never stored; the step is repeated
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
        const ulonglong2 h = hdr ? hdr[r] : *reinterpret_cast<const ulonglong2*>(rec + r);
        const int p = (int)(unsigned)h.y;
        const int slot = offs[p] + atomicAdd(fill + p, 1);
        perm[slot] = (int)r;
        skey[slot] = h.x;
    }
}


// per-point / per-body record counts of an externally gathered record set (multi-GPU import)
__global__ void k_count_records(const PointRec* __restrict__ prec, long long nprec, const BodyRec* __restrict__ brec,
                                long long nbrec, int* cnt, int* cnt_rg)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < nprec; r += stride) atomicAdd(cnt + prec[r].point, 1);
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < nbrec; r += stride) atomicAdd(cnt_rg + brec[r].body, 1);
}

// Multi-GPU owner-computes exchange: bucket this rank's point records by the rank that owns the point
// (contiguous vertex ranges of `per_rank` vertices).  Two passes: histogram, then scatter with one
// running offset per bucket (order inside a bucket is free -- the reduction sorts by key anyway).
__global__ void k_owner_count(const PointRec* __restrict__ rec, long long n, int per_rank, unsigned long long* counts)
{
    __shared__ unsigned s_c[64];
    if (threadIdx.x < 64) s_c[threadIdx.x] = 0;
    __syncthreads();
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x)
        atomicAdd(&s_c[rec[r].point / per_rank], 1u);
    __syncthreads();
    if (threadIdx.x < 64 && s_c[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)s_c[threadIdx.x]);
}
__global__ void k_owner_scatter(const PointRec* __restrict__ rec, long long n, int per_rank, unsigned long long* cursor,
                                PointRec* __restrict__ out)
{
    // one global atomic per (tile, owner): ranks inside the tile come from shared-memory counters
    __shared__ unsigned s_c[64];
    __shared__ unsigned long long s_base[64];
    for (long long tile = (long long)blockIdx.x * blockDim.x; tile < n; tile += (long long)gridDim.x * blockDim.x) {
        if (threadIdx.x < 64) s_c[threadIdx.x] = 0;
        __syncthreads();
        const long long r = tile + threadIdx.x;
        ulonglong2 a, b, c2, d;
        int owner = 0;
        unsigned local = 0;
        if (r < n) {
            const ulonglong2* src = reinterpret_cast<const ulonglong2*>(rec + r);
            a = src[0]; b = src[1]; c2 = src[2]; d = src[3];
            owner = (int)(unsigned)a.y / per_rank;
            local = atomicAdd(&s_c[owner], 1u);
        }
        __syncthreads();
        if (threadIdx.x < 64 && s_c[threadIdx.x]) s_base[threadIdx.x] = atomicAdd(&cursor[threadIdx.x], (unsigned long long)s_c[threadIdx.x]);
        __syncthreads();
        if (r < n) {
            ulonglong2* dst = reinterpret_cast<ulonglong2*>(out + (s_base[owner] + local));
            dst[0] = a; dst[1] = b; dst[2] = c2; dst[3] = d;
        }
        __syncthreads();
    }
}

// One warp per point.  Rank the point's keys (all-pairs compare, keys are unique), store the
// records in rank order, then lanes 0..5 each sum one of imp.xyz / fric.xyz sequentially.
// mode 0: apply to avgVel (updateAverageVelocity :707-724); mode 1: write the sums to acc arrays.
// SEG: the records sit in per-point segments already (pipeline 2).  SORTED: perm[] is already in key order per point
// (cub::DeviceSegmentedSort, used when movable rigid bodies are present: a sphere vertex that sweeps through a cloth stack
// collects tens of thousands of records, and the all-pairs ranking is quadratic in that number).
template <bool SEG, bool SORTED = false>
__global__ void __launch_bounds__(256)
k_reduce_points(const PointRec* __restrict__ rec, const int* __restrict__ offs, const int* __restrict__ cnt, int V,
                const int* __restrict__ perm, int* __restrict__ perm_sorted, const unsigned long long* __restrict__ skey,
                const uint8_t* __restrict__ vflags, Vec4* av, uint8_t* has, uint8_t* dirty, int mode, double* __restrict__ acc_imp,
                double* __restrict__ acc_fric, unsigned long long* counters, const unsigned long long* __restrict__ n_rec_ptr,
                long long cap)
{
    if ((long long)*n_rec_ptr > cap) return;   // overflowed list (see k_scatter)
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    for (int p = blockIdx.x * warps_per_block + (threadIdx.x >> 5); p < V; p += gridDim.x * warps_per_block) {
        const int n = cnt[p];
        if (n == 0) continue;
        const int base = offs[p];
        const int* order = perm_sorted;
        if (SORTED) {
            order = perm;
        } else {
            // rank = number of smaller keys
            for (int t = lane; t < n; t += 32) {
                // SEG: the point's records are the contiguous segment rec[base .. base + n) (written there by k_emit<., true>)
                const unsigned long long k = SEG ? rec[base + t].key : skey[base + t];
                int rank = 0;
                for (int u = 0; u < n; ++u) rank += (SEG ? rec[base + u].key : skey[base + u]) < k ? 1 : 0;
                perm_sorted[base + rank] = SEG ? base + t : perm[base + t];
            }
            __syncwarp();
        }
        // One lane per component walks the records in key order.  (Fetching 32 records with the whole warp and handing the
        // addends over by shuffles was measured: 1.7x SLOWER -- twelve shuffles per record cost more issue slots than the
        // two dependent loads cost latency at 64 resident warps per SM.)
        double sum = 0.0;
        if (lane < 6) {
            for (int t = 0; t < n; ++t) {
                const int r = order[base + t];
                const double* val = reinterpret_cast<const double*>(rec + r) + 2;  // imp[3], fric[3]
                sum += val[lane];
            }
        }
        const double fr = __shfl_down_sync(0xffffffffu, sum, 3);  // lanes 0..2 get fric.xyz
        if (mode == 0) {
            if (lane < 3 && !(vflags[p] & 1)) {
                double* a = reinterpret_cast<double*>(av + p) + lane;
                const double v = *a + (sum + fr) / n;
                *a = v;
                if (isinf(v) || isnan(v)) atomicAdd(&counters[CTR_ERROR], 1ull);
                if (lane == 0) { has[p] = 1; dirty[p] = 1; }
            }
        } else if (lane < 6) {
            if (lane < 3) acc_imp[3 * (size_t)p + lane] = sum;
            else acc_fric[3 * (size_t)p + lane - 3] = sum;
        }
        __syncwarp();
    }
}


// keep the segment offsets inside the record buffers when a list overflowed (the step is repeated then, but the segmented
// sort must not run past the end of its arrays in the meantime)
__global__ void k_clamp_offsets(int n, int* offs, int cap)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && offs[i] > cap) offs[i] = cap;
}

// Rigid-rigid contacts: per-body sums in key order.  imp_rg accumulates across passes and steps -- the reference never
// zeroes collsnImpulse_RG (dcollid.cpp:726-733).  The records are brought into (body, key) order by two stable radix
// sorts (key, then body; clsn.cu: reduce_bodies), then one thread per body walks its contiguous run sequentially --
// the summation order is part of the result, the search for it need not be quadratic (a lattice of fast spheres
// produces 10^5 sphere-sphere records per pass).
__global__ void k_body_keys(const BodyRec* __restrict__ rec, const unsigned long long* __restrict__ n_rec_ptr, long long cap,
                            unsigned long long* __restrict__ key, int* __restrict__ idx)
{
    long long n = (long long)*n_rec_ptr;
    if (n > cap) n = cap;
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < cap; r += (long long)gridDim.x * blockDim.x) {
        key[r] = r < n ? rec[r].key : ~0ull;   // padding sorts behind every record
        idx[r] = r < n ? (int)r : -1;
    }
}
__global__ void k_body_ids(const BodyRec* __restrict__ rec, const int* __restrict__ idx, long long cap, int* __restrict__ body)
{
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < cap; r += (long long)gridDim.x * blockDim.x)
        body[r] = idx[r] >= 0 ? rec[idx[r]].body : 0x7fffffff;
}
// body[] ascending (stable in key), idx[] = record index
__global__ void k_reduce_bodies_sorted(const BodyRec* __restrict__ rec, const int* __restrict__ body, const int* __restrict__ idx,
                                       long long cap, int nbody, double* imp_rg)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbody) return;
    long long lo = 0, hi = cap;   // first position with body[pos] >= b
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (body[mid] < b) lo = mid + 1; else hi = mid;
    }
    double s[3] = {imp_rg[3 * b], imp_rg[3 * b + 1], imp_rg[3 * b + 2]};
    for (long long i = lo; i < cap && body[i] == b; ++i) {
        const BodyRec& r = rec[idx[i]];
        s[0] += r.v[0]; s[1] += r.v[1]; s[2] += r.v[2];
    }
    imp_rg[3 * b] = s[0]; imp_rg[3 * b + 1] = s[1]; imp_rg[3 * b + 2] = s[2];
}

// updateAverageVelocity :726-733: avgVel += collsnImpulse_RG / collsn_num_RG for every non-static
// point of a body that was hit this pass
__global__ void k_apply_bodies(int V, const uint8_t* __restrict__ vflags, const int* __restrict__ vbody,
                               const double* __restrict__ imp_rg, const int* __restrict__ cnt_rg, Vec4* av, uint8_t* has,
                               uint8_t* dirty)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= V) return;
    if (vflags[p] & 1) return;
    const int b = vbody[p];
    const int n = cnt_rg[b];
    if (n <= 0) return;
    has[p] = 1;
    dirty[p] = 1;
    Vec4 v = av[p];
    v.x += imp_rg[3 * b] / n;
    v.y += imp_rg[3 * b + 1] / n;
    v.z += imp_rg[3 * b + 2] / n;
    av[p] = v;
}

// start-of-apply state of the 'touched' flags: fixed and rigid-body points always count as touched
__global__ void k_reset_dirty(int V, const uint8_t* __restrict__ vflags, uint8_t* __restrict__ dirty)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < V) dirty[p] = (vflags[p] & 3) ? 1 : 0;
}

// computeAverageVelocity, dcollid.cpp:160-220
__global__ void k_avg_velocity(int V, const Vec4* __restrict__ xo, const Vec4* __restrict__ xn, Vec4* __restrict__ av,
                               double dt, unsigned long long* counters)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= V) return;
    Vec4 a = xo[p], b = xn[p], v;
    if (dt > CLSN_ROUND_EPS) {
        v.x = (b.x - a.x) / dt; v.y = (b.y - a.y) / dt; v.z = (b.z - a.z) / dt;
    } else {
        v.x = v.y = v.z = 0.0;
    }
    v.w = 0.0;
    if (isnan(v.x) || isinf(v.x) || isnan(v.y) || isinf(v.y) || isnan(v.z) || isinf(v.z))
        atomicAdd(&counters[CTR_ERROR], 1ull);
    av[p] = v;
}

// detectDomainBoundaryCollision (dcollid.cpp:116-158) -- once per unique point, see SURVEY a14
__global__ void k_boundary(int V, const uint8_t* __restrict__ vflags, const Vec4* __restrict__ xo, Vec4* av, uint8_t* has,
                           double dt, double mu, double lo0, double lo1, double lo2, double hi0, double hi1, double hi2)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= V) return;
    if (vflags[p] & 2) return;  // isMovableRigidBody(pt): skipped (:123)
    const Vec4 x = xo[p];
    Vec4 v4 = av[p];
    double v[3] = {v4.x, v4.y, v4.z};
    const double xs[3] = {x.x, x.y, x.z};
    const double L[3] = {lo0, lo1, lo2}, U[3] = {hi0, hi1, hi2};
    double dv = 0;
    bool hit = false;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const double cand = xs[j] + dt * v[j];
        if (cand <= L[j] || cand >= U[j]) {
            hit = true;
            dv = fabs(v[j]);
            v[j] = 0.0;
        }
    }
    const double preVt = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (preVt > CLSN_MACH_EPS) {
        const double f = stdmax(1.0 - mu * dv / preVt, 0.0);
#pragma unroll
        for (int j = 0; j < 3; ++j) v[j] *= f;
    }
    if (hit) has[p] = 1;
    v4.x = v[0]; v4.y = v[1]; v4.z = v[2];
    av[p] = v4;
}

// updateFinalPosition, dcollid.cpp:562-584
__global__ void k_final_position(int V, const Vec4* __restrict__ xo, const Vec4* __restrict__ av, Vec4* __restrict__ xn, double dt)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= V) return;
    const Vec4 x = xo[p], v = av[p];
    Vec4 r;
    r.x = x.x + v.x * dt; r.y = x.y + v.y * dt; r.z = x.z + v.z * dt; r.w = 0.0;
    xn[p] = r;
}

// packed xyz (3V doubles) <-> padded Vec4
__global__ void k_pack(int V, const double* __restrict__ src, Vec4* __restrict__ dst)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= V) return;
    dst[p] = Vec4{src[3 * (size_t)p], src[3 * (size_t)p + 1], src[3 * (size_t)p + 2], 0.0};
}
__global__ void k_unpack(int V, const Vec4* __restrict__ src, double* __restrict__ dst)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= V) return;
    const Vec4 v = src[p];
    dst[3 * (size_t)p] = v.x; dst[3 * (size_t)p + 1] = v.y; dst[3 * (size_t)p + 2] = v.z;
}

} // namespace clsn
