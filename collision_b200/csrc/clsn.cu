// collision_b200: context, pass orchestration and the C ABI (include/collision_b200.h).
//
// One context = one B200.  All state lives in HBM as padded per-vertex records (Vec4, one
// 32-byte sector per gather); the host side only gathers/scatters caller arrays.
// One detection pass (enqueue_detect), nothing read back by the host:
//   [tree order: k_scene_bounds -> k_morton -> cub radix sort -> k_gather_elems; kept across steps]
//   k_refit8<static|moving>  exact FP64 leaf boxes + quantised 64-byte nodes of the implicit 8-ary tree (lbvh.cuh)
//   k_traverse8              self query, exact FP64 leaf test -> element pairs (a < b)
//   k_cull<static|moving>    feature-level swept-box cull, CCD: trig-free classifier of the coplanarity cubic -> work list
//   proximity: k_contact<false>;  CCD: k_fast (plain-FP64 fast path) -> k_exact (correctly rounded solve) -> k_emit
//   (contact + impulse records; narrow.cuh, cubic.cuh, fastpath.cuh, crmath.cuh)
// and of the apply step (apply_impl): cub exclusive scan over per-point counts -> k_scatter -> k_reduce_points
// (canonical-order sums, applied to avgVel) -> body records -> [rigid-body kernels] (reduce.cuh, rigid.cuh).
// resolve_impl enqueues a whole step (proximity + up to five CCD passes, device-side gated) and synchronises once.
// Compiled with --fmad=false: every FP64 expression rounds exactly like the reference's.
#include "collision_b200.h"

#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <math.h>
#include <stdio.h>
#include <unistd.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <algorithm>
#include <thread>
#include <vector>

#include "narrow.cuh"
#include "lbvh.cuh"
#include "reduce.cuh"
#include "rigid.cuh"
#include "strain.cuh"
#include "dist.cuh"

using namespace clsn;

// ------------------------------------------------------------------ narrow-phase kernel
// Feature tables: local point slots (0..2 = element a, 3..5 = element b) of the 4 points of each test, in the
// reference's loop order.  Documentation (and the reference data of
// tests/test_host_cpu.py::test_feature_slot_arithmetic_matches_the_tables); the kernels use feature_slots() below.
/*
// tri-tri CCD (MovingTriToTri :286-323): k=0 tri a + vertex b_i, k=1 tri b + vertex a_i, then edges
c_feat_tt_moving[15][4] = {
    {0, 1, 2, 3}, {0, 1, 2, 4}, {0, 1, 2, 5}, {3, 4, 5, 0}, {3, 4, 5, 1}, {3, 4, 5, 2},
    {0, 1, 3, 4}, {0, 1, 4, 5}, {0, 1, 5, 3}, {1, 2, 3, 4}, {1, 2, 4, 5}, {1, 2, 5, 3},
    {2, 0, 3, 4}, {2, 0, 4, 5}, {2, 0, 5, 3}};
// tri-tri proximity (TriToTri :592-625): k=0 tri b (tri2) + vertex a_i, k=1 tri a + vertex b_i
c_feat_tt_static[15][4] = {
    {3, 4, 5, 0}, {3, 4, 5, 1}, {3, 4, 5, 2}, {0, 1, 2, 3}, {0, 1, 2, 4}, {0, 1, 2, 5},
    {0, 1, 3, 4}, {0, 1, 4, 5}, {0, 1, 5, 3}, {1, 2, 3, 4}, {1, 2, 4, 5}, {1, 2, 5, 3},
    {2, 0, 3, 4}, {2, 0, 4, 5}, {2, 0, 5, 3}};
// tri (a) - bond (b = slots 3,4) (TriToBond :508-530, MovingTriToBond :219-241)
c_feat_tb[5][4] = {{0, 1, 2, 3}, {0, 1, 2, 4}, {0, 1, 3, 4}, {1, 2, 3, 4}, {2, 0, 3, 4}};
*/

// The tables computed arithmetically: a lane-varying index into a __constant__ table is replayed once per
// distinct address in the warp, these few integer operations are not.  type 0 tri-tri, 1 tri-bond, 2 bond-bond.
// (tests/test_host_cpu.py checks this function against the tables above.)
template <bool MOVING>
__host__ __device__ __forceinline__ void feature_slots(int type, int f, int sl[4])
{
    if (type == 0) {
        if (f < 6) {
            const bool tri_a = MOVING ? f < 3 : f >= 3;   // whose triangle; the vertex comes from the other element
            const int v = f < 3 ? f : f - 3;
            const int t0 = tri_a ? 0 : 3;
            sl[0] = t0; sl[1] = t0 + 1; sl[2] = t0 + 2; sl[3] = (tri_a ? 3 : 0) + v;
        } else {
            const int e = f - 6, i = e / 3, j = e - 3 * i;
            sl[0] = i; sl[1] = i == 2 ? 0 : i + 1; sl[2] = 3 + j; sl[3] = j == 2 ? 3 : 4 + j;
        }
    } else if (type == 1) {
        if (f < 2) {
            sl[0] = 0; sl[1] = 1; sl[2] = 2; sl[3] = 3 + f;
        } else {
            const int i = f - 2;
            sl[0] = i; sl[1] = i == 2 ? 0 : i + 1; sl[2] = 3; sl[3] = 4;
        }
    } else {
        sl[0] = 0; sl[1] = 1; sl[2] = 3; sl[3] = 4;
    }
}

#define CULL_THREADS 128
#ifndef NARROW_GRID_MULT
#define NARROW_GRID_MULT 16  // grid-stride narrow-phase kernels: blocks per SM
#endif
#define FEAT_THREADS 128
#ifndef FEAT_MIN_BLOCKS
#define FEAT_MIN_BLOCKS 4
#endif
#ifndef CULL_PREFILTER
#define CULL_PREFILTER 0
#endif
#ifndef CULL_SAT
#define CULL_SAT 1   // FP32 separating-axis test along the feature normal in front of the classifier (cubic.cuh: sat_normal_far)
#endif
#define CULL_KEEP_CAP 128  // per-warp buffer of kept features between slot reservations
#define CULL_ROW 19  // doubles per staged pair row (18 used): odd stride -> conflict-free column access

// Work-list records.  They carry the four point ids of the feature test so that the consumer's gathers
// start right after one record load (no pairs -> elem -> points pointer chase).
struct FeatRec {  // 24 B
    unsigned entry;  // index into the pass's pair list (up to 2^32 - 1 pairs)
    int id[4];       // points as passed to PointToTri / EdgeToEdge
    unsigned edge;   // bit 0: 1 = edge-edge test, 0 = point-triangle; bits 1..4: feature index inside the pair
};
struct RootRec {  // 48 B: a feature whose coplanarity cubic has a usable root
    FeatRec f;
    double r0, r1, r2;  // sorted, invalid = -1 (isCoplanar's output)
};

// Conservative FP32 swept box of one point over [0, dt] (static: the point itself): rounded outward,
// so every position the narrow phase can evaluate lies inside.
struct FBox {
    float lo[3], hi[3];
};
__device__ __forceinline__ FBox fbox_union(const FBox& a, const FBox& b)
{
    FBox r;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        r.lo[d] = fminf(a.lo[d], b.lo[d]);
        r.hi[d] = fmaxf(a.hi[d], b.hi[d]);
    }
    return r;
}
// Can the two sub-features come within the contact distance at any t in [0, dt]?  A feature test fires
// only if, at some t in [0, dt], the point lies within h of the triangle's hull inflated by the
// barycentric slack eps (PointToTri :911-919), resp. the two segments come within h (EdgeToEdge :747).
// Every position used by the test lies in the swept boxes, so a gap larger than
// margin = 1.001 h + (3 eps + 2e-3) * extent along any axis means "the reference returns false" -- the
// 2e-3 * extent term covers the rounding of the test itself, including the barycentric coordinates of
// nearly degenerate triangles (DESIGN.md "exact culls").
// FP32 with directed rounding: the gap is rounded down, the margin up, so the cull stays conservative.
__device__ __forceinline__ bool boxes_far(const FBox& a, const FBox& b, float h2, float rel)
{
    float ext[3], emax = 0.f;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        ext[d] = fmaxf(__fsub_ru(a.hi[d], a.lo[d]), __fsub_ru(b.hi[d], b.lo[d]));
        emax = fmaxf(emax, ext[d]);
    }
    // The 2e-3 term covers the noise of the reference's barycentric coordinates on the slivers that still pass its
    // ABSOLUTE degeneracy gate |det| >= 1000 MACH_EPS (dcollid3d.cpp:797-799): relative error <= eps L^4 / det <= L^4 / 1000
    // for edge length L, i.e. it is a unit-scale figure.  For features larger than 1 the slack grows with L^4.
    if (emax > 1.f) {
        const float e2 = __fmul_ru(emax, emax);
        rel = __fmul_ru(rel, __fmul_ru(e2, e2));
    }
    bool far = false;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float m = __fmaf_ru(rel, ext[d], h2);
        far = far || (__fsub_rd(a.lo[d], b.hi[d]) > m) || (__fsub_rd(b.lo[d], a.hi[d]) > m);
    }
    return far;
}

// Stage 1: (i) one pair per thread: feature-level swept-box culling -> 15-bit mask; (ii) CCD only: the
// box survivors of the warp's 32 pairs are pooled and dealt out evenly to the 32 lanes (the six points
// of every pair are staged in shared memory, one row per pair), each lane computes the coefficients of
// the coplanarity cubic and runs the trig-free classifier coplanar_maybe().  Features that can still
// fire are appended to the work list (FeatRec: pair index, the four points, test kind | feature index).
#ifndef CULL_MIN_BLOCKS
#define CULL_MIN_BLOCKS 4
#endif
template <bool MOVING>
__global__ void __launch_bounds__(CULL_THREADS, CULL_MIN_BLOCKS)
k_cull(const int2* __restrict__ pairs, long long cap_pairs, const int4* __restrict__ elem, const Vec4* __restrict__ xo,
       const Vec4* __restrict__ av, NarrowParams P, FeatRec* __restrict__ feats, long long cap_feats,
       unsigned long long* counters, bool split_by_kind, long long pair_lo, long long pair_hi)
{
    __shared__ double s_x[CULL_THREADS][CULL_ROW];
    __shared__ double s_v[MOVING ? CULL_THREADS : 1][CULL_ROW];
    __shared__ unsigned s_mask[CULL_THREADS];
    __shared__ int s_id[CULL_THREADS][7];
    __shared__ unsigned short s_keep[CULL_THREADS / 32][CULL_KEEP_CAP];
    __shared__ unsigned short s_list[CULL_THREADS / 32][32 * 15];   // (pair row << 4) | feature of every box survivor of the batch
    const int tid = threadIdx.x, lane = tid & 31, wb = tid & ~31;
    long long n_pairs = (long long)counters[CTR_PAIRS];
    if (n_pairs > cap_pairs) n_pairs = cap_pairs;
    if (n_pairs > pair_hi) n_pairs = pair_hi;   // this launch handles the chunk [pair_lo, pair_hi) of the pair list
    // contact distance with 0.1 % head-room (covers the rounding of the distance itself), and the relative
    // slack: 3 eps for the barycentric tolerance + 2e-3 for the error of the barycentric coordinates of a
    // nearly degenerate triangle (cond <= L^4 / (1000 MACH_EPS), i.e. <= 4.5e-4 for L <= 1)
    const float h2 = __double2float_ru(1.001 * (MOVING ? P.eps : P.thickness));
    const float rel = __double2float_ru(3.0 * P.eps + 2e-3);
    unsigned long long n_box = 0;
    for (long long base = pair_lo + (long long)blockIdx.x * CULL_THREADS; base < n_pairs; base += (long long)gridDim.x * CULL_THREADS) {
        const long long pi = base + tid;
        unsigned mask = 0;
        if (pi < n_pairs) {
            const int2 pr = __ldg(pairs + pi);
            const int4 A = __ldg(elem + pr.x), B = __ldg(elem + pr.y);
            const int ids[6] = {A.x, A.y, A.z, B.x, B.y, B.z};
#pragma unroll
            for (int s = 0; s < 6; ++s) s_id[tid][s] = ids[s];
            FBox pb[6];
#pragma unroll
            for (int s = 0; s < 6; ++s) {
                if (ids[s] < 0) { pb[s] = pb[0]; continue; }
                const Vec4 x = ldg_vec4(xo + ids[s]);
                const double x0[3] = {x.x, x.y, x.z};
                s_x[tid][3 * s] = x.x; s_x[tid][3 * s + 1] = x.y; s_x[tid][3 * s + 2] = x.z;
                if (MOVING) {
                    const Vec4 v = ldg_vec4(av + ids[s]);
                    const double vv[3] = {v.x, v.y, v.z};
                    s_v[tid][3 * s] = v.x; s_v[tid][3 * s + 1] = v.y; s_v[tid][3 * s + 2] = v.z;
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        const double x1 = x0[d] + vv[d] * P.dt;
                        pb[s].lo[d] = __double2float_rd(fmin(x0[d], x1));
                        pb[s].hi[d] = __double2float_ru(fmax(x0[d], x1));
                    }
                } else {
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        pb[s].lo[d] = __double2float_rd(x0[d]);
                        pb[s].hi[d] = __double2float_ru(x0[d]);
                    }
                }
            }
            if (A.z >= 0 && B.z >= 0) {
                FBox ea[3], eb[3];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    ea[i] = fbox_union(pb[i], pb[(i + 1) % 3]);
                    eb[i] = fbox_union(pb[3 + i], pb[3 + (i + 1) % 3]);
                }
                const FBox ta = fbox_union(ea[0], pb[2]), tb = fbox_union(eb[0], pb[5]);
                // point-triangle features 0..5 (order differs between proximity and CCD, see the tables above)
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const bool a_tri_b_pt = !boxes_far(ta, pb[3 + i], h2, rel);  // triangle a, vertex b_i
                    const bool b_tri_a_pt = !boxes_far(tb, pb[i], h2, rel);      // triangle b, vertex a_i
                    if (MOVING) {
                        mask |= (a_tri_b_pt ? 1u : 0u) << i;
                        mask |= (b_tri_a_pt ? 1u : 0u) << (3 + i);
                    } else {
                        mask |= (b_tri_a_pt ? 1u : 0u) << i;
                        mask |= (a_tri_b_pt ? 1u : 0u) << (3 + i);
                    }
                }
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j)
                        if (!boxes_far(ea[i], eb[j], h2, rel)) mask |= 1u << (6 + 3 * i + j);
            } else if (A.z >= 0) {
                // triangle a, bond b: features 0,1 = vertex, 2..4 = tri edge x bond
                const FBox ta = fbox_union(fbox_union(pb[0], pb[1]), pb[2]), bb = fbox_union(pb[3], pb[4]);
                if (!boxes_far(ta, pb[3], h2, rel)) mask |= 1u;
                if (!boxes_far(ta, pb[4], h2, rel)) mask |= 2u;
#pragma unroll
                for (int i = 0; i < 3; ++i)
                    if (!boxes_far(fbox_union(pb[i], pb[(i + 1) % 3]), bb, h2, rel)) mask |= 1u << (2 + i);
                mask |= 1u << 16;  // pair type 1
            } else {
                mask = 1u | (2u << 16);  // bond-bond (type 2): the leaf boxes are the feature boxes
            }
        }
        // pool the warp's survivors: exclusive prefix of the per-pair counts
        const int cnt = __popc(mask & 0x7fffu);
        n_box += cnt;
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        s_mask[tid] = mask;
        {   // This lane's survivors, in feature order, at its place in the warp's list.  (Round 2 searched the owner row in
            // every pooled round instead -- binary search over the prefix sums + n-th set bit: 10 % of the kernel's
            // instructions behind a chain of six dependent shared-memory loads; measured: cull 4.74 -> 4.41 ms per step.)
            int pos = incl - cnt;
            for (unsigned m = mask & 0x7fffu; m; m &= m - 1u) s_list[tid >> 5][pos++] = (unsigned short)((lane << 4) | (__ffs(m) - 1));
        }
        __syncwarp();
        // Features that stay are buffered per warp as 10-bit codes (pair row, feature, list end) and written out
        // in batches: one slot reservation per list end and batch instead of one per round of 32 -- a
        // same-address atomic is serialised at the L2 (~0.85 cycles per op chip-wide) and sits in every round's
        // critical path otherwise.  Two list ends, so that the consumer's warps are homogeneous:
        // proximity -> point-triangle | edge-edge (k_contact); CCD -> point-triangle | edge-edge (k_fast)
        // or, for the staged pipeline, trig branch | other branches of the cubic (k_roots).
        int nk = 0;
        for (int k0 = 0; k0 < total; k0 += 32) {
            const int k = k0 + lane;
            bool keep = false;
            unsigned code = 0;
            if (k < total) {
                const unsigned lc = s_list[tid >> 5][k];
                const int o = (int)(lc >> 4), f = (int)(lc & 15u);   // owner row inside the warp's batch, feature index
                const unsigned m = s_mask[wb + o];
                const int type = (int)(m >> 16);
                const bool edge = type == 0 ? f >= 6 : (type == 1 ? f >= 2 : true);
                bool back = edge;
                keep = true;
                int sl[4];
                feature_slots<MOVING>(type, f, sl);
                Quad q;
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        q.xo[i][d] = s_x[wb + o][3 * sl[i] + d];
                        q.av[i][d] = MOVING ? s_v[MOVING ? wb + o : 0][3 * sl[i] + d] : 0.0;
                    }
#if CULL_SAT
                // Proximity pass: separated along the feature's own normal (cubic.cuh) -> the static test cannot fire.  Measured
                // on config 4: 18.1 M -> 7 M features for k_contact<false>, -0.12 ms.  Not used in front of the CCD classifier:
                // it rejects 43 % of the box survivors there, but the classifier rejects the same features for about twice the
                // price of the test, and the lanes of a warp wait for each other (measured: cull +1.1 ms, nothing saved after it).
                if (!MOVING && sat_normal_far(q, edge, false, P.dt, P.thickness, P.eps)) {
                    keep = false;
                } else
#endif
                if (MOVING) {
#if CULL_PREFILTER
                    // experimental (tools/build_variants.py "pf"): FP32 Bernstein test in front of the FP64 classifier
                    if (coplanar_prefilter32(q, P.dt)) {
                        keep = false;
                    } else
#endif
                    {
                        double ca, cb, cc, cd;
                        coplanar_coeffs(q, ca, cb, cc, cd);
                        const int kindc = coplanar_maybe(ca, cb, cc, cd, P.dt);
                        keep = kindc != 0;
                        if (!split_by_kind) back = kindc == 2;
                    }
                }
                code = ((unsigned)o << 5) | ((unsigned)f << 1) | (back ? 1u : 0u);
            }
            const unsigned kb = __ballot_sync(0xffffffffu, keep);
            if (keep) s_keep[tid >> 5][nk + __popc(kb & ((1u << lane) - 1u))] = (unsigned short)code;
            nk += __popc(kb);
            __syncwarp();
            if (nk > CULL_KEEP_CAP - 32 || k0 + 32 >= total) {
                int nb = 0;
                for (int e = lane; e < nk; e += 32) nb += s_keep[tid >> 5][e] & 1;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) nb += __shfl_xor_sync(0xffffffffu, nb, o);
                const int nf = nk - nb;
                unsigned long long base_f = 0, base_b = 0;
                if (lane == 0) {
                    if (nf) base_f = atomicAdd(&counters[CTR_FEATS], (unsigned long long)nf);
                    if (nb) base_b = atomicAdd(&counters[CTR_FEATS_EE], (unsigned long long)nb);
                }
                base_f = __shfl_sync(0xffffffffu, base_f, 0);
                base_b = __shfl_sync(0xffffffffu, base_b, 0);
                int run_f = 0, run_b = 0;
                for (int e0 = 0; e0 < nk; e0 += 32) {
                    const int e = e0 + lane;
                    const bool valid = e < nk;
                    const unsigned cde = valid ? s_keep[tid >> 5][e] : 0u;
                    const bool isb = valid && (cde & 1u);
                    const unsigned bb = __ballot_sync(0xffffffffu, isb), fb = __ballot_sync(0xffffffffu, valid && !isb);
                    if (valid) {
                        const int o = (int)(cde >> 5), f = (int)((cde >> 1) & 15u);
                        const int type = (int)(s_mask[wb + o] >> 16);
                        int sl[4];
                        feature_slots<MOVING>(type, f, sl);
                        const unsigned edge = type == 0 ? f >= 6 : (type == 1 ? f >= 2 : 1);
                        const unsigned lt = (1u << lane) - 1u;
                        const long long slot = isb ? (long long)base_b + run_b + __popc(bb & lt) : (long long)base_f + run_f + __popc(fb & lt);
                        if (slot < cap_feats) {
                            uint2* dst = reinterpret_cast<uint2*>(feats + (isb ? cap_feats - 1 - slot : slot));
                            dst[0] = make_uint2((unsigned)(base + wb + o), (unsigned)s_id[wb + o][sl[0]]);
                            dst[1] = make_uint2((unsigned)s_id[wb + o][sl[1]], (unsigned)s_id[wb + o][sl[2]]);
                            dst[2] = make_uint2((unsigned)s_id[wb + o][sl[3]], edge | ((unsigned)f << 1));
                        }
                    }
                    run_f += __popc(fb);
                    run_b += __popc(bb);
                }
                nk = 0;
                __syncwarp();
            }
        }
        __syncwarp();
    }
    for (int o = 16; o > 0; o >>= 1) n_box += __shfl_xor_sync(0xffffffffu, n_box, o);
    if (lane == 0 && n_box) atomicAdd(&counters[CTR_BOXSURV], n_box);
}

__device__ __forceinline__ FeatRec load_featrec(const FeatRec* p)
{
    const uint2* q = reinterpret_cast<const uint2*>(p);
    const uint2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    FeatRec r;
    r.entry = a.x; r.id[0] = (int)a.y; r.id[1] = (int)b.x; r.id[2] = (int)b.y; r.id[3] = (int)c.x; r.edge = c.y;
    return r;
}

// Stage 2 (CCD): correctly rounded solve of the coplanarity cubic for every feature the classifier
// let through.  Only positions and velocities are needed here; the kernel is nothing but the
// double-double math, which keeps its instruction footprint small.
#ifndef ROOTS_MIN_BLOCKS
#define ROOTS_MIN_BLOCKS 5
#endif
__global__ void __launch_bounds__(FEAT_THREADS, ROOTS_MIN_BLOCKS)
k_roots(const FeatRec* __restrict__ feats, long long cap_feats, const Vec4* __restrict__ xo, const Vec4* __restrict__ av, double dt,
        RootRec* __restrict__ out, long long cap_out, unsigned long long* counters)
{
    const long long n_pt = (long long)counters[CTR_FEATS], n_ee = (long long)counters[CTR_FEATS_EE];
    if (n_pt + n_ee > cap_feats) return;  // overflow: the host grows the list and repeats the pass
    const long long n = n_pt + n_ee;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const FeatRec fr = load_featrec(t < n_pt ? feats + t : feats + (cap_feats - 1 - (t - n_pt)));
        Quad q;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const Vec4 x = ldg_vec4(xo + fr.id[i]), v = ldg_vec4(av + fr.id[i]);
            q.xo[i][0] = x.x; q.xo[i][1] = x.y; q.xo[i][2] = x.z;
            q.av[i][0] = v.x; q.av[i][1] = v.y; q.av[i][2] = v.z;
        }
        double roots[3] = {-1, -1, -1};
        if (!is_coplanar<false>(q, dt, roots)) continue;
        const bool ee = (fr.edge & 1u) != 0;
        // reserve() aggregates the converged lanes onto ONE counter, so the two kinds reserve separately
        long long slot = 0;
        if (ee) slot = (long long)reserve1(&counters[CTR_ROOTS_EE]);
        else slot = (long long)reserve1(&counters[CTR_ROOTS]);
        if (slot < cap_out) {
            RootRec* dst = ee ? out + (cap_out - 1 - slot) : out + slot;
            uint2* o = reinterpret_cast<uint2*>(dst);
            o[0] = make_uint2(fr.entry, (unsigned)fr.id[0]);
            o[1] = make_uint2((unsigned)fr.id[1], (unsigned)fr.id[2]);
            o[2] = make_uint2((unsigned)fr.id[3], fr.edge);
            double* r = reinterpret_cast<double*>(dst) + 3;
            r[0] = roots[0]; r[1] = roots[1]; r[2] = roots[2];
        }
    }
}

#define CONTACT_QCAP 64

template <bool MOVING>
__device__ __forceinline__ void load_contact_item(const FeatRec* __restrict__ feats, const RootRec* __restrict__ rootrecs, long long at,
                                                  const Vec4* __restrict__ xo, const Vec4* __restrict__ av,
                                                  const uint8_t* __restrict__ vflags, FeatRec& fr, Quad& q, double& r0, double& r1, double& r2)
{
    r0 = r1 = r2 = 0;
    if (MOVING) {
        fr = load_featrec(&rootrecs[at].f);
        const double* r = reinterpret_cast<const double*>(rootrecs + at) + 3;
        r0 = __ldg(r); r1 = __ldg(r + 1); r2 = __ldg(r + 2);
    } else {
        fr = load_featrec(feats + at);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        q.id[i] = fr.id[i];
        const Vec4 x = ldg_vec4(xo + fr.id[i]), v = ldg_vec4(av + fr.id[i]);
        q.xo[i][0] = x.x; q.xo[i][1] = x.y; q.xo[i][2] = x.z;
        q.av[i][0] = v.x; q.av[i][1] = v.y; q.av[i][2] = v.z;
        q.flags[i] = __ldg(vflags + fr.id[i]);
        q.body[i] = 0;
    }
}

// Stage 3: static tests at the root times (CCD) or at x_old (proximity), then impulse records.
// Two phases per warp: (1) every lane finds the first time at which its feature fires (decision only) and
// pushes (list index, time) to a per-warp shared-memory queue; (2) whenever 32 hits are queued the whole
// warp turns them into contact + impulse records -- the expensive response code runs with all lanes
// busy instead of the ~15 % that hit.
template <bool MOVING>
__global__ void __launch_bounds__(FEAT_THREADS, FEAT_MIN_BLOCKS)
k_contact(const FeatRec* __restrict__ feats, const RootRec* __restrict__ rootrecs, long long cap_in, const int2* __restrict__ pairs,
          const Vec4* __restrict__ xo, const Vec4* __restrict__ av, const uint8_t* __restrict__ vflags,
          const int* __restrict__ vbody, NarrowParams P, Emit E, unsigned* __restrict__ pair_hit)
{
    __shared__ long long s_at[FEAT_THREADS / 32][CONTACT_QCAP];
    __shared__ double s_time[FEAT_THREADS / 32][CONTACT_QCAP];
    __shared__ int s_n[FEAT_THREADS / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long n_pt = (long long)E.counters[MOVING ? CTR_ROOTS : CTR_FEATS];
    const long long n_ee = (long long)E.counters[MOVING ? CTR_ROOTS_EE : CTR_FEATS_EE];
    if (n_pt + n_ee > cap_in) return;  // overflow: the host grows the list and repeats the pass
    const long long n = n_pt + n_ee;
    const double h = MOVING ? P.eps : P.thickness;
    if (lane == 0) s_n[w] = 0;
    __syncwarp();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t0 = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31); ; t0 += stride) {
        const long long t = t0 + lane;
        if (t0 < n && t < n) {
            const long long at = t < n_pt ? t : cap_in - 1 - (t - n_pt);
            FeatRec fr;
            Quad q;
            double r0, r1, r2;
            load_contact_item<MOVING>(feats, rootrecs, at, xo, av, vflags, fr, q, r0, r1, r2);
            const double th = feature_first_hit<MOVING>(P, E, q, (fr.edge & 1u) != 0, h, r0, r1, r2);
            if (th >= 0) {
                const int pos = atomicAdd(&s_n[w], 1);
                s_at[w][pos] = at;
                s_time[w][pos] = th;
            }
        }
        __syncwarp();
        const bool done = t0 + stride >= n;  // no further round for this warp
        int nq = s_n[w];
        while (nq >= 32 || (done && nq > 0)) {
            const int take = nq < 32 ? nq : 32;
            const int base = nq - take;
            if (lane < take) {
                const long long at = s_at[w][base + lane];
                const double th = s_time[w][base + lane];
                FeatRec fr;
                Quad q;
                double r0, r1, r2;
                load_contact_item<MOVING>(feats, rootrecs, at, xo, av, vflags, fr, q, r0, r1, r2);
                if ((q.flags[0] & 3) && (q.flags[1] & 3) && (q.flags[2] & 3) && (q.flags[3] & 3)) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) q.body[i] = __ldg(vbody + q.id[i]);
                }
                const unsigned pi = fr.entry;
                const int f = (int)(fr.edge >> 1);
                const int2 pr = __ldg(pairs + pi);
                const unsigned long long key = ((unsigned long long)(unsigned)pr.x << 34) | ((unsigned long long)(unsigned)pr.y << 4) |
                                               (unsigned long long)f;
                feature_emit<MOVING>(P, E, q, key, (fr.edge & 1u) != 0, h, th);
                atomicOr(pair_hit + (pi >> 5), 1u << (pi & 31));
            }
            nq = base;
            __syncwarp();
        }
        __syncwarp();  // every lane has read s_n[w] before lane 0 rewrites it
        if (lane == 0) s_n[w] = nq;
        __syncwarp();
        if (done) break;
    }
}

// ------------------------------------------------------------------ CCD features with the fast path (pipeline 1)
// fastpath.cuh settles three features out of four in plain FP64: no valid root (nothing to do), or "misses at
// every root" -- then the outcome is the reference's own static test at t = dt, run right there.  Only the
// features that may fire AT a root need the correctly rounded solve.  Two lean kernels, so that neither carries
// the other's code or registers: k_fast (fast path + test at dt; the undecided features go to a list) and
// k_exact (correctly rounded solve + the reference's walk over the roots for that list).  Both append to the hit
// list (feature + first hit time), point-triangle entries from the front, edge-edge entries from the back;
// k_emit turns it into contact and impulse records.
struct HitRec {  // 32 B
    FeatRec f;
    double t;
};

__device__ __forceinline__ void load_quad(const FeatRec& fr, const Vec4* __restrict__ xo, const Vec4* __restrict__ av, Quad& q)
{
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        q.id[i] = fr.id[i];
        const Vec4 x = ldg_vec4(xo + fr.id[i]), v = ldg_vec4(av + fr.id[i]);
        q.xo[i][0] = x.x; q.xo[i][1] = x.y; q.xo[i][2] = x.z;
        q.av[i][0] = v.x; q.av[i][1] = v.y; q.av[i][2] = v.z;
        q.flags[i] = 0;
        q.body[i] = 0;
    }
}

__device__ __forceinline__ void store_featrec(FeatRec* dst, const FeatRec& fr)
{
    uint2* o = reinterpret_cast<uint2*>(dst);
    o[0] = make_uint2(fr.entry, (unsigned)fr.id[0]);
    o[1] = make_uint2((unsigned)fr.id[1], (unsigned)fr.id[2]);
    o[2] = make_uint2((unsigned)fr.id[3], fr.edge);
}

__device__ __forceinline__ void push_hit(HitRec* __restrict__ hits, long long cap_hits, unsigned long long* counters, const FeatRec& fr,
                                         double t)
{
    const bool ee = (fr.edge & 1u) != 0;
    long long slot;
    if (ee) slot = (long long)reserve1(&counters[CTR_HITS_EE]);
    else slot = (long long)reserve1(&counters[CTR_HITS]);
    if (slot < cap_hits) {
        HitRec* dst = ee ? hits + (cap_hits - 1 - slot) : hits + slot;
        store_featrec(&dst->f, fr);
        reinterpret_cast<double*>(dst)[3] = t;
    }
}

// correctly rounded solve + the reference's walk over the roots (MovingPointToTri / MovingEdgeToEdge,
// dcollid3d.cpp:327-369): first time at which the static test fires, or -1
template <bool EDGE>
__device__ __forceinline__ double exact_first_hit(const NarrowParams& P, const Emit& E, const Quad& q, bool& coplanar)
{
    double roots[3] = {-1, -1, -1};
    coplanar = is_coplanar<false>(q, P.dt, roots);
    if (!coplanar) return -1.0;
    double X[4][3];
#pragma unroll 1
    for (int i = 0; i < 4; ++i) {
        const double t = i < 3 ? roots[i] : P.dt;
        if (t < 0) continue;
        positions_at<true>(q, t, X);
        const bool hit = EDGE ? edge_to_edge<false>(P, E, q, 0ull, X, P.eps, t) : point_to_tri<false>(P, E, q, 0ull, X, P.eps, t);
        if (hit) return t;
    }
    return -1.0;
}

#ifndef FAST_MIN_BLOCKS
#define FAST_MIN_BLOCKS 4
#endif
#ifndef EXACT_MIN_BLOCKS
#define EXACT_MIN_BLOCKS 4
#endif
#define FAST_QCAP 64

// EDGE = false: the point-triangle entries (front of the work list); EDGE = true: the edge-edge entries (back).
// The two outputs of a round (hits at dt, undecided features) are buffered per warp as work-list indices and
// written out 32 at a time: one slot reservation per 32 entries instead of one or two per round (a same-address
// atomic is serialised at the L2 and its round trip would sit in every round's critical path).
template <bool EDGE>
__global__ void __launch_bounds__(FEAT_THREADS, FAST_MIN_BLOCKS)
k_fast(const FeatRec* __restrict__ feats, long long cap_feats, const Vec4* __restrict__ xo, const Vec4* __restrict__ av,
       NarrowParams P, Emit E, HitRec* __restrict__ hits, long long cap_hits, FeatRec* __restrict__ unc, long long cap_unc)
{
    const int lane = threadIdx.x & 31;
    const long long n_pt = (long long)E.counters[CTR_FEATS], n_ee = (long long)E.counters[CTR_FEATS_EE];
    if (n_pt + n_ee > cap_feats) return;  // overflow: the host grows the list and repeats the pass
    const long long n = EDGE ? n_ee : n_pt;
    unsigned long long n_cop = 0;
    __shared__ long long s_hit[FEAT_THREADS / 32][FAST_QCAP], s_unc[FEAT_THREADS / 32][FAST_QCAP];
    const int w = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    int nh = 0, nu = 0;  // warp-uniform fill of the two buffers
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t0 = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31); t0 < n; t0 += stride) {
        const long long t = t0 + lane;
        bool hit = false, undecided = false;
        long long at = 0;
        if (t < n) {
            at = EDGE ? cap_feats - 1 - t : t;
            const FeatRec fr = load_featrec(feats + at);
            Quad q;
            load_quad(fr, xo, av, q);
            const int st = feature_fast(q, EDGE, P.dt, P.eps, P.eps);
            if (st == FAST_DT_ONLY) {
                ++n_cop;
                double X[4][3];
                positions_at<true>(q, P.dt, X);
                hit = EDGE ? edge_to_edge<false>(P, E, q, 0ull, X, P.eps, P.dt) : point_to_tri<false>(P, E, q, 0ull, X, P.eps, P.dt);
            } else {
                undecided = st == FAST_UNCERTAIN;
            }
        }
        const unsigned hb = __ballot_sync(0xffffffffu, hit), ub = __ballot_sync(0xffffffffu, undecided);
        if (hit) s_hit[w][nh + __popc(hb & lt)] = at;
        if (undecided) s_unc[w][nu + __popc(ub & lt)] = at;
        nh += __popc(hb);
        nu += __popc(ub);
        __syncwarp();
        const bool last = t0 + stride >= n;
        while (nh >= 32 || (last && nh > 0)) {
            const int take = nh < 32 ? nh : 32;
            unsigned long long s0 = 0;
            if (lane == 0) s0 = atomicAdd(&E.counters[EDGE ? CTR_HITS_EE : CTR_HITS], (unsigned long long)take);
            s0 = __shfl_sync(0xffffffffu, s0, 0);
            const long long slot = (long long)s0 + lane;
            if (lane < take && slot < cap_hits) {
                const FeatRec fr = load_featrec(feats + s_hit[w][nh - take + lane]);
                HitRec* dst = EDGE ? hits + (cap_hits - 1 - slot) : hits + slot;
                store_featrec(&dst->f, fr);
                reinterpret_cast<double*>(dst)[3] = P.dt;
            }
            nh -= take;
            __syncwarp();
        }
        while (nu >= 32 || (last && nu > 0)) {
            const int take = nu < 32 ? nu : 32;
            unsigned long long s0 = 0;
            if (lane == 0) s0 = atomicAdd(&E.counters[EDGE ? CTR_UNC_EE : CTR_UNC], (unsigned long long)take);
            s0 = __shfl_sync(0xffffffffu, s0, 0);
            const long long slot = (long long)s0 + lane;
            if (lane < take && slot < cap_unc) {
                const FeatRec fr = load_featrec(feats + s_unc[w][nu - take + lane]);
                store_featrec(EDGE ? unc + (cap_unc - 1 - slot) : unc + slot, fr);
            }
            nu -= take;
            __syncwarp();
        }
    }
    for (int o = 16; o > 0; o >>= 1) n_cop += __shfl_xor_sync(0xffffffffu, n_cop, o);
    if (lane == 0 && n_cop) atomicAdd(&E.counters[CTR_ROOTS], n_cop);   // features with isCoplanar == true (stats)
}

template <bool EDGE>
__global__ void __launch_bounds__(FEAT_THREADS, EXACT_MIN_BLOCKS)
k_exact(const FeatRec* __restrict__ unc, long long cap_unc, const Vec4* __restrict__ xo, const Vec4* __restrict__ av, NarrowParams P,
        Emit E, HitRec* __restrict__ hits, long long cap_hits)
{
    const int lane = threadIdx.x & 31;
    const long long n_pt = (long long)E.counters[CTR_UNC], n_ee = (long long)E.counters[CTR_UNC_EE];
    if (n_pt + n_ee > cap_unc) return;  // overflow: the host grows the list and repeats the pass
    const long long n = EDGE ? n_ee : n_pt;
    unsigned long long n_cop = 0, n_exact = 0;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const long long at = EDGE ? cap_unc - 1 - t : t;
        const FeatRec fr = load_featrec(unc + at);
        Quad q;
        load_quad(fr, xo, av, q);
        ++n_exact;
        bool cop;
        const double th = exact_first_hit<EDGE>(P, E, q, cop);
        if (cop) ++n_cop;
        if (th >= 0) push_hit(hits, cap_hits, E.counters, fr, th);
    }
    for (int o = 16; o > 0; o >>= 1) {
        n_cop += __shfl_xor_sync(0xffffffffu, n_cop, o);
        n_exact += __shfl_xor_sync(0xffffffffu, n_exact, o);
    }
    if (lane == 0) {
        if (n_cop) atomicAdd(&E.counters[CTR_ROOTS], n_cop);
        if (n_exact) atomicAdd(&E.counters[CTR_EXACT], n_exact);
    }
}

// Pipeline 2 (experimental): per-point record counts of the hit list, so that k_emit<EDGE, true> can write every record
// straight into its point's segment.  Mirrors the emission rules of PointToTriImpulse / EdgeToEdgeImpulse: one point
// record per non-static point (dcollid3d.cpp:1053-1079, 1235-1284), none when all four points are rigid (body records).
template <bool EDGE>
__global__ void k_count_hits(const HitRec* __restrict__ hits, long long cap_hits, const uint8_t* __restrict__ vflags, int* cnt,
                             unsigned long long* counters)
{
    const long long n_pt = (long long)counters[CTR_HITS], n_ee = (long long)counters[CTR_HITS_EE];
    if (n_pt + n_ee > cap_hits) return;
    const long long n = EDGE ? n_ee : n_pt;
    unsigned long long total = 0;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const long long at = EDGE ? cap_hits - 1 - t : t;
        const FeatRec fr = load_featrec(&hits[at].f);
        int fl[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) fl[i] = __ldg(vflags + fr.id[i]);
        if ((fl[0] & 3) && (fl[1] & 3) && (fl[2] & 3) && (fl[3] & 3)) continue;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (!(fl[i] & 1)) {
                atomicAdd(cnt + fr.id[i], 1);
                ++total;
            }
    }
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    if ((threadIdx.x & 31) == 0 && total) atomicAdd(&counters[CTR_PREC], total);
}

// contact + impulse records of the hit list (the second half of k_contact, on a dense list).
// SEG: records go to per-point segments (offsets from the scan of k_count_hits' counts) instead of the append list.
template <bool EDGE, bool SEG = false>
__global__ void __launch_bounds__(FEAT_THREADS, FEAT_MIN_BLOCKS)
k_emit(const HitRec* __restrict__ hits, long long cap_hits, const int2* __restrict__ pairs, const Vec4* __restrict__ xo,
       const Vec4* __restrict__ av, const uint8_t* __restrict__ vflags, const int* __restrict__ vbody, NarrowParams P, Emit E,
       unsigned* __restrict__ pair_hit, SegOut S = SegOut{nullptr, nullptr})
{
    const long long n_pt = (long long)E.counters[CTR_HITS], n_ee = (long long)E.counters[CTR_HITS_EE];
    if (n_pt + n_ee > cap_hits) return;  // overflow: the host grows the list and repeats the pass
    const long long n = EDGE ? n_ee : n_pt;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const long long at = EDGE ? cap_hits - 1 - t : t;
        const FeatRec fr = load_featrec(&hits[at].f);
        const double th = __ldg(reinterpret_cast<const double*>(hits + at) + 3);
        Quad q;
        load_quad(fr, xo, av, q);
#pragma unroll
        for (int i = 0; i < 4; ++i) q.flags[i] = __ldg(vflags + fr.id[i]);
        if ((q.flags[0] & 3) && (q.flags[1] & 3) && (q.flags[2] & 3) && (q.flags[3] & 3)) {
#pragma unroll
            for (int i = 0; i < 4; ++i) q.body[i] = __ldg(vbody + q.id[i]);
        }
        const unsigned pi = fr.entry;
        const int f = (int)(fr.edge >> 1);
        const int2 pr = __ldg(pairs + pi);
        const unsigned long long key = ((unsigned long long)(unsigned)pr.x << 34) | ((unsigned long long)(unsigned)pr.y << 4) |
                                       (unsigned long long)f;
        feature_emit<true, SEG>(P, E, q, key, EDGE, P.eps, th, S);
        atomicOr(pair_hit + (pi >> 5), 1u << (pi & 31));
    }
}

// number of pairs for which isProximity / isCollision returned true (the tree's `count`, AABB.cpp:296-297)
__global__ void k_count_true(const unsigned* __restrict__ pair_hit, long long n_words, unsigned long long* counters)
{
    unsigned long long c = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += (long long)gridDim.x * blockDim.x)
        c += __popc(pair_hit[i]);
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&counters[CTR_TRUE], c);
}

// End of a chunk of the pair list: remember the sums of the six list cursors and the largest chunk, restart the cursors
// (last chunk: leave the sums in them, which is what the statistics report).
__global__ void k_fold_chunk(unsigned long long* ctr, int last)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int cur[6] = {CTR_FEATS, CTR_FEATS_EE, CTR_UNC, CTR_UNC_EE, CTR_HITS, CTR_HITS_EE};
    const unsigned long long f = ctr[CTR_FEATS] + ctr[CTR_FEATS_EE], u = ctr[CTR_UNC] + ctr[CTR_UNC_EE], h = ctr[CTR_HITS] + ctr[CTR_HITS_EE];
    if (f > ctr[CTR_MAX_FEATS]) ctr[CTR_MAX_FEATS] = f;
    if (u > ctr[CTR_MAX_UNC]) ctr[CTR_MAX_UNC] = u;
    if (h > ctr[CTR_MAX_HITS]) ctr[CTR_MAX_HITS] = h;
    for (int i = 0; i < 6; ++i) {
        ctr[CTR_TOT + i] += ctr[cur[i]];
        ctr[cur[i]] = last ? ctr[CTR_TOT + i] : 0ull;
    }
}

// ------------------------------------------------------------------ context
template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaError_t reserve(size_t want)
    {
        if (want <= n) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
        cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
        if (e == cudaSuccess) n = want;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
};

enum { PH_AVG = 0, PH_BUILD, PH_REFIT, PH_TRAVERSE, PH_CULL, PH_ROOTS, PH_CONTACT, PH_REDUCE, PH_FINAL, PH_OTHER, PH_COUNT };

struct clsn_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;  // created by clsn_create; `stream` may be replaced by clsn_set_stream
    std::string err;
    clsn_params prm;
    int V = 0, T = 0, B = 0, N = 0, nbody = 0;
    bool has_movable = false;
    // topology
    DevBuf<int4> elem;
    DevBuf<uint8_t> vflags;
    DevBuf<int> vbody;
    DevBuf<double> body_mass;
    RigidTopo rigid;
    // updateFinalForRG (dcollid.cpp:626-675): movable points of every body in hseList walk order (first occurrences)
    std::vector<int> rg_offs, rg_pts;      // CSR over bodies
    DevBuf<int> d_rg_pts;
    DevBuf<double> d_rg_state;             // gathered per listed point: avgVel xyz + has_collsn (4 doubles)
    std::vector<double> mrg_com;           // CollisionSolver::mrg_com (collid.h:174): lives across steps
    std::vector<uint8_t> mrg_valid;
    // impact zones (host fail-safe): host topology + union-find, device lists
    std::vector<int> h_tri, h_tri_surf, h_bond;
    std::vector<uint8_t> h_vflags;
    HostUF zone_uf;
    bool zone_uf_ready = false, zone_lists_stale = true;
    RigidTopo zone_lists;
    bool impact_zones = false;
    int zone_max_iter = 0;
    // strain limiting
    StrainTopo strain;
    bool strain_limiting = false, strain_pending = false;
    // state
    DevBuf<Vec4> xo, xn, av;
    DevBuf<uint8_t> has, dirty;
    bool dirty_valid = false;   // dirty[] describes exactly the change between the last two CCD passes
    bool exact_stats = false;   // count every overlapping pair in `candidates` even where the query could be pruned
    int last_detect_mode = -1;
    DevBuf<double> imp_rg, imp_rg_keep;   // keep: start-of-step copy (a step that has to be repeated restores it)
    DevBuf<Vec4> xf;                      // final positions (x_new stays intact for a repeated step)
    bool final_valid = false;
    bool phase_timing = false;            // record CUDA-event marks between kernel groups inside clsn_resolve (clsn_set_phase_timing)
    DevBuf<int> cnt_rg;
    DevBuf<double> stage;  // 3V doubles, host<->device staging of packed arrays
    // bvh
    DevBuf<unsigned> code, code_sorted;
    DevBuf<int> idx, leaf_elem;
    DevBuf<int4> selem;             // elem[] in Morton (leaf) order
    DevBuf<Node8> nodes;
    DevBuf<float> tree_scratch;     // upper-level ping-pong boxes of k_refit8 + root box (last 8 floats)
    DevBuf<uint8_t> tree_scratch_t;
    DevBuf<unsigned> tree_ticket;
    Tree8 tree;
    DevBuf<double> lbox;
    DevBuf<unsigned long long> bounds;
    DevBuf<unsigned char> cub_tmp;
    bool tree_built = false;
    bool keep_tree = true;     // reuse the leaf order across steps (CLSN_KEEP_TREE=0: rebuild every step)
    int tree_age = 0;          // steps refitted since the last build
    double tree_volume = 0.0;  // volume of the root box at the first refit after the build
    // pass buffers
    DevBuf<int2> pairs, dbg_cand;
    DevBuf<FeatRec> feats;
    DevBuf<unsigned> pair_hit;
    DevBuf<RootRec> rootrecs;
    DevBuf<HitRec> hits;
    DevBuf<FeatRec> unc;        // pipeline 1: features the fast path could not settle
    bool seg_records = false;   // pipeline 2: the pending records already sit in per-point segments (offs valid)
    long long pair_chunk = 1ll << 24;   // pairs per narrow-phase chunk (CLSN_PAIR_CHUNK; see enqueue_detect)
    int pipeline = 1;           // 1 = fast path first (k_fast + k_exact + k_emit), 0 = staged (k_roots + k_contact),
                                // 2 = 1 + records emitted into per-point segments (experimental)
    DevBuf<PointRec> prec, prec_sorted;
    DevBuf<ulonglong2> phdr;    // (key, point) headers of prec[], dense
    DevBuf<BodyRec> brec;
    DevBuf<Contact> contacts;
    DevBuf<int> cnt, offs, fill, perm, perm_sorted;
    DevBuf<unsigned long long> skey, skey_sorted;
    DevBuf<unsigned char> cub_tmp2;   // cub::DeviceSegmentedSort (movable-body scenes)
    DevBuf<unsigned long long> bkey;  // body records in (body, key) order: reduce_bodies()
    DevBuf<int> bidx, bbody;
    DevBuf<unsigned long long> counters;
    DevBuf<double> acc_imp, acc_fric;
    unsigned long long* h_counters = nullptr;  // pinned: PASS_SLOTS blocks of CTR_STRIDE counters
    unsigned long long* ctr = nullptr;         // counter block of the pass being enqueued / applied
    const unsigned long long* pending_gate = nullptr;  // gate of the pass whose records are pending (see enqueue_detect)
    long long last_true = 0;
    bool cnt_rg_preset = false;   // clsn_set_body_accumulators wrote collsn_num_RG (parity tests)
    double* h_pin = nullptr;                   // pinned staging for host arrays (3V doubles x 2)
    size_t h_pin_n = 0;
    // imported record set (multi-GPU)
    const PointRec* imp_prec = nullptr;
    const BodyRec* imp_brec = nullptr;
    long long imp_nprec = -1, imp_nbrec = 0;
    bool records_pending = false;
    long long last_nprec = 0, last_nbrec = 0;
    int rank = 0, nranks = 1;
    // in-library multi-GPU (dist.cuh): owner-computes exchange through peer memory + NCCL
    struct Dist {
        bool on = false;
        int per_rank = 0;
        NcclApi api;
        ncclComm_t comm = nullptr;
        long long cap_region = 0;
        DevBuf<PointRec> recv;                   // nranks regions of cap_region records, written by the peers
        DevBuf<unsigned long long> hdr;          // [nranks] records per region of the current pass (written by the peers)
        void* peer_recv[CLSN_MAX_RANKS];         // peers' recv / hdr buffers mapped into this process (own rank: local pointer)
        void* peer_hdr[CLSN_MAX_RANKS];
        bool peer_ipc = false;
        bool peer_local[CLSN_MAX_RANKS];         // peer context in this process: no IPC mapping to close
        bool direct = false;
        DevBuf<PointRec*> d_peer_region;         // [r] -> this rank's region inside rank r's receive buffer (peer mapping)
        DevBuf<PointRec> stage;                  // records staged per owner before k_push_regions
        DevBuf<PointRec*> d_stage_region;        // [r] -> where the emitting kernels put records owned by r
        DevBuf<unsigned long long*> d_peer_hdr;
        DevBuf<unsigned long long> send_cnt;     // [nranks] cursors of the current pass
        DevBuf<unsigned long long> gsum;         // PASS_SLOTS all-reduced counter blocks
        DevBuf<unsigned long long> maxblk;       // [4] local maxima of the pass; all-gathered into ...
        DevBuf<unsigned long long> allmax;       // ... PASS_SLOTS x nranks x 4
        long long cap_brec_x = 65536;            // body records exchanged per rank (fixed-size all-gather)
        DevBuf<BodyRec> brec_all, brec_dense;
        DevBuf<unsigned long long> brec_n;       // dense body-record count + a zero word
        unsigned long long* h_gsum = nullptr;    // pinned: gsum blocks + allmax
    } dist;
    bool dbg_candidates = false, dbg_contacts = false;
    long long n_dbg_cand = 0, n_contacts = 0;
    cudaEvent_t ev[2 * PH_COUNT + 2];
    // phase timing: events recorded on the stream between kernel groups, resolved after the step
    std::vector<cudaEvent_t> marks;
    std::vector<int> mark_phase;
    size_t n_marks = 0;
    bool timing = false;
    long long launches = 0;        // kernels launched by this library (incl. CUB's) since the last reset
    long long kernel_count[16] = {0};
    cudaEvent_t bracket[2];
    int sm_count = 148;
};

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t _e = (call);                                                              \
        if (_e != cudaSuccess) {                                                              \
            c->err = std::string(#call) + ": " + cudaGetErrorString(_e);                      \
            return _e == cudaErrorMemoryAllocation ? CLSN_E_NOMEM : CLSN_E_CUDA;              \
        }                                                                                     \
    } while (0)

// Counter blocks.  Every detection pass of a step has its own block of CTR_STRIDE counters, so that a whole step can be
// enqueued without reading anything back: slot 0 = proximity, 1..5 = CCD passes, PASS_SLOT_SOLO = a pass driven through
// the per-phase ABI or the impact-zone loop, PASS_SLOT_STEP = step-wide error flags (avgVel, boundary ...).
#define CTR_STRIDE 32
#define PASS_SLOTS 8
#define PASS_SLOT_SOLO 6
#define PASS_SLOT_STEP 7
#define CTR_LEGACY 256   // counters.p[0 .. 255]: import / owner-bucket scratch (multi-GPU exchange)
static_assert(CTR_COUNT <= CTR_STRIDE, "counter block too small");
static inline unsigned long long* pass_block(clsn_ctx* c, int slot) { return c->counters.p + CTR_LEGACY + CTR_STRIDE * slot; }
static inline unsigned long long* host_block(clsn_ctx* c, int slot) { return c->h_counters + CTR_STRIDE * slot; }

// record "everything enqueued since the previous mark belongs to `phase`"
static void mark(clsn_ctx* c, int phase)
{
    if (!c->timing) return;
    if (c->n_marks == c->marks.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        c->marks.push_back(e);
        c->mark_phase.push_back(0);
    }
    cudaEventRecord(c->marks[c->n_marks], c->stream);
    c->mark_phase[c->n_marks] = phase;
    ++c->n_marks;
}

static inline int nblk(long long n, int t) { return (int)((n + t - 1) / t > 0 ? (n + t - 1) / t : 1); }

static int fail(clsn_ctx* c, int code, const char* msg)
{
    c->err = msg;
    return code;
}

extern "C" int clsn_create(clsn_ctx** out, int device)
{
    if (!out) return CLSN_E_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return CLSN_E_CUDA;
    clsn_ctx* c = new clsn_ctx();
    c->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete c;
        return CLSN_E_CUDA;
    }
    c->stream = c->own_stream;
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
#ifdef CULL_SMEM_CARVEOUT
    // tuning experiment (tools/build_variants.py): k_cull keeps 44.5 KB of shared memory per block, so a fifth
    // resident block (CULL_MIN_BLOCKS=5) needs the full shared-memory carve-out
    cudaFuncSetAttribute(k_cull<true>, cudaFuncAttributePreferredSharedMemoryCarveout, CULL_SMEM_CARVEOUT);
    cudaFuncSetAttribute(k_cull<false>, cudaFuncAttributePreferredSharedMemoryCarveout, CULL_SMEM_CARVEOUT);
#endif
    for (auto& e : c->ev) cudaEventCreate(&e);
    cudaEventCreate(&c->bracket[0]);
    cudaEventCreate(&c->bracket[1]);
    cudaMallocHost((void**)&c->h_counters, (PASS_SLOTS * CTR_STRIDE + 8) * sizeof(unsigned long long));   // + root box of the last refit
    c->counters.reserve(CTR_LEGACY + PASS_SLOTS * CTR_STRIDE);
    cudaMemset(c->counters.p, 0, (CTR_LEGACY + PASS_SLOTS * CTR_STRIDE) * sizeof(unsigned long long));
    c->ctr = pass_block(c, PASS_SLOT_STEP);
    c->bounds.reserve(8);
    clsn_params p;
    p.eps = 1e-6; p.thickness = 1e-4; p.dt = 1e-3; p.k = 1000; p.m = 0.01; p.lambda = 0.02; p.cr = 0.0;
    for (int i = 0; i < 3; ++i) { p.lo[i] = -1e30; p.hi[i] = 1e30; }
    c->prm = p;
    if (const char* e = getenv("CLSN_PAIR_CHUNK")) c->pair_chunk = std::max<long long>(1024, atoll(e));
    if (const char* e = getenv("CLSN_KEEP_TREE")) c->keep_tree = atoi(e) != 0;
    if (const char* e = getenv("CLSN_PHASE_TIMING")) c->phase_timing = atoi(e) != 0;
    if (getenv("CLSN_TRACE")) c->phase_timing = true;   // the per-mark timeline needs the marks
    if (const char* e = getenv("CLSN_PIPELINE")) {
        const int v = atoi(e);
        if (v >= 0 && v <= 2) c->pipeline = v;
    }
    *out = c;
    return CLSN_OK;
}

extern "C" void clsn_destroy(clsn_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->elem.release(); c->vflags.release(); c->vbody.release(); c->body_mass.release();
    c->xo.release(); c->xn.release(); c->av.release(); c->has.release(); c->dirty.release(); c->imp_rg.release(); c->imp_rg_keep.release(); c->xf.release(); c->cnt_rg.release();
    c->stage.release(); c->code.release(); c->code_sorted.release(); c->idx.release(); c->leaf_elem.release();
    c->selem.release(); c->nodes.release(); c->tree_scratch.release(); c->tree_scratch_t.release();
    c->tree_ticket.release(); c->lbox.release(); c->bounds.release();
    c->cub_tmp.release(); c->pairs.release(); c->dbg_cand.release(); c->feats.release(); c->pair_hit.release(); c->rootrecs.release(); c->hits.release(); c->unc.release(); c->prec.release(); c->prec_sorted.release(); c->phdr.release(); c->brec.release();
    c->contacts.release(); c->cnt.release(); c->offs.release(); c->fill.release(); c->perm.release();
    c->perm_sorted.release(); c->skey.release(); c->skey_sorted.release(); c->cub_tmp2.release(); c->bkey.release(); c->bidx.release(); c->bbody.release(); c->counters.release(); c->acc_imp.release(); c->acc_fric.release();
    c->rigid.release(); c->zone_lists.release(); c->strain.release(); c->d_rg_pts.release(); c->d_rg_state.release();
    if (c->dist.on) {
        clsn_ctx::Dist& d = c->dist;
        if (d.peer_ipc)
            for (int r = 0; r < c->nranks; ++r)
                if (r != c->rank && d.peer_recv[r] && !d.peer_local[r]) { cudaIpcCloseMemHandle(d.peer_recv[r]); cudaIpcCloseMemHandle(d.peer_hdr[r]); }
        if (d.comm) d.api.CommDestroy(d.comm);
        d.recv.release(); d.hdr.release(); d.d_peer_region.release(); d.d_peer_hdr.release(); d.send_cnt.release();
        d.stage.release(); d.d_stage_region.release();
        d.gsum.release(); d.maxblk.release(); d.allmax.release(); d.brec_all.release(); d.brec_dense.release(); d.brec_n.release();
        if (d.h_gsum) cudaFreeHost(d.h_gsum);
    }
    if (c->h_counters) cudaFreeHost(c->h_counters);
    if (c->h_pin) cudaFreeHost(c->h_pin);
    for (auto& e : c->ev) cudaEventDestroy(e);
    for (auto& e : c->marks) cudaEventDestroy(e);
    cudaEventDestroy(c->bracket[0]);
    cudaEventDestroy(c->bracket[1]);
    cudaStreamDestroy(c->own_stream);
    delete c;
}

extern "C" const char* clsn_last_error(const clsn_ctx* c) { return c ? c->err.c_str() : "null context"; }

extern "C" int clsn_set_params(clsn_ctx* c, const clsn_params* p)
{
    if (!c || !p) return CLSN_E_ARG;
    c->prm = *p;
    return CLSN_OK;
}

extern "C" int clsn_set_slice(clsn_ctx* c, int rank, int nranks)
{
    if (!c || nranks < 1 || rank < 0 || rank >= nranks) return CLSN_E_ARG;
    c->rank = rank;
    c->nranks = nranks;
    return CLSN_OK;
}

// ------------------------------------------------------------------ in-library multi-GPU (dist.cuh)
#define NCK(call)                                                                                   \
    do {                                                                                            \
        ncclResult_t _r = (call);                                                                   \
        if (_r != ncclSuccess) {                                                                    \
            c->err = std::string(#call) + ": " + c->dist.api.GetErrorString(_r);                    \
            return CLSN_E_CUDA;                                                                     \
        }                                                                                           \
    } while (0)

static inline unsigned long long* gsum_block(clsn_ctx* c, int slot) { return c->dist.gsum.p + CTR_STRIDE * slot; }

extern "C" int clsn_dist_unique_id(void* id128)
{
    if (!id128) return CLSN_E_ARG;
    static NcclApi api;
    std::string err;
    if (!api.load(err)) return CLSN_E_UNSUPPORTED;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId");
    return api.GetUniqueId(reinterpret_cast<ncclUniqueId*>(id128)) == ncclSuccess ? CLSN_OK : CLSN_E_CUDA;
}

// (re)allocate the receive regions and exchange their addresses: COLLECTIVE over the communicator.  Between processes
// the buffers are mapped with cudaIpc; the handles travel through an NCCL all-gather.
static int dist_alloc_regions(clsn_ctx* c, long long cap_region)
{
    clsn_ctx::Dist& d = c->dist;
    const int G = c->nranks, me = c->rank;
    CK(cudaStreamSynchronize(c->stream));
    if (d.peer_ipc)
        for (int r = 0; r < G; ++r)
            if (r != me && d.peer_recv[r] && !d.peer_local[r]) { cudaIpcCloseMemHandle(d.peer_recv[r]); cudaIpcCloseMemHandle(d.peer_hdr[r]); }
    const bool had_buffers = d.recv.p != nullptr;
    for (int r = 0; r < CLSN_MAX_RANKS; ++r) d.peer_recv[r] = d.peer_hdr[r] = nullptr;
    if (had_buffers) {
        // nobody frees a buffer that a peer still has mapped: all ranks have closed their mappings before any rank goes on
        NCK(d.api.AllReduce(d.maxblk.p, d.maxblk.p, 4, ncclUint64, ncclMax, d.comm, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    d.recv.release();
    d.hdr.release();
    d.cap_region = cap_region;
    CK(d.recv.reserve((size_t)G * (size_t)cap_region));
    CK(d.hdr.reserve(CLSN_MAX_RANKS));
    CK(cudaMemset(d.hdr.p, 0, CLSN_MAX_RANKS * sizeof(unsigned long long)));
    // handles: [recv handle | hdr handle] per rank
    struct Handles { cudaIpcMemHandle_t recv, hdr; long long pid; int device; int pad; void* recv_ptr; void* hdr_ptr; };
    Handles mine;
    memset(&mine, 0, sizeof(mine));
    CK(cudaIpcGetMemHandle(&mine.recv, d.recv.p));
    CK(cudaIpcGetMemHandle(&mine.hdr, d.hdr.p));
    mine.pid = (long long)getpid();
    mine.device = c->device;
    mine.recv_ptr = d.recv.p;
    mine.hdr_ptr = d.hdr.p;
    DevBuf<Handles> dh;
    CK(dh.reserve((size_t)G));
    CK(cudaMemcpy(dh.p + me, &mine, sizeof(mine), cudaMemcpyHostToDevice));
    NCK(d.api.AllGather(dh.p + me, dh.p, sizeof(Handles), ncclUint8, d.comm, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    std::vector<Handles> all((size_t)G);
    CK(cudaMemcpy(all.data(), dh.p, (size_t)G * sizeof(Handles), cudaMemcpyDeviceToHost));
    dh.release();
    d.peer_ipc = true;
    bool* same_process = d.peer_local;
    for (int r = 0; r < CLSN_MAX_RANKS; ++r) same_process[r] = false;
    std::vector<PointRec*> region((size_t)G);
    std::vector<unsigned long long*> hdrp((size_t)G);
    for (int r = 0; r < G; ++r) {
        if (r == me) {
            d.peer_recv[r] = d.recv.p;
            d.peer_hdr[r] = d.hdr.p;
        } else if (all[r].pid == mine.pid) {
            // the peer context lives in this process (one host thread per GPU): plain peer access, no IPC mapping
            const cudaError_t e = cudaDeviceEnablePeerAccess(all[r].device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
            cudaGetLastError();
            d.peer_recv[r] = all[r].recv_ptr;
            d.peer_hdr[r] = all[r].hdr_ptr;
            same_process[r] = true;
        } else {
            CK(cudaIpcOpenMemHandle(&d.peer_recv[r], all[r].recv, cudaIpcMemLazyEnablePeerAccess));
            CK(cudaIpcOpenMemHandle(&d.peer_hdr[r], all[r].hdr, cudaIpcMemLazyEnablePeerAccess));
        }
        region[r] = reinterpret_cast<PointRec*>(d.peer_recv[r]) + (size_t)me * (size_t)cap_region;   // my region inside rank r
        hdrp[r] = reinterpret_cast<unsigned long long*>(d.peer_hdr[r]);
    }
    CK(d.d_peer_region.reserve(CLSN_MAX_RANKS));
    CK(d.d_peer_hdr.reserve(CLSN_MAX_RANKS));
    CK(cudaMemcpy(d.d_peer_region.p, region.data(), (size_t)G * sizeof(PointRec*), cudaMemcpyHostToDevice));
    d.stage.release();
    CK(d.stage.reserve((size_t)G * (size_t)cap_region));
    d.direct = getenv("CLSN_DIST_DIRECT") != nullptr;   // A/B: store every record straight into the owner's memory, no staging
    if (!d.direct)
        for (int r = 0; r < G; ++r) region[r] = r == me ? d.recv.p + (size_t)me * (size_t)cap_region : d.stage.p + (size_t)r * (size_t)cap_region;
    CK(d.d_stage_region.reserve(CLSN_MAX_RANKS));
    CK(cudaMemcpy(d.d_stage_region.p, region.data(), (size_t)G * sizeof(PointRec*), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.d_peer_hdr.p, hdrp.data(), (size_t)G * sizeof(unsigned long long*), cudaMemcpyHostToDevice));
    const size_t tot = (size_t)G * (size_t)cap_region;
    CK(c->perm.reserve(tot)); CK(c->perm_sorted.reserve(tot)); CK(c->skey.reserve(tot));
    return CLSN_OK;
}

// Join a communicator of `nranks` contexts, one per GPU of this node (one process or thread each).  id128: the 128 bytes
// of clsn_dist_unique_id() called once by any rank and handed to all of them (MPI_Bcast, torch.distributed ...).
// Call after clsn_set_topology.  Afterwards clsn_resolve / clsn_step_host run the sliced step: every rank uploads the
// same state and ends with the same, complete state (bit-identical to one GPU).
extern "C" int clsn_dist_init(clsn_ctx* c, int rank, int nranks, const void* id128)
{
    if (!c || !c->V || !id128 || nranks < 1 || nranks > CLSN_MAX_RANKS || rank < 0 || rank >= nranks) return CLSN_E_ARG;
    cudaSetDevice(c->device);
    clsn_ctx::Dist& d = c->dist;
    if (d.on) return fail(c, CLSN_E_ARG, "clsn_dist_init called twice");
    if (!d.api.load(c->err)) return CLSN_E_UNSUPPORTED;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    NCK(d.api.CommInitRank(&d.comm, nranks, id, rank));
    c->rank = rank;
    c->nranks = nranks;
    d.per_rank = (c->V + nranks - 1) / nranks;
    CK(d.send_cnt.reserve(CLSN_MAX_RANKS));
    CK(d.gsum.reserve((size_t)PASS_SLOTS * CTR_STRIDE));
    CK(d.maxblk.reserve(4));
    CK(d.allmax.reserve((size_t)PASS_SLOTS * CLSN_MAX_RANKS * 4));
    CK(d.brec_all.reserve((size_t)nranks * (size_t)d.cap_brec_x));
    CK(d.brec_dense.reserve((size_t)nranks * (size_t)d.cap_brec_x));
    CK(d.brec_n.reserve(4));
    CK(cudaMemset(d.brec_n.p, 0, 4 * sizeof(unsigned long long)));
    CK(cudaMemset(d.gsum.p, 0, (size_t)PASS_SLOTS * CTR_STRIDE * sizeof(unsigned long long)));
    CK(cudaMallocHost((void**)&d.h_gsum, ((size_t)PASS_SLOTS * CTR_STRIDE + (size_t)PASS_SLOTS * CLSN_MAX_RANKS * 4) * sizeof(unsigned long long)));
    if (c->brec.n < (size_t)d.cap_brec_x) CK(c->brec.reserve((size_t)d.cap_brec_x));
    d.on = true;
    // receive regions: ~64 records per element spread over G x G (source, owner) regions; grown on demand
    long long cap = std::max<long long>(1 << 14, (64ll * c->N) / ((long long)nranks * nranks) + 1024);
    if (const char* e = getenv("CLSN_DIST_REGION_CAP")) cap = std::max<long long>(64, atoll(e));   // tests: force the growth path
    return dist_alloc_regions(c, cap);
}

extern "C" int clsn_dist_nranks(const clsn_ctx* c) { return c && c->dist.on ? c->nranks : 1; }

// Run on the caller's stream (e.g. torch.cuda.current_stream().cuda_stream) so that the caller's own
// device work -- the NCCL exchange of the multi-GPU step -- is ordered with the library's kernels without
// host synchronisation.  stream == NULL restores the context's own stream.
extern "C" int clsn_set_stream(clsn_ctx* c, void* stream)
{
    if (!c) return CLSN_E_ARG;
    cudaSetDevice(c->device);
    CK(cudaStreamSynchronize(c->stream));
    c->stream = stream ? (cudaStream_t)stream : c->own_stream;
    return CLSN_OK;
}

extern "C" int clsn_set_exact_stats(clsn_ctx* c, int on)
{
    if (!c) return CLSN_E_ARG;
    c->exact_stats = on != 0;
    return CLSN_OK;
}

extern "C" int clsn_set_phase_timing(clsn_ctx* c, int on)
{
    if (!c) return CLSN_E_ARG;
    c->phase_timing = on != 0;
    return CLSN_OK;
}

extern "C" int clsn_set_pipeline(clsn_ctx* c, int pipeline)
{
    if (!c || pipeline < 0 || pipeline > 2) return CLSN_E_ARG;
    c->pipeline = pipeline;
    return CLSN_OK;
}

extern "C" int clsn_set_debug(clsn_ctx* c, int cand, int contacts)
{
    if (!c) return CLSN_E_ARG;
    c->dbg_candidates = cand != 0;
    c->dbg_contacts = contacts != 0;
    return CLSN_OK;
}

extern "C" int clsn_set_topology(clsn_ctx* c, int V, int T, const int32_t* tri_idx, const int32_t* tri_surf, int B,
                                 const int32_t* bond_idx, const uint8_t* vflags, const int32_t* vbody, int nbody,
                                 const double* body_mass)
{
    if (!c || V <= 0 || T < 0 || B < 0 || nbody <= 0 || !vflags || !vbody || !body_mass) return CLSN_E_ARG;
    if ((T > 0 && (!tri_idx || !tri_surf)) || (B > 0 && !bond_idx)) return CLSN_E_ARG;
    if ((long long)T + B >= (1ll << 30)) return fail(c, CLSN_E_ARG, "too many elements for the 30-bit pair key");
    cudaSetDevice(c->device);
    const int N = T + B;
    std::vector<int4> el((size_t)(N > 0 ? N : 1));
    for (int t = 0; t < T; ++t) {
        int4 e;
        e.x = tri_idx[3 * t]; e.y = tri_idx[3 * t + 1]; e.z = tri_idx[3 * t + 2];
        if (e.x < 0 || e.x >= V || e.y < 0 || e.y >= V || e.z < 0 || e.z >= V) return fail(c, CLSN_E_ARG, "triangle index out of range");
        if (tri_surf[t] < 0 || tri_surf[t] >= (1 << 28)) return fail(c, CLSN_E_ARG, "surface id out of range");
        // isRigidBody(CD_HSE*): any vertex fixed or movable-RG (dcollid.cpp:1074-1099)
        const bool rigid = ((vflags[e.x] | vflags[e.y] | vflags[e.z]) & 3) != 0;
        e.w = tri_surf[t] | (rigid ? 0x10000000 : 0);
        el[t] = e;
    }
    for (int b = 0; b < B; ++b) {
        int4 e;
        e.x = bond_idx[2 * b]; e.y = bond_idx[2 * b + 1]; e.z = -1;
        if (e.x < 0 || e.x >= V || e.y < 0 || e.y >= V) return fail(c, CLSN_E_ARG, "bond index out of range");
        e.w = 0x20000000;
        el[T + b] = e;
    }
    c->has_movable = false;
    for (int v = 0; v < V; ++v) {
        if (vbody[v] < 0 || vbody[v] >= nbody) return fail(c, CLSN_E_ARG, "vbody out of range");
        if (vflags[v] & 2) c->has_movable = true;
    }
    c->V = V; c->T = T; c->B = B; c->N = N; c->nbody = nbody;
    CK(c->elem.reserve(el.size()));
    CK(c->vflags.reserve(V)); CK(c->vbody.reserve(V)); CK(c->body_mass.reserve(nbody));
    CK(cudaMemcpy(c->elem.p, el.data(), el.size() * sizeof(int4), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->vflags.p, vflags, V, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->vbody.p, vbody, V * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->body_mass.p, body_mass, nbody * sizeof(double), cudaMemcpyHostToDevice));
    CK(c->xo.reserve(V)); CK(c->xn.reserve(V)); CK(c->xf.reserve(V));
    // + CLSN_MAX_RANKS: the in-place all-gather of the multi-GPU step works on ceil(V / G) * G entries
    CK(c->av.reserve((size_t)V + CLSN_MAX_RANKS)); CK(c->has.reserve((size_t)V + CLSN_MAX_RANKS)); CK(c->dirty.reserve((size_t)V + CLSN_MAX_RANKS));
    CK(c->imp_rg.reserve(3 * (size_t)nbody)); CK(c->cnt_rg.reserve(nbody));
    CK(cudaMemset(c->imp_rg.p, 0, 3 * (size_t)nbody * sizeof(double)));
    CK(cudaMemset(c->cnt_rg.p, 0, nbody * sizeof(int)));
    CK(cudaMemset(c->av.p, 0, V * sizeof(Vec4)));
    CK(cudaMemset(c->has.p, 0, V));
    CK(c->stage.reserve(6 * (size_t)V + 2 * CLSN_MAX_RANKS));   // + padding of the sliced multi-GPU upload
    const size_t n1 = (size_t)(N > 0 ? N : 1);
    CK(c->code.reserve(n1)); CK(c->code_sorted.reserve(n1)); CK(c->idx.reserve(n1)); CK(c->leaf_elem.reserve(n1));
    CK(c->selem.reserve(n1)); CK(c->lbox.reserve(6 * n1));
    {   // shape of the implicit 8-ary tree: a function of N alone
        Tree8& tr = c->tree;
        memset(&tr, 0, sizeof(tr));
        tr.N = N;
        tr.cnt[0] = N;
        int total = 0, l = 0;
        do {
            ++l;
            tr.cnt[l] = (tr.cnt[l - 1] + 7) / 8;
            if (tr.cnt[l] < 1) tr.cnt[l] = 1;
            tr.off[l] = total;
            total += tr.cnt[l];
        } while ((tr.cnt[l] > 1 || l < 3) && l + 1 < TREE_MAXLEV);
        tr.nlev = l;
        CK(c->nodes.reserve((size_t)total));
        CK(c->tree_scratch.reserve(12 * (size_t)tr.cnt[3] + 16)); CK(c->tree_scratch_t.reserve(2 * (size_t)tr.cnt[3] + 16));
        CK(c->tree_ticket.reserve(4));
        CK(cudaMemset(c->tree_ticket.p, 0, 4 * sizeof(unsigned)));
        tr.nodes = c->nodes.p;
    }
    size_t tmp1 = 0, tmp2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp1, c->code.p, c->code_sorted.p, c->idx.p, c->leaf_elem.p, N > 0 ? N : 1, 0, 30);
    CK(c->cnt.reserve((size_t)V + 1)); CK(c->offs.reserve((size_t)V + 1)); CK(c->fill.reserve((size_t)V + 1));
    cub::DeviceScan::ExclusiveSum(nullptr, tmp2, c->cnt.p, c->offs.p, V + 1);
    CK(c->cub_tmp.reserve((tmp1 > tmp2 ? tmp1 : tmp2) + 256));
    CK(cudaMemset(c->cnt.p, 0, ((size_t)V + 1) * sizeof(int)));
    if (c->pairs.n == 0) CK(c->pairs.reserve((size_t)16 * n1 + 1024));
    if (c->feats.n == 0) CK(c->feats.reserve((size_t)64 * n1 + 1024));
    if (c->pipeline == 0 && c->rootrecs.n == 0) CK(c->rootrecs.reserve((size_t)16 * n1 + 1024));
    if (c->pipeline >= 1 && c->hits.n == 0) CK(c->hits.reserve((size_t)4 * n1 + 1024));
    CK(c->pair_hit.reserve(c->pairs.n / 32 + 2));
    if (c->prec.n == 0) CK(c->prec.reserve((size_t)8 * n1 + 1024));
    CK(c->perm.reserve(c->prec.n)); CK(c->perm_sorted.reserve(c->prec.n)); CK(c->skey.reserve(c->prec.n));
    if (c->brec.n == 0) CK(c->brec.reserve(4096));
    c->tree_built = false;
    c->records_pending = false;
    c->h_tri.assign(tri_idx, tri_idx + 3 * (size_t)T);
    c->h_tri_surf.assign(tri_surf, tri_surf + T);
    c->h_bond.assign(bond_idx, bond_idx + 2 * (size_t)B);
    c->h_vflags.assign(vflags, vflags + V);
    c->zone_uf_ready = false;
    c->strain.release_schedule();  // rest lengths belong to the old elements: clsn_set_rest_lengths again
    c->strain.have_len0 = false;
    {   // Rigid-rigid impulses are kept per hyper-surface (k_reduce_bodies / k_apply_bodies), the reference spreads them over
        // the union-find list of the hit point (SpreadImpactZoneImpulse, dcollid.cpp:1101-1115).  The two agree iff every
        // movable hyper-surface is ONE connected triangle component and holds movable points only: checked here.
        HostUF uf;
        uf.reset(V);
        uf.merge_movable_bodies(T, tri_idx, tri_surf, vflags);
        std::vector<int> body_root((size_t)nbody, -1);
        std::vector<uint8_t> body_movable((size_t)nbody, 0), body_other((size_t)nbody, 0);
        for (int v = 0; v < V; ++v) {
            if (vflags[v] & 2) {
                body_movable[vbody[v]] = 1;
                const int r = uf.find(v);
                if (body_root[vbody[v]] < 0) body_root[vbody[v]] = r;
                else if (body_root[vbody[v]] != r)
                    return fail(c, CLSN_E_UNSUPPORTED, "a movable rigid hyper-surface must be one connected component (vbody spans several)");
            } else if (!(vflags[v] & 1)) {
                body_other[vbody[v]] = 1;
            }
        }
        for (int b = 0; b < nbody; ++b)
            if (body_movable[b] && body_other[b])
                return fail(c, CLSN_E_UNSUPPORTED, "a hyper-surface mixes movable rigid points with fabric points");
    }
    {   // movable points per body in the order updateFinalForRG meets them (hseList order, points in element order)
        std::vector<std::vector<int>> per((size_t)nbody);
        std::vector<uint8_t> seen((size_t)V, 0);
        auto visit = [&](int p) {
            if (!(vflags[p] & 2) || seen[p]) return;
            seen[p] = 1;
            per[vbody[p]].push_back(p);
        };
        for (size_t i = 0; i < 3 * (size_t)T; ++i) visit(tri_idx[i]);
        for (size_t i = 0; i < 2 * (size_t)B; ++i) visit(bond_idx[i]);
        c->rg_offs.assign(1, 0);
        c->rg_pts.clear();
        for (int b = 0; b < nbody; ++b) {
            c->rg_pts.insert(c->rg_pts.end(), per[b].begin(), per[b].end());
            c->rg_offs.push_back((int)c->rg_pts.size());
        }
        if (!c->rg_pts.empty()) {
            CK(c->d_rg_pts.reserve(c->rg_pts.size()));
            CK(c->d_rg_state.reserve(4 * c->rg_pts.size()));
            CK(cudaMemcpy(c->d_rg_pts.p, c->rg_pts.data(), c->rg_pts.size() * sizeof(int), cudaMemcpyHostToDevice));
        }
        c->mrg_com.assign(3 * (size_t)nbody, 0.0);
        c->mrg_valid.assign((size_t)nbody, 0);
    }
    if (c->dist.on) c->dist.per_rank = (V + c->nranks - 1) / c->nranks;   // vertex ownership follows the new vertex count
    // host-side restatement of createImpZoneForRG's union-find lists (topology only)
    int r = c->rigid.build(V, T, tri_idx, tri_surf, vflags);
    if (r != 0) return fail(c, CLSN_E_CUDA, "rigid-body topology upload failed");
    if (c->h_pin_n < 6 * (size_t)V) {
        if (c->h_pin) cudaFreeHost(c->h_pin);
        CK(cudaMallocHost((void**)&c->h_pin, 6 * (size_t)V * sizeof(double)));
        c->h_pin_n = 6 * (size_t)V;
    }
    return CLSN_OK;
}

static int begin_step(clsn_ctx* c)
{
    // per-step accumulators cleared as FT_Propagate's point hook / recordOriginPosition do
    // (test.cpp:192-224, dcollid.cpp:100): has_collsn; imp/fric/cnt live only inside a pass here
    CK(cudaMemsetAsync(c->has.p, 0, c->V, c->stream));
    if (!c->keep_tree || ++c->tree_age >= 64) c->tree_built = false;
    c->final_valid = false;
    c->records_pending = false;
    c->pending_gate = nullptr;
    c->dirty_valid = false;
    c->last_detect_mode = -1;
    c->zone_uf_ready = false;  // makeSet: the impact zones live for one step (dcollid3d.cpp:44)
    return CLSN_OK;
}

extern "C" int clsn_upload_state_device(clsn_ctx* c, const double* d_x_old, const double* d_x_new)
{
    if (!c || !c->V || !d_x_old || !d_x_new) return CLSN_E_ARG;
    cudaSetDevice(c->device);
    k_pack<<<nblk(c->V, 256), 256, 0, c->stream>>>(c->V, d_x_old, c->xo.p);
    k_pack<<<nblk(c->V, 256), 256, 0, c->stream>>>(c->V, d_x_new, c->xn.p);
    CK(cudaGetLastError());
    c->launches += 2;
    return begin_step(c);
}

extern "C" int clsn_upload_state(clsn_ctx* c, const double* x_old, const double* x_new)
{
    if (!c || !c->V || !x_old || !x_new) return CLSN_E_ARG;
    cudaSetDevice(c->device);
    const size_t n = 3 * (size_t)c->V;
    if (c->dist.on && c->nranks > 1) {
        // Multi-GPU step: every rank is handed the same arrays, so each one moves only ITS share across PCIe and the
        // ranks all-gather the rest over NVLink (G uploads of 24 MB at once were 40 % of the 8-GPU end-to-end step).
        const size_t per = (n + (size_t)c->nranks - 1) / (size_t)c->nranks;     // doubles per rank
        const size_t lo = std::min(n, per * (size_t)c->rank), hi = std::min(n, lo + per);
        double* so = c->stage.p;                         // [G * per] doubles each, in place
        double* sn = c->stage.p + per * (size_t)c->nranks;
        if (hi > lo) {
            CK(cudaMemcpyAsync(so + lo, x_old + lo, (hi - lo) * sizeof(double), cudaMemcpyHostToDevice, c->stream));
            CK(cudaMemcpyAsync(sn + lo, x_new + lo, (hi - lo) * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        }
        NCK(c->dist.api.GroupStart());
        NCK(c->dist.api.AllGather(so + per * (size_t)c->rank, so, per, ncclFloat64, c->dist.comm, c->stream));
        NCK(c->dist.api.AllGather(sn + per * (size_t)c->rank, sn, per, ncclFloat64, c->dist.comm, c->stream));
        NCK(c->dist.api.GroupEnd());
        return clsn_upload_state_device(c, so, sn);
    }
    // straight from the caller's arrays: a true async DMA when they are pinned, driver-staged otherwise
    CK(cudaMemcpyAsync(c->stage.p, x_old, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->stage.p + n, x_new, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    return clsn_upload_state_device(c, c->stage.p, c->stage.p + n);
}

extern "C" int clsn_download_state_device(clsn_ctx* c, double* d_x, double* d_avgvel)
{
    if (!c || !c->V) return CLSN_E_ARG;
    cudaSetDevice(c->device);
    if (d_x) k_unpack<<<nblk(c->V, 256), 256, 0, c->stream>>>(c->V, c->final_valid ? c->xf.p : c->xn.p, d_x);
    if (d_avgvel) k_unpack<<<nblk(c->V, 256), 256, 0, c->stream>>>(c->V, c->av.p, d_avgvel);
    CK(cudaGetLastError());
    c->launches += (d_x ? 1 : 0) + (d_avgvel ? 1 : 0);
    return CLSN_OK;
}

extern "C" int clsn_download_state(clsn_ctx* c, double* x, double* avgvel, uint8_t* has_collsn)
{
    if (!c || !c->V) return CLSN_E_ARG;
    cudaSetDevice(c->device);
    const size_t n = 3 * (size_t)c->V;
    int r = clsn_download_state_device(c, x ? c->stage.p : nullptr, avgvel ? c->stage.p + n : nullptr);
    if (r) return r;
    if (x) CK(cudaMemcpyAsync(x, c->stage.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (avgvel) CK(cudaMemcpyAsync(avgvel, c->stage.p + n, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (has_collsn) CK(cudaMemcpyAsync(has_collsn, c->has.p, c->V, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return CLSN_OK;
}

extern "C" int clsn_set_avgvel(clsn_ctx* c, const double* avgvel)
{
    if (!c || !c->V || !avgvel) return CLSN_E_ARG;
    cudaSetDevice(c->device);
    const size_t n = 3 * (size_t)c->V;
    CK(cudaMemcpyAsync(c->stage.p, avgvel, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    k_pack<<<nblk(c->V, 256), 256, 0, c->stream>>>(c->V, c->stage.p, c->av.p);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    c->dirty_valid = false;
    return CLSN_OK;
}

extern "C" int clsn_avg_velocity(clsn_ctx* c)
{
    if (!c || !c->V) return CLSN_E_ARG;
    cudaSetDevice(c->device);
    k_avg_velocity<<<nblk(c->V, 256), 256, 0, c->stream>>>(c->V, c->xo.p, c->xn.p, c->av.p, c->prm.dt, pass_block(c, PASS_SLOT_STEP));
    CK(cudaGetLastError());
    c->launches += 1;
    c->dirty_valid = false;
    mark(c, PH_AVG);
    return CLSN_OK;
}

// ------------------------------------------------------------------ detection pass
static int build_tree(clsn_ctx* c)
{
    const int N = c->N;
    static const unsigned long long init[6] = {~0ull, ~0ull, ~0ull, 0ull, 0ull, 0ull};
    CK(cudaMemcpyAsync(c->bounds.p, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
    k_scene_bounds<<<c->sm_count * 4, 256, 0, c->stream>>>(c->xo.p, c->V, c->bounds.p);
    k_morton<<<nblk(N, 256), 256, 0, c->stream>>>(c->elem.p, N, c->xo.p, c->bounds.p, c->code.p, c->idx.p);
    size_t tmp = c->cub_tmp.n;
    CK(cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, tmp, c->code.p, c->code_sorted.p, c->idx.p, c->leaf_elem.p, N, 0, 30, c->stream));
    k_gather_elems<<<nblk(N, 256), 256, 0, c->stream>>>(c->elem.p, c->leaf_elem.p, N, c->selem.p);
    CK(cudaGetLastError());
    c->launches += 2 + 4 + 1;  // bounds, morton, radix sort (4 onesweep kernels), element gather
    mark(c, PH_BUILD);
    c->tree_built = true;
    c->tree_age = 0;
    c->tree_volume = -1.0;   // set from the first refit's root box
    return CLSN_OK;
}

// Enqueue one detection pass (refit, self query, narrow phase, record emission) on the stream.  Nothing is read back:
// every kernel takes its input count from the pass's counter block, clamps its output to the buffer capacity and skips
// its work when an upstream list overflowed; finish_detect() looks at the counters later.  gate (device pointer or
// null): the pass runs only if *gate != 0 -- detectCollision's `while (is_collision)` (dcollid.cpp:448) on the device.
static int enqueue_detect(clsn_ctx* c, int mode, int slot, const unsigned long long* gate)
{
    const int N = c->N, V = c->V;
    const bool moving = mode == CLSN_COLLISION;
    if (N < 1) return fail(c, CLSN_E_ARG, "no elements");
    NarrowParams P{c->prm.eps, c->prm.thickness, c->prm.dt, c->prm.k, c->prm.m, c->prm.lambda, c->prm.cr};
    if (!c->tree_built) {
        int r = build_tree(c);
        if (r) return r;
    }
    unsigned long long* ctr = c->ctr = pass_block(c, slot);
    if (c->dbg_candidates && c->dbg_cand.n == 0) CK(c->dbg_cand.reserve((size_t)32 * N + 1024));
    if (c->dbg_contacts && c->contacts.n == 0) CK(c->contacts.reserve((size_t)16 * N + 1024));
    const int q_lo = (int)((long long)N * c->rank / c->nranks), q_hi = (int)((long long)N * (c->rank + 1) / c->nranks);
    CK(cudaMemsetAsync(c->cnt.p, 0, ((size_t)V + 1) * sizeof(int), c->stream));
    CK(cudaMemsetAsync(c->cnt_rg.p, 0, c->nbody * sizeof(int), c->stream));
    // pairs whose points did not change since the previous CCD pass repeat their (hit-free) outcome
    const bool skip_clean = moving && c->dirty_valid && c->last_detect_mode == CLSN_COLLISION;
    // ... and an untouched query needs to visit touched subtrees only -- unless exact candidate counts are wanted
    const bool prune = skip_clean && !c->exact_stats && !c->dbg_candidates;
    float* root_box = c->tree_scratch.p + 12 * (size_t)c->tree.cnt[3];
    if (moving)
        k_refit8<true><<<c->tree.cnt[3], REFIT_LEAVES, 0, c->stream>>>(c->selem.p, c->tree, c->xo.p, c->av.p, c->prm.dt, c->lbox.p,
                                                                        prune ? c->dirty.p : nullptr, c->tree_scratch.p,
                                                                        c->tree_scratch_t.p, c->tree_ticket.p, root_box, gate);
    else
        k_refit8<false><<<c->tree.cnt[3], REFIT_LEAVES, 0, c->stream>>>(c->selem.p, c->tree, c->xo.p, c->av.p, c->prm.dt, c->lbox.p,
                                                                         nullptr, c->tree_scratch.p, c->tree_scratch_t.p,
                                                                         c->tree_ticket.p, root_box, gate);
    c->launches += 1;
    mark(c, PH_REFIT);
    TraverseOut to;
    to.pairs = c->pairs.p; to.cap_pairs = (long long)c->pairs.n;
    to.dbg_cand = c->dbg_candidates ? c->dbg_cand.p : nullptr; to.cap_dbg = (long long)c->dbg_cand.n;
    to.counters = ctr;
    to.dirty = skip_clean ? c->dirty.p : nullptr;
    to.prune = prune ? 1 : 0;
    to.gate = gate;
    if (q_hi > q_lo)
        k_traverse8<<<nblk(q_hi - q_lo, TRAV_THREADS), TRAV_THREADS, 0, c->stream>>>(c->tree, c->lbox.p, c->leaf_elem.p, c->selem.p, q_lo, q_hi, to);
    c->launches += (q_hi > q_lo) ? 1 : 0;
    mark(c, PH_TRAVERSE);
    Emit E;
    E.prec = c->prec.p; E.brec = c->brec.p; E.contacts = c->dbg_contacts ? c->contacts.p : nullptr;
    CK(c->phdr.reserve(c->prec.n));
    E.phdr = c->phdr.p;
    E.counters = ctr;
    E.cap_prec = (long long)c->prec.n; E.cap_brec = (long long)c->brec.n; E.cap_contacts = (long long)c->contacts.n;
    E.cnt = c->cnt.p; E.cnt_rg = c->cnt_rg.p; E.body_mass = c->body_mass.p;
    E.D.nranks = 1; E.D.per_rank = V; E.D.cap_region = 0; E.D.peer_region = nullptr; E.D.send_cnt = nullptr;
    if (c->dist.on) {
        E.D.nranks = c->nranks; E.D.per_rank = c->dist.per_rank; E.D.cap_region = c->dist.cap_region;
        E.D.peer_region = c->dist.d_stage_region.p; E.D.send_cnt = c->dist.send_cnt.p;
        CK(cudaMemsetAsync(c->dist.send_cnt.p, 0, CLSN_MAX_RANKS * sizeof(unsigned long long), c->stream));
    }
    const long long hit_words = (long long)(c->pairs.n / 32 + 1);
    CK(cudaMemsetAsync(c->pair_hit.p, 0, (size_t)hit_words * sizeof(unsigned), c->stream));
    const int grid = c->sm_count * NARROW_GRID_MULT;
    const bool fused = c->pipeline >= 1;
    // segments only where this context reduces its own records (multi-GPU ranks exchange the plain list)
    const bool seg = moving && c->pipeline == 2 && c->nranks == 1;
    // The pair list is processed in chunks of `pair_chunk` pairs, so that the feature / hit lists stay bounded however many
    // pairs a pass produces (fast rigid bodies in a cloth stack: 10^8 pairs, 15 surviving features each).  The number of
    // chunks follows from the CAPACITY of the pair list (the count lives on the device); a chunk beyond the count is a
    // handful of empty launches.  One chunk covers every scene of the size of config 4.
    const long long cap_pairs = (long long)c->pairs.n;
    const int nchunk = (c->pipeline == 1 || !moving) ? (int)std::max<long long>(1, (cap_pairs + c->pair_chunk - 1) / c->pair_chunk) : 1;
    if (moving) {
        if (fused && c->hits.n == 0) CK(c->hits.reserve((size_t)4 * N + 1024));
        if (fused && c->unc.n == 0) CK(c->unc.reserve((size_t)4 * N + 1024));
        if (!fused && c->rootrecs.n == 0) CK(c->rootrecs.reserve((size_t)16 * N + 1024));
    }
    for (int ch = 0; ch < nchunk; ++ch) {
        const long long lo = nchunk == 1 ? 0 : (long long)ch * c->pair_chunk;
        const long long hi = nchunk == 1 ? cap_pairs : std::min<long long>(cap_pairs, lo + c->pair_chunk);
        if (moving) {
            k_cull<true><<<grid, CULL_THREADS, 0, c->stream>>>(c->pairs.p, cap_pairs, c->elem.p, c->xo.p, c->av.p, P, c->feats.p,
                                                                (long long)c->feats.n, ctr, fused, lo, hi);
            if (ch == nchunk - 1) mark(c, PH_CULL);
            if (fused) {
                k_fast<false><<<grid, FEAT_THREADS, 0, c->stream>>>(c->feats.p, (long long)c->feats.n, c->xo.p, c->av.p, P, E, c->hits.p,
                                                                     (long long)c->hits.n, c->unc.p, (long long)c->unc.n);
                k_fast<true><<<grid, FEAT_THREADS, 0, c->stream>>>(c->feats.p, (long long)c->feats.n, c->xo.p, c->av.p, P, E, c->hits.p,
                                                                    (long long)c->hits.n, c->unc.p, (long long)c->unc.n);
                k_exact<false><<<grid, FEAT_THREADS, 0, c->stream>>>(c->unc.p, (long long)c->unc.n, c->xo.p, c->av.p, P, E, c->hits.p,
                                                                      (long long)c->hits.n);
                k_exact<true><<<grid, FEAT_THREADS, 0, c->stream>>>(c->unc.p, (long long)c->unc.n, c->xo.p, c->av.p, P, E, c->hits.p,
                                                                     (long long)c->hits.n);
                if (ch == nchunk - 1) mark(c, PH_ROOTS);
                if (seg) {
                    // per-point record counts of the hit list -> segment offsets -> records written in place
                    k_count_hits<false><<<grid, 256, 0, c->stream>>>(c->hits.p, (long long)c->hits.n, c->vflags.p, c->cnt.p, ctr);
                    k_count_hits<true><<<grid, 256, 0, c->stream>>>(c->hits.p, (long long)c->hits.n, c->vflags.p, c->cnt.p, ctr);
                    size_t tmp = c->cub_tmp.n;
                    CK(cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, tmp, c->cnt.p, c->offs.p, V + 1, c->stream));
                    CK(cudaMemsetAsync(c->fill.p, 0, ((size_t)V + 1) * sizeof(int), c->stream));
                    const SegOut S{c->offs.p, c->fill.p};
                    k_emit<false, true><<<grid, FEAT_THREADS, 0, c->stream>>>(c->hits.p, (long long)c->hits.n, c->pairs.p, c->xo.p,
                                                                               c->av.p, c->vflags.p, c->vbody.p, P, E, c->pair_hit.p, S);
                    k_emit<true, true><<<grid, FEAT_THREADS, 0, c->stream>>>(c->hits.p, (long long)c->hits.n, c->pairs.p, c->xo.p,
                                                                              c->av.p, c->vflags.p, c->vbody.p, P, E, c->pair_hit.p, S);
                    c->launches += 4;
                } else {
                    k_emit<false><<<grid, FEAT_THREADS, 0, c->stream>>>(c->hits.p, (long long)c->hits.n, c->pairs.p, c->xo.p, c->av.p,
                                                                         c->vflags.p, c->vbody.p, P, E, c->pair_hit.p);
                    k_emit<true><<<grid, FEAT_THREADS, 0, c->stream>>>(c->hits.p, (long long)c->hits.n, c->pairs.p, c->xo.p, c->av.p,
                                                                        c->vflags.p, c->vbody.p, P, E, c->pair_hit.p);
                }
            } else {
                k_roots<<<grid, FEAT_THREADS, 0, c->stream>>>(c->feats.p, (long long)c->feats.n, c->xo.p, c->av.p, P.dt, c->rootrecs.p,
                                                               (long long)c->rootrecs.n, ctr);
                mark(c, PH_ROOTS);
                k_contact<true><<<grid, FEAT_THREADS, 0, c->stream>>>(c->feats.p, c->rootrecs.p, (long long)c->rootrecs.n, c->pairs.p,
                                                                       c->xo.p, c->av.p, c->vflags.p, c->vbody.p, P, E, c->pair_hit.p);
            }
        } else {
            k_cull<false><<<grid, CULL_THREADS, 0, c->stream>>>(c->pairs.p, cap_pairs, c->elem.p, c->xo.p, c->av.p, P, c->feats.p,
                                                                 (long long)c->feats.n, ctr, false, lo, hi);
            if (ch == nchunk - 1) mark(c, PH_CULL);
            k_contact<false><<<grid, FEAT_THREADS, 0, c->stream>>>(c->feats.p, c->rootrecs.p, (long long)c->feats.n, c->pairs.p,
                                                                    c->xo.p, c->av.p, c->vflags.p, c->vbody.p, P, E, c->pair_hit.p);
        }
        k_fold_chunk<<<1, 32, 0, c->stream>>>(ctr, ch == nchunk - 1 ? 1 : 0);
        c->launches += 1 + (moving ? (fused ? 7 : 3) : 2);
    }
    k_count_true<<<c->sm_count * 2, 256, 0, c->stream>>>(c->pair_hit.p, hit_words, ctr);
    CK(cudaGetLastError());
    c->launches += 1;
    if (c->dist.on) {
        // counts to the owners, then the all-reduce of the pass's counter block: global counts for the device-side gate
        // of the next pass, and the barrier after which every record pushed to this rank has landed
        clsn_ctx::Dist& d = c->dist;
        PublishCaps caps;
        caps.pairs = (long long)c->pairs.n; caps.feats = (long long)c->feats.n;
        caps.unc = moving && fused ? (long long)c->unc.n : (1ll << 62);
        caps.hits = moving && fused ? (long long)c->hits.n : (1ll << 62);
        caps.brec = std::min<long long>((long long)c->brec.n, d.cap_brec_x); caps.region = d.cap_region;
        if (!d.direct)
            k_push_regions<<<dim3((unsigned)std::max(1, c->sm_count * 8 / c->nranks), (unsigned)c->nranks), 256, 0, c->stream>>>(
                c->rank, d.cap_region, d.stage.p, d.d_peer_region.p, d.send_cnt.p);
        k_publish<<<1, CLSN_MAX_RANKS, 0, c->stream>>>(c->nranks, c->rank, d.send_cnt.p, d.d_peer_hdr.p, ctr, d.maxblk.p, caps);
        CK(cudaGetLastError());
        c->launches += 2;
        NCK(d.api.GroupStart());
        NCK(d.api.AllReduce(ctr, gsum_block(c, slot), CTR_STRIDE, ncclUint64, ncclSum, d.comm, c->stream));
        NCK(d.api.AllGather(d.maxblk.p, d.allmax.p + (size_t)slot * CLSN_MAX_RANKS * 4, 4, ncclUint64, d.comm, c->stream));
        if (c->has_movable)
            NCK(d.api.AllGather(c->brec.p, d.brec_all.p, (size_t)d.cap_brec_x * sizeof(BodyRec), ncclUint8, d.comm, c->stream));
        NCK(d.api.GroupEnd());
    }
    mark(c, PH_CONTACT);
    c->records_pending = true;
    c->seg_records = seg;
    c->imp_nprec = -1;
    c->last_detect_mode = mode;
    c->dirty_valid = false;  // becomes valid again once these records have been applied
    c->pending_gate = gate;
    return CLSN_OK;
}

// Look at a pass's counters (host copy h): grow every list that overflowed.  Returns 1 when the pass (or the step that
// contains it) has to be repeated, 0 when its results stand, < 0 on error.
static int grow_after_pass(clsn_ctx* c, const unsigned long long* h, bool moving)
{
    const bool fused = c->pipeline >= 1;
    bool redo = false;
    if (h[CTR_PAIRS] >= (1ull << 32)) return fail(c, CLSN_E_NOMEM, "more than 2^32 - 1 pairs in one pass");
    if (h[CTR_PAIRS] > c->pairs.n) {
        size_t want = (size_t)(h[CTR_PAIRS] * 5 / 4 + 1024);
        if (want >= (1ull << 32)) want = (1ull << 32) - 1;
        CK(c->pairs.reserve(want));
        CK(c->pair_hit.reserve(c->pairs.n / 32 + 2));
        redo = true;
    }
    if (h[CTR_MAX_FEATS] > c->feats.n) {   // lists that restart with every chunk of the pair list: the largest chunk counts
        CK(c->feats.reserve((size_t)(h[CTR_MAX_FEATS] * 5 / 4 + 1024)));
        redo = true;
    }
    if (!fused && moving && h[CTR_ROOTS] + h[CTR_ROOTS_EE] > c->rootrecs.n) {
        CK(c->rootrecs.reserve((size_t)((h[CTR_ROOTS] + h[CTR_ROOTS_EE]) * 5 / 4 + 1024)));
        redo = true;
    }
    if (fused && moving && h[CTR_MAX_UNC] > c->unc.n) {
        CK(c->unc.reserve((size_t)(h[CTR_MAX_UNC] * 5 / 4 + 1024)));
        redo = true;
    }
    if (fused && moving && h[CTR_MAX_HITS] > c->hits.n) {
        CK(c->hits.reserve((size_t)(h[CTR_MAX_HITS] * 5 / 4 + 1024)));
        redo = true;
    }
    if (!c->dist.on && h[CTR_PREC] > c->prec.n) {   // (multi-GPU: point records live in the receive regions, not here)
        size_t want = (size_t)(h[CTR_PREC] * 5 / 4 + 1024);
        CK(c->prec.reserve(want)); CK(c->perm.reserve(want)); CK(c->perm_sorted.reserve(want)); CK(c->skey.reserve(want));
        redo = true;
    }
    if (h[CTR_BREC] > c->brec.n) { CK(c->brec.reserve((size_t)(h[CTR_BREC] * 5 / 4 + 1024))); redo = true; }
    if (c->dbg_candidates && h[CTR_DBG_CAND] > c->dbg_cand.n) { CK(c->dbg_cand.reserve((size_t)(h[CTR_DBG_CAND] * 5 / 4 + 1024))); redo = true; }
    if (c->dbg_contacts && h[CTR_CONTACTS] > c->contacts.n) { CK(c->contacts.reserve((size_t)(h[CTR_CONTACTS] * 5 / 4 + 1024))); redo = true; }
    return redo ? 1 : 0;
}

static void fill_pass_stats(clsn_ctx* c, const unsigned long long* h, bool moving, clsn_pass_stats* st)
{
    if (!st) return;
    const bool fused = c->pipeline >= 1;
    st->candidates = (int64_t)h[CTR_CAND];
    st->pairs_tested = (int64_t)h[CTR_PAIRS];
    st->true_pairs = (int64_t)h[CTR_TRUE];
    st->contacts = (int64_t)h[CTR_CONTACTS];
    st->contributions = (int64_t)(h[CTR_PREC] + h[CTR_BREC]);
    st->features = (int64_t)(h[CTR_FEATS] + h[CTR_FEATS_EE]);
    st->box_survivors = (int64_t)h[CTR_BOXSURV];
    st->coplanar = (int64_t)(h[CTR_ROOTS] + h[CTR_ROOTS_EE]);
    st->exact_solves = fused && moving ? (int64_t)h[CTR_EXACT] : st->coplanar;
}

static int read_blocks(clsn_ctx* c);
static int dist_grow(clsn_ctx* c, int first_slot, int last_slot);

// One pass through the per-phase ABI (and the impact-zone loop): enqueue, read the counters back, repeat if a list was
// too small.  The whole-step entry (resolve_impl) enqueues all passes first and reads every block once at the end.
static int run_detect(clsn_ctx* c, int mode, clsn_pass_stats* st)
{
    const bool moving = mode == CLSN_COLLISION;
    const bool dirty_valid = c->dirty_valid;
    const int last_mode = c->last_detect_mode;
    for (int attempt = 0; attempt < 8; ++attempt) {
        unsigned long long* blk = pass_block(c, PASS_SLOT_SOLO);
        CK(cudaMemsetAsync(blk, 0, CTR_STRIDE * sizeof(unsigned long long), c->stream));
        c->dirty_valid = dirty_valid;   // enqueue_detect clears them; a repeated pass starts from the same flags
        c->last_detect_mode = last_mode;
        int r = enqueue_detect(c, mode, PASS_SLOT_SOLO, nullptr);
        if (r) return r;
        unsigned long long* h = host_block(c, PASS_SLOT_SOLO);
        if (c->dist.on) {
            if ((r = read_blocks(c))) return r;
            r = grow_after_pass(c, h, moving);
            if (r < 0) return r;
            r = dist_grow(c, PASS_SLOT_SOLO, PASS_SLOT_SOLO);
            if (r < 0) return r;
            if (r == 1) continue;
            h = c->dist.h_gsum + CTR_STRIDE * PASS_SLOT_SOLO;   // global counts
        } else {
            CK(cudaMemcpyAsync(h, blk, CTR_STRIDE * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
            CK(cudaStreamSynchronize(c->stream));
            r = grow_after_pass(c, h, moving);
            if (r < 0) return r;
            if (r == 1) continue;
        }
        if (h[CTR_ERROR]) return fail(c, CLSN_E_NUMERIC, "NaN/Inf or degenerate normal (reference: clean_up(ERROR))");
        fill_pass_stats(c, h, moving, st);
        c->last_nprec = (long long)h[CTR_PREC];
        c->last_nbrec = (long long)h[CTR_BREC];
        c->last_true = (long long)h[CTR_TRUE];
        c->n_dbg_cand = (long long)h[CTR_DBG_CAND];
        c->n_contacts = (long long)h[CTR_CONTACTS];
        return CLSN_OK;
    }
    return fail(c, CLSN_E_NOMEM, "pair/record buffers kept overflowing");
}

extern "C" int clsn_detect(clsn_ctx* c, int mode, clsn_pass_stats* st)
{
    if (!c || !c->V || (mode != CLSN_PROXIMITY && mode != CLSN_COLLISION)) return CLSN_E_ARG;
    cudaSetDevice(c->device);
    return run_detect(c, mode, st);
}

// collsnImpulse_RG += the body records of the pass, per body in canonical key order (reduce.cuh)
static int reduce_bodies(clsn_ctx* c, const BodyRec* rec, const unsigned long long* n_dev, long long cap, double* imp_rg)
{
    if (cap <= 0) return CLSN_OK;
    const size_t n = (size_t)cap;
    CK(c->bkey.reserve(2 * n)); CK(c->bidx.reserve(3 * n)); CK(c->bbody.reserve(2 * n));
    unsigned long long *k0 = c->bkey.p, *k1 = c->bkey.p + n;
    int *i0 = c->bidx.p, *i1 = c->bidx.p + n, *i2 = c->bidx.p + 2 * n, *b0 = c->bbody.p, *b1 = c->bbody.p + n;
    size_t t1 = 0, t2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t1, k0, k1, i0, i1, (int)n, 0, 64, c->stream);
    cub::DeviceRadixSort::SortPairs(nullptr, t2, b0, b1, i1, i2, (int)n, 0, 31, c->stream);
    CK(c->cub_tmp2.reserve(std::max(t1, t2) + 256));
    const int grid = std::max(1, std::min(c->sm_count * 4, (int)((n + 255) / 256)));
    k_body_keys<<<grid, 256, 0, c->stream>>>(rec, n_dev, cap, k0, i0);
    size_t tmp = c->cub_tmp2.n;
    CK(cub::DeviceRadixSort::SortPairs(c->cub_tmp2.p, tmp, k0, k1, i0, i1, (int)n, 0, 64, c->stream));
    k_body_ids<<<grid, 256, 0, c->stream>>>(rec, i1, cap, b0);
    tmp = c->cub_tmp2.n;
    CK(cub::DeviceRadixSort::SortPairs(c->cub_tmp2.p, tmp, b0, b1, i1, i2, (int)n, 0, 31, c->stream));
    k_reduce_bodies_sorted<<<nblk(c->nbody, 64), 64, 0, c->stream>>>(rec, b1, i2, cap, c->nbody, imp_rg);
    CK(cudaGetLastError());
    c->launches += 3 + 2 * 9;
    return CLSN_OK;
}

// Sum the grouped records per point in canonical key order and apply / store the sums (k_reduce_points).  With movable
// rigid bodies a single vertex can own tens of thousands of records (a sphere sweeping through a cloth stack): the
// per-point order then comes from cub::DeviceSegmentedSort over (key, record index) instead of the warp's all-pairs
// ranking, which is quadratic in the records of one point.
static int launch_reduce_points(clsn_ctx* c, const PointRec* rec, bool seg, int mode, const unsigned long long* n_rec_ptr, long long cap)
{
    const int V = c->V;
    const bool sorted = c->has_movable && !seg;
    const int* perm = c->perm.p;
    if (sorted) {
        CK(c->skey_sorted.reserve(c->skey.n));
        size_t need = 0;
        cub::DeviceSegmentedSort::SortPairs(nullptr, need, c->skey.p, c->skey_sorted.p, c->perm.p, c->perm_sorted.p, (long long)c->perm.n, V,
                                            c->offs.p, c->offs.p + 1, c->stream);
        CK(c->cub_tmp2.reserve(need + 256));
        k_clamp_offsets<<<nblk(V + 1, 256), 256, 0, c->stream>>>(V + 1, c->offs.p, (int)std::min<size_t>(c->perm.n, 0x7fffffff));
        size_t tmp = c->cub_tmp2.n;
        CK(cub::DeviceSegmentedSort::SortPairs(c->cub_tmp2.p, tmp, c->skey.p, c->skey_sorted.p, c->perm.p, c->perm_sorted.p,
                                               (long long)c->perm.n, V, c->offs.p, c->offs.p + 1, c->stream));
        perm = c->perm_sorted.p;
        k_reduce_points<false, true><<<c->sm_count * 8, 256, 0, c->stream>>>(rec, c->offs.p, c->cnt.p, V, perm, nullptr, c->skey.p, c->vflags.p,
                                                                               c->av.p, c->has.p, c->dirty.p, mode, c->acc_imp.p,
                                                                               c->acc_fric.p, c->ctr, n_rec_ptr, cap);
    } else if (seg) {
        k_reduce_points<true><<<c->sm_count * 8, 256, 0, c->stream>>>(rec, c->offs.p, c->cnt.p, V, perm, c->perm_sorted.p, c->skey.p,
                                                                        c->vflags.p, c->av.p, c->has.p, c->dirty.p, mode, c->acc_imp.p,
                                                                        c->acc_fric.p, c->ctr, n_rec_ptr, cap);
    } else {
        k_reduce_points<false><<<c->sm_count * 8, 256, 0, c->stream>>>(rec, c->offs.p, c->cnt.p, V, perm, c->perm_sorted.p, c->skey.p,
                                                                         c->vflags.p, c->av.p, c->has.p, c->dirty.p, mode, c->acc_imp.p,
                                                                         c->acc_fric.p, c->ctr, n_rec_ptr, cap);
    }
    CK(cudaGetLastError());
    return CLSN_OK;
}

// group + reduce the pending records.  mode 0: apply to avgVel; mode 1: into acc arrays only.
// No host-side counts are consulted for the context's own records: every kernel reads its count from the pass's counter
// block, so the launches are the same whether or not the pass found anything (an empty or gated pass costs a few empty
// kernels).  An imported record set (multi-GPU exchange) carries its counts explicitly.
static int reduce_records(clsn_ctx* c, int mode, int what = 3)  // what: bit 0 point records, bit 1 body records
{
    const int V = c->V;
    const PointRec* prec = c->prec.p;
    const BodyRec* brec = c->brec.p;
    const bool imported = c->imp_nprec >= 0;
    unsigned long long* n_prec_dev = c->ctr + CTR_PREC;
    unsigned long long* n_brec_dev = c->ctr + CTR_BREC;
    long long cap_p = (long long)c->prec.n, cap_b = (long long)c->brec.n;
    if (imported) {
        // externally gathered record set: recount per point
        prec = c->imp_prec; brec = c->imp_brec;
        const long long nprec = c->imp_nprec, nbrec = c->imp_nbrec;
        cap_p = nprec; cap_b = nbrec;
        n_prec_dev = c->counters.p + 32;
        n_brec_dev = c->counters.p + 33;
        if (what & 1) {
            unsigned long long hc[2] = {(unsigned long long)nprec, (unsigned long long)nbrec};
            CK(cudaMemcpyAsync(c->counters.p + 32, hc, sizeof(hc), cudaMemcpyHostToDevice, c->stream));
            CK(cudaMemsetAsync(c->cnt.p, 0, ((size_t)V + 1) * sizeof(int), c->stream));
            CK(cudaMemsetAsync(c->cnt_rg.p, 0, c->nbody * sizeof(int), c->stream));
            if (nprec > 0 || nbrec > 0)
                k_count_records<<<c->sm_count * 4, 256, 0, c->stream>>>(prec, nprec, brec, nbrec, c->cnt.p, c->cnt_rg.p);
            if ((size_t)nprec > c->perm.n) {
                CK(c->perm.reserve((size_t)nprec)); CK(c->perm_sorted.reserve((size_t)nprec)); CK(c->skey.reserve((size_t)nprec));
            }
        }
    }
    const bool seg = c->seg_records && !imported;   // records already grouped per point by k_emit<., true>
    if (what & 1) {
        if (!seg) {
            size_t tmp = c->cub_tmp.n;
            CK(cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, tmp, c->cnt.p, c->offs.p, V + 1, c->stream));
            CK(cudaMemsetAsync(c->fill.p, 0, ((size_t)V + 1) * sizeof(int), c->stream));
            k_scatter<<<c->sm_count * 8, 256, 0, c->stream>>>(prec, imported ? nullptr : c->phdr.p, n_prec_dev, cap_p, c->offs.p, c->fill.p,
                                                              c->perm.p, c->skey.p);
        }
        if (mode == 1) {
            CK(c->acc_imp.reserve(3 * (size_t)V)); CK(c->acc_fric.reserve(3 * (size_t)V));
            CK(cudaMemsetAsync(c->acc_imp.p, 0, 3 * (size_t)V * sizeof(double), c->stream));
            CK(cudaMemsetAsync(c->acc_fric.p, 0, 3 * (size_t)V * sizeof(double), c->stream));
        }
        {
            int r = launch_reduce_points(c, prec, seg, mode, n_prec_dev, cap_p);
            if (r) return r;
        }
        c->launches += 2 + 2;  // scan (init + scan), scatter, reduce
    }
    if (mode == 0 && (what & 2) && (c->has_movable || (imported && c->imp_nbrec > 0))) {
        // body records only exist where a movable rigid body does (emit_body needs a movable point)
        int r = reduce_bodies(c, brec, n_brec_dev, cap_b, c->imp_rg.p);
        if (r) return r;
    }
    CK(cudaGetLastError());
    return CLSN_OK;
}

// updateAverageVelocity of the multi-GPU step (dist.cuh): this rank reduces the records the peers pushed into its receive
// regions (its own vertex range), one in-place all-gather makes avgVel / has_collsn / touched whole again on every rank,
// then the few rigid-rigid body records (all-gathered by enqueue_detect) and the rigid bodies are handled identically
// everywhere.  No host read-back: region counts come from the headers the peers wrote.
static int apply_dist(clsn_ctx* c, int rigidify)
{
    clsn_ctx::Dist& d = c->dist;
    const int V = c->V, G = c->nranks;
    const unsigned long long* gate = c->pending_gate;
    if (c->records_pending) {
        k_reset_dirty<<<nblk(V, 256), 256, 0, c->stream>>>(V, c->vflags.p, c->dirty.p);
        const dim3 grid(c->sm_count * 2, G);
        k_count_regions<<<grid, 256, 0, c->stream>>>(d.recv.p, d.cap_region, d.hdr.p, c->cnt.p);
        size_t tmp = c->cub_tmp.n;
        CK(cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, tmp, c->cnt.p, c->offs.p, V + 1, c->stream));
        CK(cudaMemsetAsync(c->fill.p, 0, ((size_t)V + 1) * sizeof(int), c->stream));
        k_scatter_regions<<<grid, 256, 0, c->stream>>>(d.recv.p, d.cap_region, d.hdr.p, c->offs.p, c->fill.p, c->perm.p, c->skey.p);
        {
            int r = launch_reduce_points(c, d.recv.p, false, 0, d.brec_n.p + 1 /* a zero word */, 1);
            if (r) return r;
        }
        CK(cudaGetLastError());
        c->launches += 6;
        // the state is whole again: every rank contributes the slice it owns (in place)
        const size_t per = (size_t)d.per_rank;
        NCK(d.api.GroupStart());
        NCK(d.api.AllGather(c->av.p + per * c->rank, c->av.p, per * sizeof(Vec4), ncclUint8, d.comm, c->stream));
        NCK(d.api.AllGather(c->has.p + per * c->rank, c->has.p, per, ncclUint8, d.comm, c->stream));
        NCK(d.api.AllGather(c->dirty.p + per * c->rank, c->dirty.p, per, ncclUint8, d.comm, c->stream));
        NCK(d.api.GroupEnd());
        if (c->has_movable) {
            // rigid-rigid contacts: every rank reduces the union of all ranks' body records (ranks in order)
            CK(cudaMemsetAsync(c->cnt_rg.p, 0, c->nbody * sizeof(int), c->stream));
            const int slot = (int)((c->ctr - pass_block(c, 0)) / CTR_STRIDE);
            k_compact_bodies<<<1, 256, 0, c->stream>>>(G, d.cap_brec_x, d.brec_all.p, d.allmax.p + (size_t)slot * CLSN_MAX_RANKS * 4,
                                                       d.brec_dense.p, d.brec_n.p, c->cnt_rg.p);
            {
                int r = reduce_bodies(c, d.brec_dense.p, d.brec_n.p, (long long)d.brec_dense.n, c->imp_rg.p);
                if (r) return r;
            }
            k_apply_bodies<<<nblk(V, 256), 256, 0, c->stream>>>(V, c->vflags.p, c->vbody.p, c->imp_rg.p, c->cnt_rg.p, c->av.p,
                                                                 c->has.p, c->dirty.p);
            CK(cudaMemsetAsync(c->cnt_rg.p, 0, c->nbody * sizeof(int), c->stream));
            c->launches += 3;
        }
        CK(cudaMemsetAsync(c->cnt.p, 0, ((size_t)V + 1) * sizeof(int), c->stream));
        c->records_pending = false;
        c->imp_nprec = -1;
        c->dirty_valid = true;
    }
    if (rigidify && c->has_movable && c->prm.dt > 0.0) {
        int r = c->rigid.rigidify(c->xo.p, c->av.p, c->vflags.p, c->prm.m, c->prm.dt, c->ctr, c->stream, nullptr, gate);
        if (r != 0) return fail(c, CLSN_E_CUDA, "rigid-body kernels failed");
        c->launches += c->rigid.nlists ? 2 : 0;
    }
    c->pending_gate = nullptr;
    CK(cudaGetLastError());
    mark(c, PH_REDUCE);
    return CLSN_OK;
}

// stages: bit 0 = reduce the point records into avgVel; bit 1 = body records, rigid bodies, bookkeeping.
// The multi-GPU owner-computes path runs them separately with the avgVel exchange in between.
static int apply_impl(clsn_ctx* c, int rigidify, int stages)
{
    if (c->dist.on) {
        if (stages != 3) return fail(c, CLSN_E_ARG, "clsn_apply_stage is the external-exchange protocol; not with clsn_dist_init");
        return apply_dist(c, rigidify);
    }
    const unsigned long long* gate = c->pending_gate;   // the apply of a pass that did not run must not touch anything
    if (c->records_pending && (stages & 1)) {
        k_reset_dirty<<<nblk(c->V, 256), 256, 0, c->stream>>>(c->V, c->vflags.p, c->dirty.p);
        c->launches += 1;
        int r = reduce_records(c, 0, 1);
        if (r) return r;
    }
    if (c->records_pending && (stages & 2)) {
        int r = reduce_records(c, 0, 2);
        if (r) return r;
        if (c->has_movable || c->cnt_rg_preset) {
            // cnt_rg > 0 can also persist from set_body_accumulators (parity tests)
            k_apply_bodies<<<nblk(c->V, 256), 256, 0, c->stream>>>(c->V, c->vflags.p, c->vbody.p, c->imp_rg.p, c->cnt_rg.p,
                                                                    c->av.p, c->has.p, c->dirty.p);
            CK(cudaMemsetAsync(c->cnt_rg.p, 0, c->nbody * sizeof(int), c->stream));
            c->cnt_rg_preset = false;
            c->launches += 1;
        }
        CK(cudaMemsetAsync(c->cnt.p, 0, ((size_t)c->V + 1) * sizeof(int), c->stream));
        c->records_pending = false;
        c->imp_nprec = -1;
        c->dirty_valid = true;
    }
    if ((stages & 2) && rigidify && c->has_movable && c->prm.dt > 0.0) {
        int r = c->rigid.rigidify(c->xo.p, c->av.p, c->vflags.p, c->prm.m, c->prm.dt, c->ctr, c->stream, nullptr, gate);
        if (r != 0) return fail(c, CLSN_E_CUDA, "rigid-body kernels failed");
        c->launches += c->rigid.nlists ? 2 : 0;
    }
    if (stages & 2) c->pending_gate = nullptr;
    CK(cudaGetLastError());
    mark(c, PH_REDUCE);
    return CLSN_OK;
}

extern "C" int clsn_apply(clsn_ctx* c, int rigidify)
{
    if (!c || !c->V) return CLSN_E_ARG;
    cudaSetDevice(c->device);
    return apply_impl(c, rigidify, 3);
}

extern "C" int clsn_apply_stage(clsn_ctx* c, int stage, int rigidify)
{
    if (!c || !c->V || (stage != 1 && stage != 2)) return CLSN_E_ARG;
    cudaSetDevice(c->device);
    int r = apply_impl(c, rigidify, stage);
    if (r) return r;
    // the caller exchanges state next: on its own stream unless it shares one with us (clsn_set_stream)
    if (c->stream == c->own_stream) CK(cudaStreamSynchronize(c->stream));
    return CLSN_OK;
}

extern "C" int clsn_boundary(clsn_ctx* c)
{
    if (!c || !c->V) return CLSN_E_ARG;
    cudaSetDevice(c->device);
    const clsn_params& p = c->prm;
    k_boundary<<<nblk(c->V, 256), 256, 0, c->stream>>>(c->V, c->vflags.p, c->xo.p, c->av.p, c->has.p, p.dt, p.lambda, p.lo[0],
                                                         p.lo[1], p.lo[2], p.hi[0], p.hi[1], p.hi[2]);
    CK(cudaGetLastError());
    c->launches += 1;
    c->dirty_valid = false;   // the clamp rewrites avgVel outside the impulse reduction: no pair may be skipped as untouched
    return CLSN_OK;
}

extern "C" int clsn_final_position(clsn_ctx* c)
{
    if (!c || !c->V) return CLSN_E_ARG;
    cudaSetDevice(c->device);
    k_final_position<<<nblk(c->V, 256), 256, 0, c->stream>>>(c->V, c->xo.p, c->av.p, c->xf.p, c->prm.dt);
    c->final_valid = true;
    CK(cudaGetLastError());
    c->launches += 1;
    mark(c, PH_FINAL);
    return CLSN_OK;
}

// ------------------------------------------------------------------ impact zones
// computeImpactZone, dcollid.cpp:227-265: the fail-safe the reference enters when MAX_ITER CCD passes
// leave collisions.  Every iteration is one full CCD pass + updateAverageVelocity on the GPU, plus
//   * createImpZone (dcollid.cpp:473-484) for every feature test of a pair from its first hit on
//     (`status` is sticky inside Moving*To*, dcollid3d.cpp:203-325): the hits come back as contact
//     records, the host replays the merges in canonical order (pairs by (a,b), features in loop
//     order) on the reference's union-find, whose list order is part of the result;
//   * updateImpactZoneVelocity (dcollid.cpp:290-309): every set of more than one point -- zones and
//     movable bodies alike -- is made to move rigidly by the same kernels as rigid.cuh's bodies.
// Host work is O(contacts) per iteration (+ one O(elements) list flattening when a merge happened).
static int pair_features(int T, int ea, int eb) { return ea < T && eb < T ? 15 : (ea < T ? 5 : 1); }

static void feature_points(const clsn_ctx* c, int ea, int eb, int f, int p[4])
{
    const int T = c->T;
    const int* A = ea < T ? &c->h_tri[3 * (size_t)ea] : &c->h_bond[2 * (size_t)(ea - T)];
    const int* B = eb < T ? &c->h_tri[3 * (size_t)eb] : &c->h_bond[2 * (size_t)(eb - T)];
    if (ea < T && eb < T) {  // MovingTriToTri, dcollid3d.cpp:274-325
        if (f < 3) { p[0] = A[0]; p[1] = A[1]; p[2] = A[2]; p[3] = B[f]; }
        else if (f < 6) { p[0] = B[0]; p[1] = B[1]; p[2] = B[2]; p[3] = A[f - 3]; }
        else {
            const int i = (f - 6) / 3, j = (f - 6) % 3;
            p[0] = A[i]; p[1] = A[(i + 1) % 3]; p[2] = B[j]; p[3] = B[(j + 1) % 3];
        }
    } else if (ea < T) {     // MovingTriToBond, :203-244 (triangles precede bonds in hseList)
        if (f < 2) { p[0] = A[0]; p[1] = A[1]; p[2] = A[2]; p[3] = B[f]; }
        else { const int i = f - 2; p[0] = A[i]; p[1] = A[(i + 1) % 3]; p[2] = B[0]; p[3] = B[1]; }
    } else {                 // MovingBondToBond, :246-272
        p[0] = A[0]; p[1] = A[1]; p[2] = B[0]; p[3] = B[1];
    }
}

// all sets of more than one point, in order of first appearance in hseList (updateImpactZoneVelocity)
static int upload_zone_lists(clsn_ctx* c)
{
    HostUF& uf = c->zone_uf;
    std::vector<int> offs(1, 0), pts, pt_list;
    std::vector<uint8_t> seen((size_t)c->V, 0);
    auto visit = [&](int p) {
        const int r = uf.find(p);
        if (seen[r] || uf.weight[r] == 1) return;
        seen[r] = 1;
        for (int q = r; q >= 0; q = uf.next[q]) {
            pts.push_back(q);
            pt_list.push_back((int)offs.size() - 1);
        }
        offs.push_back((int)pts.size());
    };
    for (size_t i = 0; i < c->h_tri.size(); ++i) visit(c->h_tri[i]);
    for (size_t i = 0; i < c->h_bond.size(); ++i) visit(c->h_bond[i]);
    if (c->zone_lists.upload(offs, pts, pt_list) != 0) return fail(c, CLSN_E_NOMEM, "impact-zone list upload failed");
    c->zone_lists_stale = false;
    return CLSN_OK;
}

extern "C" int clsn_compute_impact_zone(clsn_ctx* c, int max_iter, clsn_zone_stats* out)
{
    if (!c || !c->V) return CLSN_E_ARG;
    cudaSetDevice(c->device);
    if (c->nranks != 1) return fail(c, CLSN_E_UNSUPPORTED, "impact zones need the whole pair set (not a sliced context)");
    if (max_iter <= 0) max_iter = 100000;
    clsn_zone_stats zs;
    memset(&zs, 0, sizeof(zs));
    if (!c->zone_uf_ready) {  // makeSet + createImpZoneForRG, once per step (dcollid3d.cpp:44-45)
        c->zone_uf.reset(c->V);
        c->zone_uf.merge_movable_bodies(c->T, c->h_tri.data(), c->h_tri_surf.data(), c->h_vflags.data());
        c->zone_uf_ready = true;
        c->zone_lists_stale = true;
    }
    const bool dbg_saved = c->dbg_contacts;
    if (c->contacts.n == 0) CK(c->contacts.reserve((size_t)c->N / 8 + 4096));
    c->dbg_contacts = true;
    std::vector<Contact> con;
    std::vector<unsigned long long> order;
    bool is_collision = true;
    int r = CLSN_OK;
    while (is_collision) {
        clsn_pass_stats st;
        if ((r = run_detect(c, CLSN_COLLISION, &st))) break;
        is_collision = st.true_pairs > 0;
        zs.true_pairs += st.true_pairs;
        const size_t n = (size_t)c->n_contacts;
        con.resize(n);
        if (n) {
            cudaError_t e = cudaMemcpyAsync(con.data(), c->contacts.p, n * sizeof(Contact), cudaMemcpyDeviceToHost, c->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
            if (e != cudaSuccess) { c->err = cudaGetErrorString(e); r = CLSN_E_CUDA; break; }
        }
        // first hit of every true pair, pairs in canonical order
        std::sort(con.begin(), con.end(), [](const Contact& a, const Contact& b) {
            return a.ea != b.ea ? a.ea < b.ea : (a.eb != b.eb ? a.eb < b.eb : a.feature < b.feature);
        });
        for (size_t i = 0; i < n; ++i) {
            if (i && con[i].ea == con[i - 1].ea && con[i].eb == con[i - 1].eb) continue;
            const int ea = con[i].ea, eb = con[i].eb, nf = pair_features(c->T, ea, eb);
            for (int f = con[i].feature; f < nf; ++f) {
                int p[4];
                feature_points(c, ea, eb, f, p);
                for (int a = 0; a < 4; ++a)      // createImpZone(pts, 4, first = NO)
                    for (int b = 0; b < a; ++b) {
                        if ((c->h_vflags[p[a]] & 2) || (c->h_vflags[p[b]] & 2)) continue;
                        if (c->zone_uf.merge(p[a], p[b])) { c->zone_lists_stale = true; ++zs.merges; }
                    }
            }
        }
        if ((r = clsn_apply(c, 1))) break;
        if (c->zone_lists_stale && (r = upload_zone_lists(c))) break;
        if (c->prm.dt > 0.0 && c->zone_lists.nlists) {
            if (c->zone_lists.rigidify(c->xo.p, c->av.p, c->vflags.p, c->prm.m, c->prm.dt, c->ctr, c->stream, c->dirty.p) != 0) {
                r = fail(c, CLSN_E_CUDA, "impact-zone kernels failed");
                break;
            }
            c->launches += 2;
        }
        mark(c, PH_REDUCE);
        zs.zones = c->zone_lists.nlists;
        zs.zone_points = c->zone_lists.npts;
        if (++zs.iterations >= max_iter && is_collision) {
            r = fail(c, CLSN_E_NUMERIC, "impact zones did not converge within max_iter iterations");
            break;
        }
    }
    c->dbg_contacts = dbg_saved;
    zs.converged = (r == CLSN_OK && !is_collision) ? 1 : 0;
    if (out) *out = zs;
    return r;
}

extern "C" int clsn_set_impact_zones(clsn_ctx* c, int on, int max_iter)
{
    if (!c) return CLSN_E_ARG;
    c->impact_zones = on != 0;
    c->zone_max_iter = max_iter;
    return CLSN_OK;
}

// ------------------------------------------------------------------ strain limiting
extern "C" int clsn_set_rest_lengths(clsn_ctx* c, const double* tri_len0, const double* bond_len0)
{
    if (!c || !c->V || (c->T > 0 && !tri_len0) || (c->B > 0 && !bond_len0)) return CLSN_E_ARG;
    c->strain.h_tri_len0.assign(tri_len0, tri_len0 + 3 * (size_t)c->T);
    c->strain.h_bond_len0.assign(bond_len0, bond_len0 + c->B);
    c->strain.have_len0 = true;
    c->strain.release_schedule();
    return CLSN_OK;
}

extern "C" int clsn_set_strain_limiting(clsn_ctx* c, int on)
{
    if (!c) return CLSN_E_ARG;
    c->strain_limiting = on != 0;
    return CLSN_OK;
}

// enqueue reduceSuperelast on the resident avgVel; strain_finish() reads the outcome after a stream sync
static int strain_enqueue(clsn_ctx* c)
{
    if (!c->strain.have_len0) return fail(c, CLSN_E_ARG, "strain limiting needs clsn_set_rest_lengths (TRI::side_length0 / BOND::length0)");
    if (!c->strain.built) {
        cudaSetDevice(c->device);
        int r = c->strain.build(c->V, c->T, c->B, c->h_tri.data(), c->h_bond.data(), c->h_vflags.data());
        if (r == -2) return fail(c, CLSN_E_ARG, "too many edges for the 32-bit strain-limiting schedule");
        if (r != 0) return fail(c, CLSN_E_NOMEM, "strain-limiting schedule upload failed");
    }
    if (c->strain.run(c->xo.p, c->av.p, c->prm.dt, c->sm_count, c->stream, &c->launches) != 0)
        return fail(c, CLSN_E_CUDA, "strain-limiting kernels failed to launch");
    c->strain_pending = true;
    c->dirty_valid = false;
    mark(c, PH_FINAL);
    return CLSN_OK;
}

static void strain_finish(clsn_ctx* c, int32_t* sweeps, int32_t* edges_last)
{
    const StrainResult& r = *c->strain.h_res;
    int n = c->strain.M > 0 ? r.sweeps : 1;
    if (n < 1) n = 1;
    if (sweeps) *sweeps = n;
    if (edges_last) *edges_last = c->strain.M > 0 && r.any ? r.viol[n - 1] : 0;
    c->strain_pending = false;
}

extern "C" int clsn_strain_limit(clsn_ctx* c, int32_t* sweeps, int32_t* edges_last)
{
    if (!c || !c->V) return CLSN_E_ARG;
    cudaSetDevice(c->device);
    int r = strain_enqueue(c);
    if (r) return r;
    CK(cudaStreamSynchronize(c->stream));
    strain_finish(c, sweeps, edges_last);
    return CLSN_OK;
}

// resolveCollision, dcollid.cpp:317-362 (detectProximity :390-406, detectCollision :430-468).
// The whole step is enqueued without a single read-back: the proximity pass, the five CCD passes (each gated on the
// device by the contact count of the one before it -- `while (is_collision)`), the applies, the wall clamp and the final
// positions, optionally the copies into the caller's host arrays.  ONE cudaStreamSynchronize at the end brings back the
// counter blocks of all passes; only if a work list turned out too small (first step of a new scene, or a step much more
// violent than the one before) the lists are grown and the step is enqueued again from the uploaded state, which the
// step never overwrites (final positions go to their own buffer).  With the impact-zone fail-safe enabled there is one
// more synchronisation after the CCD passes, because entering it is a host decision (and the fail-safe itself is
// host-assisted).
struct HostOut {
    double* x = nullptr;
    double* avgvel = nullptr;
    uint8_t* has = nullptr;
};

static int read_blocks(clsn_ctx* c)
{
    CK(cudaMemcpyAsync(c->h_counters, pass_block(c, 0), PASS_SLOTS * CTR_STRIDE * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                       c->stream));
    CK(cudaMemcpyAsync(c->h_counters + PASS_SLOTS * CTR_STRIDE, c->tree_scratch.p + 12 * (size_t)c->tree.cnt[3], 6 * sizeof(float),
                       cudaMemcpyDeviceToHost, c->stream));
    if (c->dist.on) {   // the all-reduced blocks (global counts) and every rank's maxima
        CK(cudaMemcpyAsync(c->dist.h_gsum, c->dist.gsum.p, PASS_SLOTS * CTR_STRIDE * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                           c->stream));
        CK(cudaMemcpyAsync(c->dist.h_gsum + PASS_SLOTS * CTR_STRIDE, c->dist.allmax.p,
                           (size_t)PASS_SLOTS * CLSN_MAX_RANKS * 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    }
    CK(cudaStreamSynchronize(c->stream));
    return CLSN_OK;
}

// host copy of a pass's counters as the step's statistics see them: the all-reduced block when the step is distributed
static inline const unsigned long long* stat_block(clsn_ctx* c, int slot)
{
    return c->dist.on ? c->dist.h_gsum + CTR_STRIDE * slot : host_block(c, slot);
}

// which CCD passes ran: pass 1 always, pass k + 1 iff pass k hit something
static int ccd_passes_run(clsn_ctx* c)
{
    int n = 1;
    while (n < CLSN_MAX_CCD_PASSES && stat_block(c, n)[CTR_CONTACTS] > 0) ++n;
    return n;
}

// multi-GPU: after the step's read-back, grow what the all-gathered maxima ask for (every rank sees the same numbers, so
// the collective re-allocation is entered by all of them or by none).  Returns 1 when the step has to be repeated.
static int dist_grow(clsn_ctx* c, int first_slot, int last_slot)
{
    clsn_ctx::Dist& d = c->dist;
    unsigned long long need_region = 0, need_brec = 0, ovf = 0;
    const unsigned long long* am = d.h_gsum + PASS_SLOTS * CTR_STRIDE;
    for (int p = first_slot; p <= last_slot; ++p) {
        ovf |= (d.h_gsum + CTR_STRIDE * p)[CTR_OVF];
        for (int r = 0; r < c->nranks; ++r) {
            need_region = std::max(need_region, am[((size_t)p * CLSN_MAX_RANKS + r) * 4]);
            need_brec = std::max(need_brec, am[((size_t)p * CLSN_MAX_RANKS + r) * 4 + 1]);
        }
    }
    if (!ovf) return 0;
    if (need_brec > (unsigned long long)d.cap_brec_x) {
        d.cap_brec_x = (long long)(need_brec * 5 / 4 + 1024);
        CK(c->brec.reserve((size_t)d.cap_brec_x));
        CK(d.brec_all.reserve((size_t)c->nranks * (size_t)d.cap_brec_x));
        CK(d.brec_dense.reserve((size_t)c->nranks * (size_t)d.cap_brec_x));
    }
    if (need_region > (unsigned long long)d.cap_region) {
        int r = dist_alloc_regions(c, (long long)(need_region * 5 / 4 + 1024));
        if (r) return r;
    }
    return 1;
}

static int resolve_impl(clsn_ctx* c, clsn_step_stats& s, const HostOut& out)
{
    int r;
    const size_t nb = 3 * (size_t)c->nbody;
    CK(c->imp_rg_keep.reserve(nb));
    CK(cudaMemcpyAsync(c->imp_rg_keep.p, c->imp_rg.p, nb * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    for (int attempt = 0; attempt < 8; ++attempt) {
        if (attempt) {   // back to the uploaded state
            CK(cudaMemcpyAsync(c->imp_rg.p, c->imp_rg_keep.p, nb * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
            if ((r = begin_step(c))) return r;
            c->n_marks = 0;
        }
        CK(cudaEventRecord(c->ev[0], c->stream));
        mark(c, PH_OTHER);
        CK(cudaMemsetAsync(pass_block(c, 0), 0, PASS_SLOTS * CTR_STRIDE * sizeof(unsigned long long), c->stream));
        if ((r = clsn_avg_velocity(c))) return r;
        if ((r = enqueue_detect(c, CLSN_PROXIMITY, 0, nullptr))) return r;
        if ((r = apply_impl(c, 1, 3))) return r;
        for (int p = 1; p <= CLSN_MAX_CCD_PASSES; ++p) {
            const unsigned long long* gate = p == 1 ? nullptr : (c->dist.on ? gsum_block(c, p - 1) : pass_block(c, p - 1)) + CTR_CONTACTS;
            if ((r = enqueue_detect(c, CLSN_COLLISION, p, gate))) return r;
            if ((r = apply_impl(c, 1, 3))) return r;
        }
        bool checked = false, redo = false;
        auto check = [&]() -> int {
            for (int p = 0; p <= CLSN_MAX_CCD_PASSES; ++p) {
                const int g = grow_after_pass(c, host_block(c, p), p > 0);   // this rank's own work lists
                if (g < 0) return g;
                if (g == 1 && !c->dist.on) redo = true;
            }
            if (c->dist.on) {   // the decision to repeat is global (CTR_OVF of the all-reduced blocks)
                const int g = dist_grow(c, 0, CLSN_MAX_CCD_PASSES);
                if (g < 0) return g;
                redo = g == 1;
            }
            checked = true;
            return CLSN_OK;
        };
        s.zone_iterations = 0;
        s.zones = 0;
        if (c->impact_zones) {  // detectCollision's tail, dcollid.cpp:464-467: a host decision
            if ((r = read_blocks(c))) return r;
            if ((r = check())) return r;
            if (redo) continue;
            const int np = ccd_passes_run(c);
            if (np == CLSN_MAX_CCD_PASSES && stat_block(c, np)[CTR_TRUE] > 0) {
                if (c->dist.on) return fail(c, CLSN_E_UNSUPPORTED, "the impact-zone fail-safe is not available in the multi-GPU step");
                clsn_zone_stats zs;
                if ((r = clsn_compute_impact_zone(c, c->zone_max_iter, &zs))) return r;
                s.zone_iterations = zs.iterations;
                s.zones = zs.zones;
            }
        }
        c->ctr = pass_block(c, PASS_SLOT_STEP);
        if ((r = clsn_boundary(c))) return r;
        if ((r = clsn_final_position(c))) return r;
        if (c->strain_limiting && (r = strain_enqueue(c))) return r;  // reduceSuperelast, dcollid.cpp:355
        CK(cudaEventRecord(c->ev[1], c->stream));
        if (out.x || out.avgvel || out.has) {
            const size_t n = 3 * (size_t)c->V;
            if ((r = clsn_download_state_device(c, out.x ? c->stage.p : nullptr, out.avgvel ? c->stage.p + n : nullptr))) return r;
            if (out.x) CK(cudaMemcpyAsync(out.x, c->stage.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            if (out.avgvel) CK(cudaMemcpyAsync(out.avgvel, c->stage.p + n, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            if (out.has) CK(cudaMemcpyAsync(out.has, c->has.p, c->V, cudaMemcpyDeviceToHost, c->stream));
        }
        if ((r = read_blocks(c))) return r;
        if (!checked && (r = check())) return r;
        if (redo) continue;
        // ---- the step stands.  Keep the tree (leaf order) for the next step unless the scene box has grown by more
        // than 20 % since it was built -- the reference's rule for its own tree (aabbCollision, dcollid.cpp:417-426) --
        // or the order is 64 steps old (the Morton order of a moving mesh ages even at constant volume).
        {
            const float* rb = reinterpret_cast<const float*>(c->h_counters + PASS_SLOTS * CTR_STRIDE);
            const double vol = (double)(rb[3] - rb[0]) * (double)(rb[4] - rb[1]) * (double)(rb[5] - rb[2]);
            if (c->tree_volume < 0.0) c->tree_volume = vol;
            if ((vol - c->tree_volume) > 0.2 * c->tree_volume) c->tree_built = false;
        }
        // statistics from the counter blocks
        unsigned long long err = 0;
        for (int p = 0; p < PASS_SLOTS; ++p) err |= host_block(c, p)[CTR_ERROR] | (c->dist.on ? stat_block(c, p)[CTR_ERROR] : 0ull);
        if (err) return fail(c, CLSN_E_NUMERIC, "NaN/Inf in the collision step (reference: clean_up(ERROR))");
        fill_pass_stats(c, stat_block(c, 0), false, &s.proximity);
        const int np = ccd_passes_run(c);
        for (int p = 1; p <= np; ++p) fill_pass_stats(c, stat_block(c, p), true, &s.ccd[p - 1]);
        s.n_ccd_passes = np;
        s.has_collision = stat_block(c, 1)[CTR_TRUE] > 0 ? 1 : 0;
        s.still_colliding = stat_block(c, np)[CTR_TRUE] > 0 ? 1 : 0;
        CK(cudaEventElapsedTime(&s.ms_total, c->ev[0], c->ev[1]));
        if (c->strain_pending) strain_finish(c, &s.strain_sweeps, &s.strain_edges);
        const bool trace = getenv("CLSN_TRACE") != nullptr && c->rank == 0;   // per-mark timeline on stderr (tuning runs)
        static const char* names[PH_COUNT] = {"avgvel", "build", "refit", "traverse", "cull", "roots", "contact", "reduce", "final", "other"};
        for (size_t i = 1; i < c->n_marks; ++i) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, c->marks[i - 1], c->marks[i]) == cudaSuccess) s.ms_phase[c->mark_phase[i]] += ms;
            if (trace) fprintf(stderr, "%s%s=%.3f", c->mark_phase[i] == PH_REFIT ? "\n  " : " ", names[c->mark_phase[i]], ms);
        }
        if (trace) fprintf(stderr, "\n");
        return CLSN_OK;
    }
    return fail(c, CLSN_E_NOMEM, "pair/record buffers kept overflowing");
}

static int resolve_io(clsn_ctx* c, clsn_step_stats* stats, const HostOut& out)
{
    clsn_step_stats s;
    memset(&s, 0, sizeof(s));
    c->timing = c->phase_timing;   // phase marks are only recorded inside a whole step
    c->n_marks = 0;
    const int r = resolve_impl(c, s, out);
    c->timing = false;  // also on the error paths: a failed step must not leave the marks armed
    c->strain_pending = false;
    if (r == CLSN_OK && stats) *stats = s;
    return r;
}

extern "C" int clsn_resolve(clsn_ctx* c, clsn_step_stats* stats)
{
    if (!c || !c->V) return CLSN_E_ARG;
    cudaSetDevice(c->device);
    return resolve_io(c, stats, HostOut());
}

// updateFinalVelocity, dcollid.cpp:598-624: vel = avgVel where has_collsn.  A 24-byte copy per point out of freshly
// DMA-written (cache-cold) memory: on a 504 K-vertex mesh one host thread needs longer for it than PCIe needs for the
// arrays themselves, so large meshes are split over a few threads (disjoint vertex ranges).
static void merge_final_velocity(int V, const double* av, const uint8_t* has, double* vel)
{
    auto run = [=](int lo, int hi) {
        for (int p = lo; p < hi; ++p)
            if (has[p])
                for (int j = 0; j < 3; ++j) vel[3 * (size_t)p + j] = av[3 * (size_t)p + j];
    };
    const unsigned hw = std::thread::hardware_concurrency();
    const int nt = V >= (1 << 16) ? (int)std::min(4u, std::max(1u, hw / 2)) : 1;
    std::vector<std::thread> th;
    int done = 0;   // vertices handed to helper threads so far
    try {
        for (int t = 1; t < nt; ++t) {
            const int lo = (int)((long long)V * (t - 1) / nt), hi = (int)((long long)V * t / nt);
            th.emplace_back(run, lo, hi);
            done = hi;
        }
    } catch (...) {   // no more threads to be had: the calling thread does the rest
    }
    run(done, V);
    for (auto& x : th) x.join();
}

extern "C" int clsn_step_host(clsn_ctx* c, const double* x_old, const double* x_new, double* x_out, double* vel_inout,
                              uint8_t* has_out, clsn_step_stats* stats)
{
    if (!c || !c->V || !x_old || !x_new || !x_out) return CLSN_E_ARG;
    int r;
    if ((r = clsn_upload_state(c, x_old, x_new))) return r;
    const size_t n = 3 * (size_t)c->V;
    // scratch for avgVel / has_collsn lives in the context's pinned buffer (3V doubles + V bytes)
    HostOut out;
    out.x = x_out;
    out.avgvel = vel_inout ? c->h_pin : nullptr;   // avgVel only travels back when the caller wants velocities updated
    out.has = has_out ? has_out : (vel_inout ? reinterpret_cast<uint8_t*>(c->h_pin + n) : nullptr);
    if ((r = resolve_io(c, stats, out))) return r;   // upload, step and download: one synchronisation
    if (vel_inout) merge_final_velocity(c->V, out.avgvel, out.has, vel_inout);
    return CLSN_OK;
}

// The same call with the raw per-point results instead of the merged velocity: x_out = final Coords, avgvel_out = avgVel
// of EVERY point, has_out = has_collsn (any of the three may be NULL).  For host mirrors that keep STATE::avgVel of all
// points like the reference does (collision_b200/host/collid_b200.cpp): upload, step and download with one synchronisation.
extern "C" int clsn_step_host_state(clsn_ctx* c, const double* x_old, const double* x_new, double* x_out, double* avgvel_out,
                                    uint8_t* has_out, clsn_step_stats* stats)
{
    if (!c || !c->V || !x_old || !x_new) return CLSN_E_ARG;
    int r;
    if ((r = clsn_upload_state(c, x_old, x_new))) return r;
    HostOut out;
    out.x = x_out;
    out.avgvel = avgvel_out;
    out.has = has_out;
    return resolve_io(c, stats, out);
}

// Page-locked host memory for callers that do not link the CUDA runtime themselves: host arrays handed to
// clsn_step_host* / clsn_upload_state / clsn_download_state from such a buffer move by true asynchronous DMA (pageable
// arrays are staged by the driver).  Portable: usable with the contexts of every device of the process.
extern "C" int clsn_host_alloc(void** out, size_t bytes)
{
    if (!out) return CLSN_E_ARG;
    *out = nullptr;
    if (bytes == 0) return CLSN_OK;
    const cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        *out = nullptr;
        cudaGetLastError();
        return e == cudaErrorMemoryAllocation ? CLSN_E_NOMEM : CLSN_E_CUDA;
    }
    return CLSN_OK;
}
extern "C" void clsn_host_free(void* p)
{
    if (p) cudaFreeHost(p);
}

// ------------------------------------------------------------------ updateFinalForRG
__global__ void k_gather_rg(int n, const int* __restrict__ pts, const Vec4* __restrict__ av, const uint8_t* __restrict__ has,
                            double* __restrict__ out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int p = pts[t];
    const Vec4 v = av[p];
    out[4 * (size_t)t] = v.x; out[4 * (size_t)t + 1] = v.y; out[4 * (size_t)t + 2] = v.z;
    out[4 * (size_t)t + 3] = has[p] ? 1.0 : 0.0;
}

// dcollid.cpp:626-675, run by updateFinalVelocity at the end of resolveCollision.  Per movable body, with its points in
// the order the reference's hseList walk meets them: q = the first point, p* = the first point with has_collsn.
//   p* exists:  center_of_mass_velo = avgVel[p*];  center_of_mass = avgVel[p*] * dt + C,  C = mrg_com[body] if p* == q
//               (the value left by the previous call) else the incoming center_of_mass;  mrg_com[body] = new centre
//   otherwise:  mrg_com[body] = incoming center_of_mass
// A body met for the first time reads mrg_com before the reference ever wrote it (an empty std::vector there): seeded with
// the incoming centre of mass.  Only the listed points' avgVel / has_collsn cross PCIe (32 B per movable point).
extern "C" int clsn_update_rigid_bodies(clsn_ctx* c, double* center_of_mass, double* center_of_mass_velo)
{
    if (!c || !c->V || !center_of_mass || !center_of_mass_velo) return CLSN_E_ARG;
    const size_t M = c->rg_pts.size();
    if (M == 0) return CLSN_OK;
    cudaSetDevice(c->device);
    k_gather_rg<<<nblk((long long)M, 256), 256, 0, c->stream>>>((int)M, c->d_rg_pts.p, c->av.p, c->has.p, c->d_rg_state.p);
    CK(cudaGetLastError());
    c->launches += 1;
    std::vector<double> st(4 * M);
    CK(cudaMemcpyAsync(st.data(), c->d_rg_state.p, 4 * M * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    const double dt = c->prm.dt;
    for (int b = 0; b < c->nbody; ++b) {
        const int beg = c->rg_offs[b], end = c->rg_offs[b + 1];
        if (beg == end) continue;
        double* com = center_of_mass + 3 * (size_t)b;
        double* mrg = &c->mrg_com[3 * (size_t)b];
        if (!c->mrg_valid[b]) {
            for (int j = 0; j < 3; ++j) mrg[j] = com[j];
            c->mrg_valid[b] = 1;
        }
        int hit = -1;
        for (int t = beg; t < end && hit < 0; ++t)
            if (st[4 * (size_t)t + 3] != 0.0) hit = t;
        if (hit != beg)   // the walk meets a point without a collision first: mrg_com <- the incoming centre of mass
            for (int j = 0; j < 3; ++j) mrg[j] = com[j];
        if (hit >= 0) {
            for (int j = 0; j < 3; ++j) {
                const double v = st[4 * (size_t)hit + j];
                center_of_mass_velo[3 * (size_t)b + j] = v;
                com[j] = v * dt + mrg[j];
            }
            for (int j = 0; j < 3; ++j) mrg[j] = com[j];
        }
    }
    return CLSN_OK;
}

// ------------------------------------------------------------------ measurement helpers
// CUDA events on the library's own stream (torch.cuda.Event only sees torch's current stream)
extern "C" int clsn_timer_start(clsn_ctx* c)
{
    if (!c) return CLSN_E_ARG;
    cudaSetDevice(c->device);
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaEventRecord(c->bracket[0], c->stream));
    return CLSN_OK;
}
extern "C" int clsn_timer_stop(clsn_ctx* c, float* ms)
{
    if (!c || !ms) return CLSN_E_ARG;
    cudaSetDevice(c->device);
    CK(cudaEventRecord(c->bracket[1], c->stream));
    CK(cudaEventSynchronize(c->bracket[1]));
    CK(cudaEventElapsedTime(ms, c->bracket[0], c->bracket[1]));
    return CLSN_OK;
}
extern "C" int64_t clsn_launch_count(clsn_ctx* c, int reset)
{
    if (!c) return 0;
    long long n = c->launches;
    if (reset) c->launches = 0;
    return n;
}
extern "C" int clsn_synchronize(clsn_ctx* c)
{
    if (!c) return CLSN_E_ARG;
    cudaSetDevice(c->device);
    CK(cudaStreamSynchronize(c->stream));
    return CLSN_OK;
}

// ------------------------------------------------------------------ multi-GPU record exchange
extern "C" int clsn_export_records(clsn_ctx* c, void** d_prec, int64_t* n_prec, void** d_brec, int64_t* n_brec, int64_t* true_pairs)
{
    if (!c || !c->records_pending) return CLSN_E_ARG;
    if (d_prec) *d_prec = c->prec.p;
    if (n_prec) *n_prec = c->last_nprec;
    if (d_brec) *d_brec = c->brec.p;
    if (n_brec) *n_brec = c->last_nbrec;
    if (true_pairs) *true_pairs = (int64_t)c->last_true;
    return CLSN_OK;
}

extern "C" int clsn_bucket_records(clsn_ctx* c, int nranks, int64_t* counts, void** d_sorted)
{
    if (!c || !c->records_pending || nranks < 1 || nranks > 64 || !counts || !d_sorted) return CLSN_E_ARG;
    cudaSetDevice(c->device);
    const long long n = c->last_nprec;
    const int per_rank = (c->V + nranks - 1) / nranks;
    CK(c->prec_sorted.reserve((size_t)(n > 0 ? n : 1)));
    unsigned long long* cnt = c->counters.p + 40;  // 64 bucket counters, then 64 cursors
    CK(cudaMemsetAsync(cnt, 0, 128 * sizeof(unsigned long long), c->stream));
    unsigned long long h[64] = {0};
    if (n > 0) {
        k_owner_count<<<c->sm_count * 4, 256, 0, c->stream>>>(c->prec.p, n, per_rank, cnt);
        CK(cudaMemcpyAsync(h, cnt, 64 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        unsigned long long cur[64], acc = 0;
        for (int r = 0; r < 64; ++r) { cur[r] = acc; acc += h[r]; }
        CK(cudaMemcpyAsync(cnt + 64, cur, 64 * sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));
        k_owner_scatter<<<c->sm_count * 4, 256, 0, c->stream>>>(c->prec.p, n, per_rank, cnt + 64, c->prec_sorted.p);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(c->stream));
        c->launches += 2;
    }
    for (int r = 0; r < nranks; ++r) counts[r] = (int64_t)h[r];
    *d_sorted = c->prec_sorted.p;
    return CLSN_OK;
}

extern "C" int clsn_state_device_ptrs(clsn_ctx* c, void** av, void** has, void** dirty)
{
    if (!c || !c->V) return CLSN_E_ARG;
    if (av) *av = c->av.p;
    if (has) *has = c->has.p;
    if (dirty) *dirty = c->dirty.p;
    return CLSN_OK;
}

extern "C" int clsn_import_records(clsn_ctx* c, const void* d_prec, int64_t n_prec, const void* d_brec, int64_t n_brec)
{
    if (!c || n_prec < 0 || n_brec < 0) return CLSN_E_ARG;
    c->imp_prec = (const PointRec*)d_prec;
    c->imp_brec = (const BodyRec*)d_brec;
    c->imp_nprec = n_prec;
    c->imp_nbrec = n_brec;
    c->records_pending = true;
    return CLSN_OK;
}

// ------------------------------------------------------------------ debug readbacks
extern "C" int64_t clsn_num_candidates(clsn_ctx* c) { return c ? c->n_dbg_cand : 0; }
extern "C" int clsn_get_candidates(clsn_ctx* c, int32_t* pairs)
{
    if (!c || !pairs || !c->dbg_candidates) return CLSN_E_ARG;
    cudaSetDevice(c->device);
    CK(cudaMemcpy(pairs, c->dbg_cand.p, (size_t)c->n_dbg_cand * sizeof(int2), cudaMemcpyDeviceToHost));
    return CLSN_OK;
}
extern "C" int64_t clsn_num_contacts(clsn_ctx* c) { return c ? c->n_contacts : 0; }
extern "C" int clsn_get_contacts(clsn_ctx* c, clsn_contact* out)
{
    if (!c || !out || !c->dbg_contacts) return CLSN_E_ARG;
    cudaSetDevice(c->device);
    static_assert(sizeof(clsn_contact) == sizeof(Contact), "contact layout");
    CK(cudaMemcpy(out, c->contacts.p, (size_t)c->n_contacts * sizeof(Contact), cudaMemcpyDeviceToHost));
    return CLSN_OK;
}

extern "C" int clsn_get_accumulators(clsn_ctx* c, double* imp, double* fric, int32_t* cnt, double* imp_rg, int32_t* cnt_rg)
{
    if (!c || !c->V || !c->records_pending) return CLSN_E_ARG;
    cudaSetDevice(c->device);
    int r = reduce_records(c, 1);
    if (r) return r;
    CK(cudaStreamSynchronize(c->stream));
    const size_t n = 3 * (size_t)c->V;
    if (imp) CK(cudaMemcpy(imp, c->acc_imp.p, n * sizeof(double), cudaMemcpyDeviceToHost));
    if (fric) CK(cudaMemcpy(fric, c->acc_fric.p, n * sizeof(double), cudaMemcpyDeviceToHost));
    if (cnt) CK(cudaMemcpy(cnt, c->cnt.p, (size_t)c->V * sizeof(int), cudaMemcpyDeviceToHost));
    if (imp_rg || cnt_rg) {
        // body sums as they will be after this pass: reduce into a scratch copy
        DevBuf<double> tmp;
        CK(tmp.reserve(3 * (size_t)c->nbody));
        CK(cudaMemcpy(tmp.p, c->imp_rg.p, 3 * (size_t)c->nbody * sizeof(double), cudaMemcpyDeviceToDevice));
        const long long nbrec = c->imp_nprec >= 0 ? c->imp_nbrec : c->last_nbrec;
        if (nbrec > 0) {
            unsigned long long* n_dev = c->imp_nprec >= 0 ? c->counters.p + 33 : c->ctr + CTR_BREC;
            const BodyRec* brec = c->imp_nprec >= 0 ? c->imp_brec : c->brec.p;
            const long long cap = c->imp_nprec >= 0 ? nbrec : (long long)c->brec.n;
            int rr = reduce_bodies(c, brec, n_dev, cap, tmp.p);
            if (rr) return rr;
            CK(cudaStreamSynchronize(c->stream));
        }
        if (imp_rg) CK(cudaMemcpy(imp_rg, tmp.p, 3 * (size_t)c->nbody * sizeof(double), cudaMemcpyDeviceToHost));
        if (cnt_rg) CK(cudaMemcpy(cnt_rg, c->cnt_rg.p, (size_t)c->nbody * sizeof(int), cudaMemcpyDeviceToHost));
        tmp.release();
    }
    return CLSN_OK;
}

extern "C" int clsn_set_body_accumulators(clsn_ctx* c, const double* imp_rg, const int32_t* cnt_rg)
{
    if (!c || !c->V) return CLSN_E_ARG;
    cudaSetDevice(c->device);
    CK(cudaStreamSynchronize(c->stream));
    if (imp_rg) CK(cudaMemcpy(c->imp_rg.p, imp_rg, 3 * (size_t)c->nbody * sizeof(double), cudaMemcpyHostToDevice));
    if (cnt_rg) CK(cudaMemcpy(c->cnt_rg.p, cnt_rg, (size_t)c->nbody * sizeof(int), cudaMemcpyHostToDevice));
    if (cnt_rg) c->cnt_rg_preset = true;
    return CLSN_OK;
}
