// Broad phase: Morton-code LBVH over the element leaf boxes (sm_100a).
//
// Replaces the reference's incremental-insertion AABBTree (AABB.cpp:129-245) and its per-leaf
// stack query (AABB.cpp:254-343).  What must be preserved is only the candidate SET: unordered
// element pairs whose FP64 leaf boxes (CD_TRI/CD_BOND::{min,max}_{static,moving}_coord -+ 1e-6,
// dcollid.cpp:852-932, AABB.cpp:6-36) overlap on closed intervals (AABB.cpp:56-60).  Tree shape
// is free.  Internal nodes carry FP32 boxes rounded outward (conservative); the leaf-leaf test is
// the exact FP64 one.
//
// Layout: leaves are sorted by Morton code; above them sits an implicit 8-ary tree (Node8 / Tree8 below).
#pragma once
#include <stdint.h>

namespace clsn {

struct alignas(32) Vec4 {  // one 32-byte sector per vertex gather
    double x, y, z, w;
};

#define CLSN_BOX_PAD 1e-6 /* AABB.cpp:10-16: fixed, independent of setRoundingTolerance */

__device__ __forceinline__ Vec4 ldg_vec4(const Vec4* p)
{
    const double2* q = reinterpret_cast<const double2*>(p);
    double2 a = __ldg(q), b = __ldg(q + 1);
    return Vec4{a.x, a.y, b.x, b.y};
}

// exact FP64 leaf box of one element.  elem = (p0, p1, p2 | -1, tag)
template <bool MOVING>
__device__ __forceinline__ void leaf_box(const int4 e, const Vec4* __restrict__ xo, const Vec4* __restrict__ av, double dt,
                                         double* lo, double* hi)
{
    const int n = e.z >= 0 ? 3 : 2;
    const int ids[3] = {e.x, e.y, e.z};
    double mn[3] = {1.0e18, 1.0e18, 1.0e18}, mx[3] = {-1.0e18, -1.0e18, -1.0e18};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (i < n) {
            Vec4 p = ldg_vec4(xo + ids[i]);
            const double x0[3] = {p.x, p.y, p.z};
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                mn[d] = fmin(mn[d], x0[d]);
                mx[d] = fmax(mx[d], x0[d]);
            }
            if (MOVING) {
                Vec4 v = ldg_vec4(av + ids[i]);
                const double vv[3] = {v.x, v.y, v.z};
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    double x1 = x0[d] + vv[d] * dt;  // dcollid.cpp:918: x_old + avgVel*dt
                    mn[d] = fmin(mn[d], x1);
                    mx[d] = fmax(mx[d], x1);
                }
            }
        }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        lo[d] = mn[d] - CLSN_BOX_PAD;
        hi[d] = mx[d] + CLSN_BOX_PAD;
    }
}

// order-preserving map double -> uint64 for atomicMin/Max
__device__ __forceinline__ unsigned long long enc_f64(double x)
{
    unsigned long long u = (unsigned long long)__double_as_longlong(x);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double dec_f64(unsigned long long u)
{
    unsigned long long v = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u;
    double d;
#if defined(__CUDA_ARCH__)
    d = __longlong_as_double((long long)v);
#else
    memcpy(&d, &v, 8);
#endif
    return d;
}

// scene bounds over vertex positions (x_old): bounds[0..2] = min, [3..5] = max (encoded)
__global__ void k_scene_bounds(const Vec4* __restrict__ xo, int V, unsigned long long* bounds)
{
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
        Vec4 p = ldg_vec4(xo + v);
        mn[0] = fmin(mn[0], p.x); mn[1] = fmin(mn[1], p.y); mn[2] = fmin(mn[2], p.z);
        mx[0] = fmax(mx[0], p.x); mx[1] = fmax(mx[1], p.y); mx[2] = fmax(mx[2], p.z);
    }
#pragma unroll
    for (int d = 0; d < 3; ++d)
        for (int o = 16; o > 0; o >>= 1) {
            mn[d] = fmin(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
            mx[d] = fmax(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
        }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            atomicMin(&bounds[d], enc_f64(mn[d]));
            atomicMax(&bounds[3 + d], enc_f64(mx[d]));
        }
    }
}

__device__ __forceinline__ unsigned expand10(unsigned v)
{
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

// 30-bit Morton code of the centroid of the element's x_old vertices
__global__ void k_morton(const int4* __restrict__ elem, int N, const Vec4* __restrict__ xo,
                         const unsigned long long* __restrict__ bounds, unsigned* __restrict__ code, int* __restrict__ idx)
{
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N) return;
    int4 el = __ldg(elem + e);
    float lo[3], inv[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        double a = dec_f64(bounds[d]), b = dec_f64(bounds[3 + d]);
        lo[d] = (float)a;
        float ext = (float)(b - a);
        inv[d] = ext > 0.f ? 1023.0f / ext : 0.f;
    }
    Vec4 p0 = ldg_vec4(xo + el.x), p1 = ldg_vec4(xo + el.y);
    float c[3] = {(float)(p0.x + p1.x), (float)(p0.y + p1.y), (float)(p0.z + p1.z)};
    float n = 2.f;
    if (el.z >= 0) {
        Vec4 p2 = ldg_vec4(xo + el.z);
        c[0] += (float)p2.x; c[1] += (float)p2.y; c[2] += (float)p2.z;
        n = 3.f;
    }
    unsigned q[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        float t = (c[d] / n - lo[d]) * inv[d];
        t = fminf(fmaxf(t, 0.f), 1023.f);
        q[d] = (unsigned)t;
    }
    code[e] = (expand10(q[0]) << 2) | (expand10(q[1]) << 1) | expand10(q[2]);
    idx[e] = e;
}

// ------------------------------------------------------------------ implicit 8-ary tree over the Morton order
// Level 0 = the leaves in Morton order; node k of level l >= 1 has the children 8k .. 8k+7 of level l-1 (leaves for
// l = 1), i.e. it covers the sorted leaves [k 8^l, (k+1) 8^l).  No pointers, no build beyond the sort: the shape is a
// function of N alone, a refit is a streaming bottom-up reduction without atomics, and the "each unordered pair once"
// rule (j > i) needs no per-node field because the leaf range of a child is implied by its index.  A node record holds
// the boxes of its 8 children quantised to 8 bits per plane relative to the node's own box (conservatively: quantised
// boxes contain the FP32 boxes, which contain the exact FP64 leaf boxes): 64 bytes per visit instead of 8 x 24, because
// the traversal is bound by L2 bandwidth, not by latency (measured: 8 FP32 boxes in 192 B ran no faster than the binary
// radix tree's 2 boxes in 64 B over three times as many dependent hops).  9 MB at 1 M leaves, 7 levels deep.
struct alignas(64) Node8 {  // 64 B: the boxes of the 8 children, quantised to 8 bits per plane inside the node's own box
    float plo[3];        // lower corner of the union of the children (FP32, rounded down)
    unsigned meta;       // e[0] | e[1] << 8 | e[2] << 16 (int8 exponents: one quantum = 2^e[d]) | touched << 24
                         // touched: bit c = child c holds a point changed by the previous pass (see k_traverse8)
    unsigned q[12];      // q[2p + h] = plane p of children 4h .. 4h+3, one byte each; planes: lo.x lo.y lo.z hi.x hi.y hi.z
                         // child box = plo + q * 2^e, lo planes rounded down, hi planes rounded up: always CONTAINS the
                         // child's FP32 box.  An empty slot has lo = 255, hi = 0.
};
#define TREE_MAXLEV 12
struct Tree8 {
    int nlev;              // internal levels 1 .. nlev; level nlev holds the single root node (nlev >= 3)
    int N;                 // leaves
    int cnt[TREE_MAXLEV];  // cnt[0] = N, cnt[l] = ceil(cnt[l-1] / 8)
    int off[TREE_MAXLEV];  // index of the first node of level l in nodes[] / touched[]
    Node8* nodes;
};
#define REFIT_LEAVES 512  // leaves per block of k_refit8 = 64 level-1 nodes = 8 level-2 nodes = 1 level-3 node

__global__ void k_gather_elems(const int4* __restrict__ elem, const int* __restrict__ leaf_elem, int N, int4* __restrict__ selem)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) selem[i] = __ldg(elem + __ldg(leaf_elem + i));
}

// union of the boxes of the 8 lanes of a group (lane & 7 == child slot); every lane of `mask` ends with the group's box
__device__ __forceinline__ void box_reduce8(float b[6], unsigned mask)
{
#pragma unroll
    for (int o = 1; o < 8; o <<= 1)
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            b[d] = fminf(b[d], __shfl_xor_sync(mask, b[d], o));
            b[3 + d] = fmaxf(b[3 + d], __shfl_xor_sync(mask, b[3 + d], o));
        }
}
__device__ __forceinline__ float pow2f(int e) { return __int_as_float((127 + e) << 23); }  // e in [-126, 127]

// Called by the 8 lanes of a group (lane & 7 = child slot c) with the child's box b and, after box_reduce8, the group's
// union P: quantise b inside P and store the record.  `tm` = touched mask of the group (all lanes), `exists` = the node
// is inside the level (lanes of a non-existing node still take part in the shuffles).  mask: the shuffle mask.
__device__ __forceinline__ void node_store(Node8* nd, int c, const float b[6], const float P[6], unsigned tm, bool exists, unsigned mask)
{
    unsigned ql[3], qh[3], ebits = 0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float ext = __fsub_ru(P[3 + d], P[d]);
        int e = -100;
        if (ext > 0.f && ext <= 3.0e38f) {
            const float t = __fdiv_ru(ext, 255.f);
            e = ((__float_as_int(t) >> 23) & 0xff) - 127 + 1;   // 2^e > t  =>  255 * 2^e > ext
            if (e < -100) e = -100;
        }
        const float inv = pow2f(-e);
        if (b[d] <= b[3 + d]) {
            int lo = __float2int_rd(__fsub_rd(b[d], P[d]) * inv);
            int hi = __float2int_ru(__fsub_ru(b[3 + d], P[d]) * inv);
            ql[d] = (unsigned)max(0, min(255, lo));
            qh[d] = (unsigned)max(0, min(255, hi));
        } else {  // empty slot
            ql[d] = 255u;
            qh[d] = 0u;
        }
        ebits |= ((unsigned)e & 0xffu) << (8 * d);
    }
    // assemble the words of each half (children 0..3 / 4..7): byte (c & 3) of word [2p + (c >> 2)]
    unsigned w[6];
#pragma unroll
    for (int p = 0; p < 6; ++p) {
        unsigned v = (p < 3 ? ql[p] : qh[p - 3]) << (8 * (c & 3));
        v |= __shfl_xor_sync(mask, v, 1);
        v |= __shfl_xor_sync(mask, v, 2);
        w[p] = v;
    }
    if (!exists) return;
    const int j = c & 3, h = c >> 2;
    if (j < 3) {
        nd->q[2 * j + h] = j == 0 ? w[0] : (j == 1 ? w[1] : w[2]);
        nd->q[2 * (j + 3) + h] = j == 0 ? w[3] : (j == 1 ? w[4] : w[5]);
    } else if (h == 0) {
        *reinterpret_cast<float4*>(nd) = make_float4(P[0], P[1], P[2], __uint_as_float(ebits | (tm << 24)));
    }
}

// Exact FP64 leaf boxes (stored in Morton order) + the whole tree above them, one launch.  Block = 512 consecutive
// leaves: leaf boxes -> level-1 records -> (8-lane shuffles, shared memory) level-2 and level-3 records; the block's
// level-3 box goes to a small scratch array and the LAST block to finish (ticket) builds the few levels above from it.
// Traffic per leaf: 16 B element + vertex gathers (L2) in, 48 B exact box + 24 B level-1 slot out.
// scratch: 2 x cnt[3] x 6 floats (ping-pong of the upper levels), scratch_t: 2 x cnt[3] bytes, ticket: one zeroed word,
// root_box: 6 floats (the scene box of this refit; volume rule of dcollid.cpp:377-385)
#ifndef REFIT_MIN_BLOCKS
#define REFIT_MIN_BLOCKS 3
#endif
template <bool MOVING>
__global__ void __launch_bounds__(REFIT_LEAVES, REFIT_MIN_BLOCKS)
k_refit8(const int4* __restrict__ selem, Tree8 tr, const Vec4* __restrict__ xo, const Vec4* __restrict__ av, double dt,
         double* __restrict__ lbox, const uint8_t* __restrict__ vdirty, float* scratch, uint8_t* scratch_t, unsigned* ticket,
         float* root_box, const unsigned long long* __restrict__ gate)
{
    if (gate && *gate == 0ull) return;   // the previous CCD pass found nothing: this pass does not run (dcollid.cpp:448)
    __shared__ float s_box[64][6];
    __shared__ uint8_t s_t[64];
    __shared__ float s_box2[8][6];
    __shared__ uint8_t s_t2[8];
    __shared__ bool s_last;
    const int t = threadIdx.x, lane = t & 31, N = tr.N;
    const int i = blockIdx.x * REFIT_LEAVES + t;
    const float INF = __int_as_float(0x7f800000);
    float b[6] = {INF, INF, INF, -INF, -INF, -INF};
    int dbit = 0;
    if (i < N) {
        const int4 el = __ldg(selem + i);
        double lo[3], hi[3];
        leaf_box<MOVING>(el, xo, av, dt, lo, hi);
        double2* lb = reinterpret_cast<double2*>(lbox + 6 * (size_t)i);
        lb[0] = make_double2(lo[0], lo[1]);
        lb[1] = make_double2(lo[2], hi[0]);
        lb[2] = make_double2(hi[1], hi[2]);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            b[d] = __double2float_rd(lo[d]);
            b[3 + d] = __double2float_ru(hi[d]);
        }
        dbit = 1;
        if (vdirty) {
            dbit = vdirty[el.x] | vdirty[el.y];
            if (el.z >= 0) dbit |= vdirty[el.z];
        }
    }
    // level 1: leaf i is child i & 7 of node i >> 3
    float P[6];
#pragma unroll
    for (int d = 0; d < 6; ++d) P[d] = b[d];
    box_reduce8(P, 0xffffffffu);
    unsigned bal = __ballot_sync(0xffffffffu, dbit != 0);
    unsigned tm = (bal >> (lane & ~7)) & 0xffu;
    const int k1 = i >> 3;
    node_store(tr.nodes + tr.off[1] + k1, i & 7, b, P, tm, k1 < tr.cnt[1], 0xffffffffu);
    if ((t & 7) == 0) {
#pragma unroll
        for (int d = 0; d < 6; ++d) s_box[t >> 3][d] = P[d];
        s_t[t >> 3] = tm != 0;
    }
    __syncthreads();
    // level 2: the block's 64 level-1 nodes
    if (t < 64) {
#pragma unroll
        for (int d = 0; d < 6; ++d) P[d] = b[d] = s_box[t][d];
        const int n1 = blockIdx.x * 64 + t, k2 = n1 >> 3;
        box_reduce8(P, 0xffffffffu);
        bal = __ballot_sync(0xffffffffu, s_t[t] != 0);
        tm = (bal >> (lane & ~7)) & 0xffu;
        node_store(tr.nodes + tr.off[2] + k2, n1 & 7, b, P, tm, k2 < tr.cnt[2], 0xffffffffu);
        if ((t & 7) == 0) {
#pragma unroll
            for (int d = 0; d < 6; ++d) s_box2[t >> 3][d] = P[d];
            s_t2[t >> 3] = tm != 0;
        }
    }
    __syncthreads();
    // level 3: the block's 8 level-2 nodes are the children of node blockIdx.x
    if (t < 8) {
#pragma unroll
        for (int d = 0; d < 6; ++d) P[d] = b[d] = s_box2[t][d];
        box_reduce8(P, 0xffu);
        bal = __ballot_sync(0xffu, s_t2[t] != 0);
        tm = bal & 0xffu;
        node_store(tr.nodes + tr.off[3] + blockIdx.x, t, b, P, tm, true, 0xffu);
        if (t == 0) {
#pragma unroll
            for (int d = 0; d < 6; ++d) scratch[6 * (size_t)blockIdx.x + d] = P[d];
            scratch_t[blockIdx.x] = tm != 0;
            if (tr.nlev == 3) {
#pragma unroll
                for (int d = 0; d < 6; ++d) root_box[d] = P[d];
            }
        }
    }
    if (tr.nlev == 3) return;
    // levels 4 .. nlev: the last block to arrive reduces the level-3 boxes (a few thousand at most)
    // only thread 0's own stores (the block's level-3 box in `scratch`) are read by the last block: one fence, by it
    if (t == 0) {
        __threadfence();
        s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    if (t == 0) *ticket = 0u;   // armed for the next launch
    __threadfence();
    const int n3 = tr.cnt[3];
    float* cur = scratch;
    uint8_t* cur_t = scratch_t;
    float* nxt = scratch + 6 * (size_t)n3;
    uint8_t* nxt_t = scratch_t + n3;
    int ncur = n3;
    for (int l = 4; l <= tr.nlev; ++l) {
        const int nn = tr.cnt[l];
        for (int base = 0; base < nn * 8; base += REFIT_LEAVES) {
            const int ch = base + t;
            float c[6] = {INF, INF, INF, -INF, -INF, -INF};
            int tb = 0;
            if (ch < ncur) {
#pragma unroll
                for (int d = 0; d < 6; ++d) c[d] = __ldcg(cur + 6 * (size_t)ch + d);
                tb = __ldcg(cur_t + ch);
            }
            float Pu[6];
#pragma unroll
            for (int d = 0; d < 6; ++d) Pu[d] = c[d];
            box_reduce8(Pu, 0xffffffffu);
            bal = __ballot_sync(0xffffffffu, tb != 0);
            tm = (bal >> (lane & ~7)) & 0xffu;
            node_store(tr.nodes + tr.off[l] + (ch >> 3), ch & 7, c, Pu, tm, (ch >> 3) < nn, 0xffffffffu);
            if ((t & 7) == 0 && (ch >> 3) < nn) {
#pragma unroll
                for (int d = 0; d < 6; ++d) nxt[6 * (size_t)(ch >> 3) + d] = Pu[d];
                nxt_t[ch >> 3] = tm != 0;
                if (l == tr.nlev) {
#pragma unroll
                    for (int d = 0; d < 6; ++d) root_box[d] = Pu[d];
                }
            }
        }
        __threadfence_block();
        __syncthreads();
        float* tf = cur; cur = nxt; nxt = tf;
        uint8_t* tt = cur_t; cur_t = nxt_t; nxt_t = tt;
        ncur = nn;
    }
}

struct TraverseOut {
    int2* pairs;            // non-adjacent, unfiltered-by-narrow-phase element pairs (a < b)
    long long cap_pairs;
    int2* dbg_cand;         // every overlapping pair (debug / parity), may be null
    long long cap_dbg;
    unsigned long long* counters;
    const uint8_t* dirty;   // non-null: drop pairs none of whose points changed since the previous CCD pass
    int prune;              // (with dirty) an untouched query only descends into touched subtrees (tree.touched) -- its
                            // other pairs would be dropped anyway.  `candidates` then counts only the pairs actually found.
    const unsigned long long* gate;  // non-null: run only if *gate != 0 (the previous CCD pass found a collision)
};

#ifndef TRAV_THREADS
#define TRAV_THREADS 128
#endif
#define TRAV_QCAP 288  // per-warp queue: <= 31 left over + 8 leaf children per lane and round

// Self-query.  Thread = one query leaf (sorted index i in [q_lo, q_hi)); finds leaves j > i whose exact FP64 boxes
// overlap (AABB::isCollid, AABB.cpp:56-60).  Children whose leaf range ends at or before i are skipped, so each
// unordered pair is found once.  One loop round = one node visit per lane (8 conservative FP32 box tests); leaf
// children that pass are pushed to a per-warp shared-memory queue, and whenever it holds >= 32 entries the whole warp
// drains 32 of them together: exact FP64 test, adjacency / same-surface-rigid filters (dcollid3d.cpp:279-284,
// dcollid.cpp:762) and a ballot-aggregated append to the pair list -- all lanes busy instead of one or two.
__global__ void __launch_bounds__(TRAV_THREADS)
k_traverse8(Tree8 tr, const double* __restrict__ lbox, const int* __restrict__ leaf_elem, const int4* __restrict__ selem,
            int q_lo, int q_hi, TraverseOut out)
{
    if (out.gate && *out.gate == 0ull) return;
    __shared__ int2 s_q[TRAV_THREADS / 32][TRAV_QCAP];
    __shared__ int s_n[TRAV_THREADS / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int i = q_lo + blockIdx.x * blockDim.x + threadIdx.x;
    const int N = tr.N;
    unsigned long long n_cand = 0;
    if (lane == 0) s_n[w] = 0;
    __syncwarp();
    float flo[3] = {0, 0, 0}, fhi[3] = {0, 0, 0};
    int cur = -1;  // (level << 27) | node; -1 = this lane has finished
    bool me_touched = true;
    if (i < q_hi && i + 1 < N) {
        if (out.prune) {
            const int4 el = __ldg(selem + i);
            unsigned d = out.dirty[el.x] | out.dirty[el.y];
            if (el.z >= 0) d |= out.dirty[el.z];
            me_touched = d != 0;
        }
        const double2* lb = reinterpret_cast<const double2*>(lbox + 6 * (size_t)i);
        const double2 b0 = __ldg(lb), b1 = __ldg(lb + 1), b2 = __ldg(lb + 2);
        flo[0] = __double2float_rd(b0.x); flo[1] = __double2float_rd(b0.y); flo[2] = __double2float_rd(b1.x);
        fhi[0] = __double2float_ru(b1.y); fhi[1] = __double2float_ru(b2.x); fhi[2] = __double2float_ru(b2.y);
        cur = tr.nlev << 27;
    }
    int stack[64];
    int sp = 0;
    bool more = __any_sync(0xffffffffu, cur >= 0);
    while (more || s_n[w] > 0) {
        if (cur >= 0) {
            const int level = cur >> 27, k = cur & 0x7ffffff;
            const uint4* r = reinterpret_cast<const uint4*>(tr.nodes + tr.off[level] + k);
            const uint4 h = __ldg(r), w0 = __ldg(r + 1), w1 = __ldg(r + 2), w2 = __ldg(r + 3);
            // h = plo.xyz, meta; w0 = lo.x[0..3] lo.x[4..7] lo.y[0..3] lo.y[4..7]; w1 = lo.z.. hi.x..; w2 = hi.y.. hi.z..
            const float plo[3] = {__uint_as_float(h.x), __uint_as_float(h.y), __uint_as_float(h.z)};
            const unsigned qlo[3][2] = {{w0.x, w0.y}, {w0.z, w0.w}, {w1.x, w1.y}};
            const unsigned qhi[3][2] = {{w1.z, w1.w}, {w2.x, w2.y}, {w2.z, w2.w}};
            unsigned m0 = 0xffffffffu, m1 = 0xffffffffu;   // per-byte pass masks of children 0..3 / 4..7
            bool none = false;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                // the query box in the node's quantised frame, rounded so that the test can only err towards "overlap":
                // child.lo <= query.hi  <=  q_lo <= floor((fhi - plo) / 2^e)  (difference rounded up)
                // child.hi >= query.lo  <=  q_hi >= ceil((flo - plo) / 2^e)   (difference rounded down)
                const int e = (int)(signed char)((h.w >> (8 * d)) & 0xffu);
                const float inv = pow2f(-e);
                const float a = __fsub_ru(fhi[d], plo[d]) * inv, bq = __fsub_rd(flo[d], plo[d]) * inv;
                if (a < 0.f || bq > 255.f) none = true;
                const unsigned Qhi = (unsigned)min(255, max(0, __float2int_rd(a)));
                const unsigned Qlo = (unsigned)min(255, max(0, __float2int_ru(bq)));
                const unsigned rh = Qhi * 0x01010101u, rl = Qlo * 0x01010101u;
                m0 &= __vcmpleu4(qlo[d][0], rh) & __vcmpgeu4(qhi[d][0], rl);
                m1 &= __vcmpleu4(qlo[d][1], rh) & __vcmpgeu4(qhi[d][1], rl);
            }
            // byte masks (0xff / 0x00) -> one bit per child
            unsigned mask = none ? 0u : ((((m0 & 0x01010101u) * 0x01020408u) >> 24) & 0xfu) | (((((m1 & 0x01010101u) * 0x01020408u) >> 24) & 0xfu) << 4);
            // slots beyond the end of the level below (an empty slot can pass the quantised test of a query that spans the node)
            const int nvalid = tr.cnt[level - 1] - 8 * k;
            if (nvalid < 8) mask &= (1u << nvalid) - 1u;
            // children whose leaves are all <= i: child c of this node covers the leaves [(8k + c) << sh, (8k + c + 1) << sh)
            const int sh = 3 * (level - 1);
            const long long cfirst = ((long long)(i + 1) >> sh) - 8ll * k;   // the child that holds leaf i + 1
            if (cfirst > 0) mask &= cfirst >= 8 ? 0u : (0xffu << (int)cfirst);
            if (!me_touched) mask &= h.w >> 24;
            int next = -1;
            if (level == 1) {
                while (mask) {
                    const int c = __ffs(mask) - 1;
                    mask &= mask - 1;
                    s_q[w][atomicAdd(&s_n[w], 1)] = make_int2(i, 8 * k + c);
                }
            } else {
                while (mask) {
                    const int c = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const int child = ((level - 1) << 27) | (8 * k + c);
                    if (next < 0) next = child;
                    else stack[sp++] = child;
                }
            }
            if (next >= 0) cur = next;
            else if (sp > 0) cur = stack[--sp];
            else cur = -1;
        }
        more = __any_sync(0xffffffffu, cur >= 0);
        __syncwarp();
        // drain: full batches while traversing, everything at the end
        int nq = s_n[w];
        while (nq >= 32 || (!more && nq > 0)) {
            const int take = nq < 32 ? nq : 32;
            const int base = nq - take;
            bool emit = false;
            int a = 0, b = 0;
            if (lane < take) {
                const int2 e = s_q[w][base + lane];
                const double2* qb = reinterpret_cast<const double2*>(lbox + 6 * (size_t)e.x);
                const double2* ob = reinterpret_cast<const double2*>(lbox + 6 * (size_t)e.y);
                const double2 q0 = __ldg(qb), q1 = __ldg(qb + 1), q2 = __ldg(qb + 2);
                const double2 c0b = __ldg(ob), c1b = __ldg(ob + 1), c2b = __ldg(ob + 2);
                // lo = (b0.x, b0.y, b1.x), hi = (b1.y, b2.x, b2.y); closed intervals
                const bool hit = q0.x <= c1b.y && q1.y >= c0b.x && q0.y <= c2b.x && q2.x >= c0b.y && q1.x <= c2b.y && q2.y >= c1b.x;
                if (hit) {
                    ++n_cand;
                    const int my_id = __ldg(leaf_elem + e.x), other_id = __ldg(leaf_elem + e.y);
                    a = min(my_id, other_id);
                    b = max(my_id, other_id);
                    if (out.dbg_cand) {
                        const unsigned long long sdb = atomicAdd(&out.counters[CTR_DBG_CAND], 1ull);
                        if ((long long)sdb < out.cap_dbg) out.dbg_cand[sdb] = make_int2(a, b);
                    }
                    const int4 me = __ldg(selem + e.x), ot = __ldg(selem + e.y);
                    // pairs sharing a vertex return false at once in every narrow-phase driver
                    // (dcollid3d.cpp:209-214, 257-264, 279-284, 491-496, 546-553, 574-579)
                    bool shared = me.x == ot.x || me.x == ot.y || me.y == ot.x || me.y == ot.y;
                    if (ot.z >= 0) shared = shared || me.x == ot.z || me.y == ot.z;
                    if (me.z >= 0) shared = shared || me.z == ot.x || me.z == ot.y || (ot.z >= 0 && me.z == ot.z);
                    emit = !shared;
                    // tri-tri on one surface with a rigid `a` is dropped (dcollid.cpp:762, 805)
                    if (emit && me.z >= 0 && ot.z >= 0) {
                        const int4 ea = my_id < other_id ? me : ot;
                        if (((me.w ^ ot.w) & 0x0fffffff) == 0 && (ea.w & 0x10000000)) emit = false;
                    }
                    // A feature test is a pure function of (x_old, avgVel) of its points.  If no point of the pair
                    // was touched by the previous pass's updateAverageVelocity, every test repeats its previous
                    // outcome, and that outcome was "no hit" (a hit would have changed one of the points; fixed and
                    // rigid-body points always count as touched).  Such pairs contribute nothing: skip them.
                    if (emit && out.dirty) {
                        unsigned d = out.dirty[me.x] | out.dirty[me.y] | out.dirty[ot.x] | out.dirty[ot.y];
                        if (me.z >= 0) d |= out.dirty[me.z];
                        if (ot.z >= 0) d |= out.dirty[ot.z];
                        if (!d) emit = false;
                    }
                }
            }
            const unsigned ballot = __ballot_sync(0xffffffffu, emit);
            if (ballot) {
                unsigned long long s0 = 0;
                if (lane == 0) s0 = atomicAdd(&out.counters[CTR_PAIRS], (unsigned long long)__popc(ballot));
                s0 = __shfl_sync(0xffffffffu, s0, 0);
                const unsigned long long sl = s0 + __popc(ballot & ((1u << lane) - 1u));
                if (emit && (long long)sl < out.cap_pairs) out.pairs[sl] = make_int2(a, b);
            }
            nq = base;
            __syncwarp();
        }
        __syncwarp();  // every lane has read s_n[w] before lane 0 rewrites it
        if (lane == 0) s_n[w] = nq;
        __syncwarp();
    }
    // candidate count: warp-reduce, one atomic per warp
    for (int o = 16; o > 0; o >>= 1) n_cand += __shfl_xor_sync(0xffffffffu, n_cand, o);
    if (lane == 0 && n_cand) atomicAdd(&out.counters[CTR_CAND], n_cand);
}

} // namespace clsn
