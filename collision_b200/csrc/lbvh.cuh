// Broad phase: Morton-code LBVH over the element leaf boxes (sm_100a).
//
// Replaces the reference's incremental-insertion AABBTree (AABB.cpp:129-245) and its per-leaf
// stack query (AABB.cpp:254-343).  What must be preserved is only the candidate SET: unordered
// element pairs whose FP64 leaf boxes (CD_TRI/CD_BOND::{min,max}_{static,moving}_coord -+ 1e-6,
// dcollid.cpp:852-932, AABB.cpp:6-36) overlap on closed intervals (AABB.cpp:56-60).  Tree shape
// is free.  Internal nodes carry FP32 boxes rounded outward (conservative); the leaf-leaf test is
// the exact FP64 one.
//
// Layout: leaves are sorted by Morton code; internal node i (Karras 2012) stores BOTH child boxes
// in one 64-byte record, so one traversal step is one 2-sector load.
#pragma once
#include <stdint.h>

namespace clsn {

struct alignas(32) Vec4 {  // one 32-byte sector per vertex gather
    double x, y, z, w;
};

struct alignas(64) WideNode {
    float lo0[3], hi0[3];  // left child box
    float lo1[3], hi1[3];  // right child box
    int c0, c1;            // child: >= 0 internal node index, < 0 leaf ~sorted_index
    int last;              // last sorted leaf index covered by this node (right child's max)
    int parent;            // parent internal node, -1 for the root
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

// Karras 2012 radix-tree construction over sorted (code, index) keys; ties broken by position.
__device__ __forceinline__ int delta_lcp(const unsigned* __restrict__ code, int N, int i, int j)
{
    if (j < 0 || j >= N) return -1;
    unsigned a = code[i], b = code[j];
    if (a == b) return 32 + __clz((unsigned)i ^ (unsigned)j);
    return __clz(a ^ b);
}

__global__ void k_hierarchy(const unsigned* __restrict__ code, int N, WideNode* __restrict__ nodes, int* __restrict__ leaf_parent)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N - 1) return;
    int d = (delta_lcp(code, N, i, i + 1) - delta_lcp(code, N, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta_lcp(code, N, i, i - d);
    int lmax = 2;
    while (delta_lcp(code, N, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta_lcp(code, N, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta_lcp(code, N, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (delta_lcp(code, N, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + min(d, 0);
    int first = min(i, j), last = max(i, j);
    int c0 = (first == gamma) ? ~gamma : gamma;
    int c1 = (last == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    nodes[i].c0 = c0;
    nodes[i].c1 = c1;
    nodes[i].last = last;
    if (i == 0) nodes[0].parent = -1;
    if (c0 < 0) leaf_parent[gamma] = i; else nodes[c0].parent = i;
    if (c1 < 0) leaf_parent[gamma + 1] = i; else nodes[c1].parent = i;
}

// Leaf boxes (exact FP64, stored in sorted order) + bottom-up refit of the FP32 child boxes.
// The second thread to reach a node (atomic flag) owns it; __threadfence orders the box stores.
template <bool MOVING>
__global__ void k_refit(const int4* __restrict__ elem, const int* __restrict__ leaf_elem, int N, const Vec4* __restrict__ xo,
                        const Vec4* __restrict__ av, double dt, double* __restrict__ lbox, WideNode* nodes,
                        const int* __restrict__ leaf_parent, int* flags, const uint8_t* __restrict__ vdirty)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int4 el = __ldg(elem + __ldg(leaf_elem + i));
    double lo[3], hi[3];
    leaf_box<MOVING>(el, xo, av, dt, lo, hi);
    // "touched" bit of the subtree (see k_traverse): does it hold a point changed by the previous pass?
    int dbit = 1;
    if (vdirty) {
        dbit = vdirty[el.x] | vdirty[el.y];
        if (el.z >= 0) dbit |= vdirty[el.z];
    }
    double2* lb = reinterpret_cast<double2*>(lbox + 6 * (size_t)i);
    lb[0] = make_double2(lo[0], lo[1]);
    lb[1] = make_double2(lo[2], hi[0]);
    lb[2] = make_double2(hi[1], hi[2]);
    if (N < 2) return;
    float flo[3], fhi[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        flo[d] = __double2float_rd(lo[d]);
        fhi[d] = __double2float_ru(hi[d]);
    }
    int node = leaf_parent[i];
    int child = ~i;
    while (true) {
        WideNode* nd = nodes + node;
        bool left = (nd->c0 == child);
        float* dlo = left ? nd->lo0 : nd->lo1;
        float* dhi = left ? nd->hi0 : nd->hi1;
#pragma unroll
        for (int d = 0; d < 3; ++d) { dlo[d] = flo[d]; dhi[d] = fhi[d]; }
        __threadfence();
        // arrival counter in the low byte; bit 8 / 9 = left / right child subtree is touched
        const int old = atomicAdd(flags + node, 1 + (dbit << (left ? 8 : 9)));
        if ((old & 0xff) == 0) return;  // sibling subtree not done yet
        dbit |= (old >> (left ? 9 : 8)) & 1;
        __threadfence();
        const volatile float* slo = left ? nd->lo1 : nd->lo0;
        const volatile float* shi = left ? nd->hi1 : nd->hi0;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            flo[d] = fminf(flo[d], slo[d]);
            fhi[d] = fmaxf(fhi[d], shi[d]);
        }
        int parent = nd->parent;
        if (parent < 0) return;
        child = node;
        node = parent;
    }
}

struct TraverseOut {
    int2* pairs;            // non-adjacent, unfiltered-by-narrow-phase element pairs (a < b)
    long long cap_pairs;
    int2* dbg_cand;         // every overlapping pair (debug / parity), may be null
    long long cap_dbg;
    unsigned long long* counters;
    const uint8_t* dirty;   // non-null: drop pairs none of whose points changed since the previous CCD pass
    const int* node_touched; // non-null (with dirty): refit's per-node flags, bit 8 / 9 = left / right subtree touched;
                             // an untouched query then only descends into touched subtrees (its other pairs would
                             // be dropped anyway).  `candidates` then counts only the pairs actually found.
};

#ifndef TRAV_THREADS
#define TRAV_THREADS 128
#endif
#define TRAV_QCAP 96

// Self-query.  Thread = one query leaf (sorted index i in [q_lo, q_hi)); finds leaves j > i whose exact
// FP64 boxes overlap (AABB::isCollid, AABB.cpp:56-60).  Subtrees whose last leaf is <= i are skipped, so
// each unordered pair is found once.
// The loop is kept lean and converged: every iteration is exactly one node visit per lane, and leaf
// children that pass the conservative FP32 test are only pushed to a per-warp shared-memory queue.
// Whenever the queue holds >= 32 entries the whole warp drains 32 of them together: exact FP64 test,
// adjacency / same-surface-rigid filters (dcollid3d.cpp:279-284, dcollid.cpp:762) and a ballot-aggregated
// append to the pair list -- all lanes busy instead of one or two.
__global__ void __launch_bounds__(TRAV_THREADS)
k_traverse(const WideNode* __restrict__ nodes, const double* __restrict__ lbox, const int* __restrict__ leaf_elem,
           const int4* __restrict__ elem, int N, int q_lo, int q_hi, TraverseOut out)
{
    __shared__ int2 s_q[TRAV_THREADS / 32][TRAV_QCAP];
    __shared__ int s_n[TRAV_THREADS / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int i = q_lo + blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long n_cand = 0;
    if (lane == 0) s_n[w] = 0;
    __syncwarp();
    float flo[3] = {0, 0, 0}, fhi[3] = {0, 0, 0};
    int node = -1;  // -1 = this lane has finished
    bool me_touched = true;
    if (i < q_hi && N >= 2 && out.node_touched) {
        const int4 el = __ldg(elem + __ldg(leaf_elem + i));
        unsigned d = out.dirty[el.x] | out.dirty[el.y];
        if (el.z >= 0) d |= out.dirty[el.z];
        me_touched = d != 0;
    }
    if (i < q_hi && N >= 2) {
        const double2* lb = reinterpret_cast<const double2*>(lbox + 6 * (size_t)i);
        const double2 b0 = __ldg(lb), b1 = __ldg(lb + 1), b2 = __ldg(lb + 2);
        flo[0] = __double2float_rd(b0.x); flo[1] = __double2float_rd(b0.y); flo[2] = __double2float_rd(b1.x);
        fhi[0] = __double2float_ru(b1.y); fhi[1] = __double2float_ru(b2.x); fhi[2] = __double2float_ru(b2.y);
        node = 0;
    }
    int stack[64];
    int sp = 0;
    bool more = __any_sync(0xffffffffu, node >= 0);
    while (more || s_n[w] > 0) {
        if (node >= 0) {
            const float4* np = reinterpret_cast<const float4*>(nodes + node);
            const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2);
            const int4 n3 = __ldg(reinterpret_cast<const int4*>(np + 3));
            // n0 = lo0.xyz hi0.x ; n1 = hi0.yz lo1.xy ; n2 = lo1.z hi1.xyz ; n3 = c0 c1 last parent
            const int c0 = n3.x, c1 = n3.y, last = n3.z;
            const int split = c0 < 0 ? ~c0 : c0;  // last leaf of the left child
            bool o0 = split > i && flo[0] <= n0.w && fhi[0] >= n0.x && flo[1] <= n1.x && fhi[1] >= n0.y &&
                      flo[2] <= n1.y && fhi[2] >= n0.z;
            bool o1 = last > i && flo[0] <= n2.y && fhi[0] >= n1.z && flo[1] <= n2.z && fhi[1] >= n1.w &&
                      flo[2] <= n2.w && fhi[2] >= n2.x;
            if (!me_touched && (o0 || o1)) {
                const int tf = __ldg(out.node_touched + node);
                o0 = o0 && ((tf >> 8) & 1);
                o1 = o1 && ((tf >> 9) & 1);
            }
            int next = -1;
            if (o0) {
                if (c0 >= 0) next = c0;
                else if (~c0 > i) s_q[w][atomicAdd(&s_n[w], 1)] = make_int2(i, ~c0);
            }
            if (o1) {
                if (c1 >= 0) { if (next < 0) next = c1; else stack[sp++] = c1; }
                else if (~c1 > i) s_q[w][atomicAdd(&s_n[w], 1)] = make_int2(i, ~c1);
            }
            if (next >= 0) node = next;
            else if (sp > 0) node = stack[--sp];
            else node = -1;
        }
        more = __any_sync(0xffffffffu, node >= 0);
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
                    const int4 me = __ldg(elem + my_id), ot = __ldg(elem + other_id);
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
