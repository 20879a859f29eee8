// Narrow phase + impulse response, FP64, one element pair per thread (sm_100a).
//
// Device restatement of the reference's dcollid3d.cpp for the CUDA path:
//   PointToTri :778-922, EdgeToEdge :643-776, isCoplanar :371-482, MovingPointToTri/EdgeToEdge
//   :327-369, PointToTriImpulse :925-1107, EdgeToEdgeImpulse :1109-1300, pair drivers :203-325,
//   :485-627, dispatch/filter dcollid.cpp:753-836.
// Operation order follows the reference's expressions exactly (left-to-right dot products, no
// FMA contraction: this TU is compiled with --fmad=false), IEEE div/sqrt; acos/cos/pow are the
// correctly rounded versions of crmath.cuh.  Instead of "+=" into per-point state through
// pointers, every fired feature emits up to four 64-byte point records (or 48-byte body records
// for rigid-rigid contacts) tagged with the canonical key (ea, eb, feature); reduce.cuh sums them
// per point in key order, which is what makes the result deterministic and rank-count independent.
#pragma once
#include <stdint.h>
#include "crmath.cuh"
#include "cubic.cuh"
#include "fastpath.cuh"

namespace clsn {


struct PointRec {  // 64 B
    unsigned long long key;
    int point;
    int pad;
    double imp[3];
    double fric[3];
};
struct BodyRec {  // 48 B
    unsigned long long key;
    int body;
    int pad;
    double v[3];
};
struct Contact {  // 96 B, == clsn_contact
    int ea, eb, feature, kind;
    int p[4];
    double root, dist, nor[3], w[3];
};

struct NarrowParams {
    double eps, thickness, dt, k, m, lambda, cr;
};

// what the record-emitting kernels need to push a record to its owner (passed by value inside Emit)
struct DistEmit {
    int nranks;                          // 1 = single GPU: records go to the local list
    int per_rank;                        // owner of vertex v = v / per_rank
    long long cap_region;                // records per (source, owner) region
    PointRec* const* peer_region;        // [owner] -> this source's region inside the owner's receive buffer
    unsigned long long* send_cnt;        // [owner] cursor of this pass
};

// one cursor bump per distinct owner among the converged lanes (a same-address atomic is serialised at the L2)
__device__ __forceinline__ unsigned long long reserve_owner(unsigned long long* cursors, int owner)
{
    const unsigned act = __activemask();
    const unsigned grp = __match_any_sync(act, owner);
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(grp) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(&cursors[owner], (unsigned long long)__popc(grp));
    base = __shfl_sync(grp, base, leader);
    return base + __popc(grp & ((1u << lane) - 1u));
}

struct Emit {
    PointRec* prec;
    ulonglong2* phdr;            // (key, point) of every record once more as a dense 16-byte array (k_scatter reads only this)
    BodyRec* brec;
    Contact* contacts;           // nullptr unless debug
    unsigned long long* counters; // see CTR_* in clsn.cu
    long long cap_prec, cap_brec, cap_contacts;
    int* cnt;                    // per-point record count (this pass)
    int* cnt_rg;                 // per-body record count (this pass)
    const double* body_mass;
    DistEmit D;                  // D.nranks > 1: point records go to their owners' receive buffers (dist.cuh)
};

enum { CTR_PAIRS = 0, CTR_CAND = 1, CTR_PREC = 2, CTR_BREC = 3, CTR_TRUE = 4, CTR_CONTACTS = 5, CTR_ERROR = 6,
       CTR_DBG_CAND = 7, CTR_FEATS = 8, CTR_BOXSURV = 9, CTR_ROOTS = 10, CTR_FEATS_EE = 11, CTR_ROOTS_EE = 12,
       CTR_HITS = 13, CTR_HITS_EE = 14, CTR_EXACT = 15, CTR_UNC = 16, CTR_UNC_EE = 17, CTR_OVF = 18,
       // the pair list is processed in chunks (clsn.cu: enqueue_detect): the six list cursors above restart with every chunk;
       // k_fold_chunk keeps their sums (CTR_TOT + 0..5: FEATS, FEATS_EE, UNC, UNC_EE, HITS, HITS_EE) and the largest chunk
       CTR_MAX_FEATS = 19, CTR_MAX_UNC = 20, CTR_MAX_HITS = 21, CTR_TOT = 22, CTR_COUNT = 28 };

__device__ __forceinline__ double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ double mag3(const double* a) { return sqrt(dot3(a, a)); }
__device__ __forceinline__ void cross3(const double* b, const double* c, double* r)
{
    r[0] = b[1] * c[2] - b[2] * c[1];
    r[1] = b[2] * c[0] - b[0] * c[2];
    r[2] = b[0] * c[1] - b[1] * c[0];
}
__device__ __forceinline__ void sub3(const double* a, const double* b, double* r)
{
    r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2];
}
__device__ __forceinline__ double stdmin(double a, double b) { return b < a ? b : a; }
__device__ __forceinline__ double stdmax(double a, double b) { return a < b ? b : a; }

__device__ __forceinline__ bool q_static(const Quad& q, int i) { return (q.flags[i] & 1) != 0; }
__device__ __forceinline__ bool q_movable(const Quad& q, int i) { return (q.flags[i] & 2) != 0; }
__device__ __forceinline__ bool q_rigid(const Quad& q, int i) { return (q.flags[i] & 3) != 0; }

// reserve n consecutive slots of a global append buffer, one atomic per converged group of lanes
__device__ __forceinline__ unsigned long long reserve(unsigned long long* ctr, int n)
{
    unsigned mask = __activemask();
    int lane = threadIdx.x & 31;
    int leader = __ffs(mask) - 1;
    int total, before;
    if (mask == 0xffffffffu) {
        // full warp: log-step inclusive scan
        int incl = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        total = __shfl_sync(0xffffffffu, incl, 31);
        before = incl - n;
    } else {
        // prefix of n over the active lanes only
        total = 0; before = 0;
        for (unsigned m = mask; m; m &= m - 1) {
            const int l = __ffs(m) - 1;
            const int v = __shfl_sync(mask, n, l);
            if (l < lane) before += v;
            total += v;
        }
    }
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(ctr, (unsigned long long)total);
    base = __shfl_sync(mask, base, leader);
    return base + before;
}

// one slot per calling lane: a single atomic per converged group.  (A same-address atomic costs ~0.85 cycles
// per LANE at the L2, chip-wide: six million un-aggregated increments of one counter are 2.7 ms.)
__device__ __forceinline__ unsigned long long reserve1(unsigned long long* ctr)
{
    const unsigned mask = __activemask();
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(ctr, (unsigned long long)__popc(mask));
    base = __shfl_sync(mask, base, leader);
    return base + __popc(mask & ((1u << lane) - 1u));
}

__device__ __forceinline__ void store_prec(PointRec* dst, unsigned long long key, int point, const double* imp, const double* fric)
{
    // 4 x 16-byte stores
    ulonglong2* d = reinterpret_cast<ulonglong2*>(dst);
    ulonglong2 h;
    h.x = key;
    h.y = (unsigned long long)(unsigned)point;
    d[0] = h;
    double2* dd = reinterpret_cast<double2*>(dst);
    dd[1] = make_double2(imp[0], imp[1]);
    dd[2] = make_double2(imp[2], fric[0]);
    dd[3] = make_double2(fric[1], fric[2]);
}

// Where an impulse record goes.  SEG = false: appended to the pass's record list (slot from reserve()), counted per
// point in E.cnt -- the reduction then groups the list with a counting sort.  SEG = true (pipeline 2): the per-point
// counts are known before emission (k_count_hits), so the record is written straight into its point's segment
// [offs[p], offs[p] + cnt[p]) of the same buffer; no grouping pass is needed afterwards.
struct SegOut {
    const int* offs;
    int* fill;
};
template <bool SEG>
__device__ __forceinline__ void put_prec(const Emit& E, const SegOut& S, unsigned long long& slot, unsigned long long key, int point,
                                         const double* imp, const double* fric)
{
    if (!SEG && E.D.nranks > 1) {
        // multi-GPU: straight into the owner's receive buffer over NVLink (dist.cuh); the owner counts per point
        const int owner = point / E.D.per_rank;
        const long long s = (long long)reserve_owner(E.D.send_cnt, owner);
        if (s < E.D.cap_region) store_prec(E.D.peer_region[owner] + s, key, point, imp, fric);
    } else if (SEG) {
        const long long s = (long long)S.offs[point] + atomicAdd(&S.fill[point], 1);
        if (s < E.cap_prec) store_prec(E.prec + s, key, point, imp, fric);
    } else {
        atomicAdd(&E.cnt[point], 1);
        if ((long long)slot < E.cap_prec) {
            store_prec(E.prec + slot, key, point, imp, fric);
            ulonglong2 h;
            h.x = key;
            h.y = (unsigned long long)(unsigned)point;
            E.phdr[slot] = h;
        }
        ++slot;
    }
}

__device__ __noinline__ void emit_contact(const Emit& E, int4 ids, unsigned long long key, int kind, double root,
                                          double dist, double n0, double n1, double n2, double w0, double w1, double w2)
{
    unsigned long long slot = reserve1(&E.counters[CTR_CONTACTS]);
    if (E.contacts && (long long)slot < E.cap_contacts) {
        Contact c;
        c.feature = (int)(key & 15ull);
        c.eb = (int)((key >> 4) & 0x3fffffffull);
        c.ea = (int)(key >> 34);
        c.kind = kind;
        c.p[0] = ids.x; c.p[1] = ids.y; c.p[2] = ids.z; c.p[3] = ids.w;
        c.root = root; c.dist = dist;
        c.nor[0] = n0; c.nor[1] = n1; c.nor[2] = n2;
        c.w[0] = w0; c.w[1] = w1; c.w[2] = w2;
        E.contacts[slot] = c;
    }
}

__device__ __forceinline__ void emit_body(const Emit& E, unsigned long long key, int body, double impulse, const double* nor)
{
    // SpreadImpactZoneImpulse, dcollid.cpp:1101-1115: one add per body instead of one per point of the body
    unsigned long long slot = reserve1(&E.counters[CTR_BREC]);
    atomicAdd(&E.cnt_rg[body], 1);
    if ((long long)slot < E.cap_brec) {
        BodyRec r;
        r.key = key; r.body = body; r.pad = 0;
        r.v[0] = impulse * nor[0]; r.v[1] = impulse * nor[1]; r.v[2] = impulse * nor[2];
        E.brec[slot] = r;
    }
}

// PointToTriImpulse, dcollid3d.cpp:925-1107.  q: 0..2 triangle, 3 point.  w is modified as in the reference.
template <bool SEG>
__device__ __forceinline__ void point_to_tri_impulse(const NarrowParams& P, const Emit& E, const SegOut& S, const Quad& q,
                                                      unsigned long long key, const double* nor, double* w, double dist)
{
    double v_rel[3] = {0.0, 0.0, 0.0}, vn, vt;
    double impulse = 0.0, m_impulse, sum_w = 0.0;
    double ri0 = 0.0, ri1 = 0.0;
    const double k = P.k, m = P.m, dt = P.dt, lambda = P.lambda, h = P.thickness, cr = P.cr;
    dist = h - dist;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        v_rel[i] += q.av[3][i];
#pragma unroll
        for (int j = 0; j < 3; ++j) v_rel[i] -= w[j] * q.av[j][i];
    }
    vn = dot3(v_rel, nor);
    if (dot3(v_rel, v_rel) > vn * vn) vt = sqrt(dot3(v_rel, v_rel) - vn * vn);
    else vt = 0.0;
    if (vn < 0) {
        if (q_static(q, 3) || (q_static(q, 0) && q_static(q, 1) && q_static(q, 2))) {
            impulse = vn; ri0 = vn; ri1 = vn;
        } else if (q_movable(q, 0) && q_movable(q, 1) && q_movable(q, 2) && q_movable(q, 3)) {
            double m1 = E.body_mass[q.body[0]], m2 = E.body_mass[q.body[3]];
            ri0 = vn * m2 / (m1 + m2);
            ri1 = vn * m1 / (m1 + m2);
        } else if (q_movable(q, 0) && q_movable(q, 1) && q_movable(q, 2)) {
            ri0 = 0.5 * vn; impulse = 0.5 * vn;
        } else if (q_movable(q, 3)) {
            impulse = 0.5 * vn; ri1 = 0.5 * vn;
        } else
            impulse = vn * 0.5;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (q_static(q, i)) w[i] = 0.0;
            sum_w += w[i];
        }
        if (fabs(sum_w) > CLSN_MACH_EPS) {
            double s = 1.0 / sum_w;
#pragma unroll
            for (int i = 0; i < 3; ++i) w[i] = s * w[i];
        }
    }
    const bool all_rigid = q_rigid(q, 0) && q_rigid(q, 1) && q_rigid(q, 2) && q_rigid(q, 3);
    if (vn * dt < 0.1 * dist) {
        if (all_rigid) {
            ri0 *= 1.0 + cr; ri1 *= 1.0 + cr;
        } else {
            double tmp = -stdmin(dt * k * dist / m, (0.1 * dist / dt - vn));
            impulse += tmp; ri0 += tmp; ri1 += tmp;
        }
    }
    if (fabs(sum_w) < CLSN_MACH_EPS) m_impulse = impulse;
    else m_impulse = 2.0 * impulse / (1.0 + dot3(w, w));
    if (all_rigid) {
        if (q_movable(q, 0)) emit_body(E, key, q.body[0], ri0, nor);
        if (q_movable(q, 3)) emit_body(E, key, q.body[3], -1.0 * ri1, nor);
        return;
    }
    int n = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) n += q_static(q, i) ? 0 : 1;
    unsigned long long slot = (SEG || E.D.nranks > 1) ? 0ull : reserve(&E.counters[CTR_PREC], n);
    const bool has_fric = fabs(vt) > CLSN_ROUND_EPS;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (q_static(q, i)) continue;
        double t_impulse = m_impulse;
        if (q_movable(q, i)) t_impulse = ri0;
        double imp[3], fric[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            imp[j] = w[i] * t_impulse * nor[j];
            fric[j] = has_fric ? stdmax(-fabs(lambda * w[i] * t_impulse / vt), -1.0) * (v_rel[j] - vn * nor[j]) : 0.0;
        }
        put_prec<SEG>(E, S, slot, key, q.id[i], imp, fric);
    }
    if (!q_static(q, 3)) {
        double t_impulse = m_impulse;
        if (q_movable(q, 3)) t_impulse = ri1;
        double imp[3], fric[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            // the reference does "collsnImpulse -= t*nor": adding the exact negation is the same operation
            imp[j] = -(t_impulse * nor[j]);
            fric[j] = has_fric ? stdmax(-fabs(lambda * t_impulse / vt), -1.0) * (v_rel[j] - vn * nor[j]) : 0.0;
        }
        put_prec<SEG>(E, S, slot, key, q.id[3], imp, fric);
    }
}

// EdgeToEdgeImpulse, dcollid3d.cpp:1109-1300.  q: edge 0-1 against edge 2-3.
template <bool SEG>
__device__ __forceinline__ void edge_to_edge_impulse(const NarrowParams& P, const Emit& E, const SegOut& S, const Quad& q,
                                                      unsigned long long key, const double* nor, double a, double b, double dist)
{
    double v_rel[3], vn, vt;
    double impulse = 0.0, m_impulse;
    double ri0 = 0.0, ri1 = 0.0;
    double wa0 = 1.0 - a, wa1 = a, wb0 = 1.0 - b, wb1 = b;
    const double k = P.k, m = P.m, dt = P.dt, lambda = P.lambda, h = P.thickness, cr = P.cr;
    dist = h - dist;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        v_rel[j] = (1.0 - b) * q.av[2][j] + b * q.av[3][j];
        v_rel[j] -= (1.0 - a) * q.av[0][j] + a * q.av[1][j];
    }
    vn = dot3(v_rel, nor);
    if (dot3(v_rel, v_rel) > vn * vn) vt = sqrt(dot3(v_rel, v_rel) - vn * vn);
    else vt = 0.0;
    if (vn < 0.0) {
        if ((q_static(q, 0) && q_static(q, 1)) || (q_static(q, 2) && q_static(q, 3))) {
            impulse = vn; ri0 = vn; ri1 = vn;
        } else if (q_movable(q, 0) && q_movable(q, 1) && q_movable(q, 2) && q_movable(q, 3)) {
            double m1 = E.body_mass[q.body[0]], m2 = E.body_mass[q.body[2]];
            ri0 = vn * m2 / (m1 + m2);
            ri1 = vn * m1 / (m1 + m2);
        } else if (q_movable(q, 0) && q_movable(q, 1)) {
            ri0 = 0.5 * vn; impulse = 0.5 * vn;
        } else if (q_movable(q, 2) && q_movable(q, 3)) {
            impulse = 0.5 * vn; ri1 = 0.5 * vn;
        } else
            impulse = vn * 0.5;
        if (q_static(q, 0)) wa0 = 0.0;
        if (q_static(q, 1)) wa1 = 0.0;
        if (q_static(q, 2)) wb0 = 0.0;
        if (q_static(q, 3)) wb1 = 0.0;
    }
    const bool all_rigid = q_rigid(q, 0) && q_rigid(q, 1) && q_rigid(q, 2) && q_rigid(q, 3);
    if (vn * dt < 0.1 * dist) {
        if (all_rigid) {
            ri0 *= 1.0 + cr; ri1 *= 1.0 + cr;
        } else {
            double tmp = -stdmin(dt * k * dist / m, (0.1 * dist / dt - vn));
            impulse += tmp; ri0 += tmp; ri1 += tmp;
        }
    }
    if (wa0 + wa1 < CLSN_MACH_EPS || wb0 + wb1 < CLSN_MACH_EPS) m_impulse = impulse;
    else m_impulse = 2.0 * impulse / (wa0 * wa0 + wa1 * wa1 + wb0 * wb0 + wb1 * wb1);
    if (all_rigid) {
        if (q_movable(q, 0)) emit_body(E, key, q.body[0], ri0, nor);
        if (q_movable(q, 2)) emit_body(E, key, q.body[2], -1.0 * ri1, nor);
        return;
    }
    int n = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) n += q_static(q, i) ? 0 : 1;
    unsigned long long slot = (SEG || E.D.nranks > 1) ? 0ull : reserve(&E.counters[CTR_PREC], n);
    const bool has_fric = fabs(vt) > CLSN_ROUND_EPS;
    const double wgt[4] = {wa0, wa1, wb0, wb1};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (q_static(q, i)) continue;
        double t_impulse = m_impulse;
        if (q_movable(q, i)) t_impulse = (i < 2) ? ri0 : ri1;
        double imp[3], fric[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            double t = wgt[i] * t_impulse * nor[j];
            imp[j] = (i < 2) ? t : -t;
            fric[j] = has_fric ? stdmax(-fabs(lambda * wgt[i] * t_impulse / vt), -1.0) * (v_rel[j] - vn * nor[j]) : 0.0;
        }
        put_prec<SEG>(E, S, slot, key, q.id[i], imp, fric);
    }
}

// PointToTri, dcollid3d.cpp:778-922.  X = positions at test time.
// EMIT = false: decision only (no records, no counters) -- used to find the first hit of a feature;
// EMIT = true: the same arithmetic followed by the contact record and the impulse.
template <bool EMIT, bool SEG = false>
__device__ __forceinline__ bool point_to_tri(const NarrowParams& P, const Emit& E, const Quad& q, unsigned long long key,
                                              const double X[4][3], double h, double root, const SegOut& S = SegOut{nullptr, nullptr})
{
    double w[3];
    double x13[3], x23[3], x43[3], nor[3], nor_mag, dist, det;
    sub3(X[0], X[2], x13);
    sub3(X[1], X[2], x23);
    sub3(X[3], X[2], x43);
    det = dot3(x13, x13) * dot3(x23, x23) - dot3(x13, x23) * dot3(x13, x23);
    if (fabs(det) < 1000 * CLSN_MACH_EPS) return false;
    cross3(x13, x23, nor);
    nor_mag = mag3(nor);
    double x43_old[3];
    sub3(q.xo[3], q.xo[2], x43_old);
    dist = dot3(x43_old, nor);
    {
        double den = nor_mag * ((dist >= 0) ? 1.0 : -1.0);
#pragma unroll
        for (int i = 0; i < 3; ++i) nor[i] /= den;
    }
    dist = fabs(dot3(x43, nor));
    w[0] = (dot3(x13, x43) * dot3(x23, x23) - dot3(x23, x43) * dot3(x13, x23)) / det;
    w[1] = (dot3(x13, x13) * dot3(x23, x43) - dot3(x13, x23) * dot3(x13, x43)) / det;
    w[2] = 1 - w[0] - w[1];
    if (fabs(w[0]) < CLSN_ROUND_EPS || fabs(w[1]) < CLSN_ROUND_EPS || fabs(w[2]) < CLSN_ROUND_EPS) {
        double vec[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) vec[j] = q.xo[3][j];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) vec[j] -= w[i] * q.xo[i][j];
        if (mag3(vec) > CLSN_ROUND_EPS) {
            nor[0] = vec[0]; nor[1] = vec[1]; nor[2] = vec[2];
        }
    }
    nor_mag = mag3(nor);
    if (nor_mag > CLSN_ROUND_EPS) {
#pragma unroll
        for (int i = 0; i < 3; ++i) nor[i] /= nor_mag;
    } else {
        if (!EMIT) atomicAdd(&E.counters[CTR_ERROR], 1ull);  // reference: clean_up(ERROR)
        return false;
    }
    if (dist > h) return false;
#pragma unroll
    for (int i = 0; i < 3; ++i)
        if (w[i] > 1 + P.eps || w[i] < -P.eps) return false;
    if (EMIT) {
        emit_contact(E, make_int4(q.id[0], q.id[1], q.id[2], q.id[3]), key, 0, root, dist, nor[0], nor[1], nor[2], w[0], w[1], w[2]);
        point_to_tri_impulse<SEG>(P, E, S, q, key, nor, w, dist);
    }
    return true;
}

// EdgeToEdge, dcollid3d.cpp:643-776
template <bool EMIT, bool SEG = false>
__device__ __forceinline__ bool edge_to_edge(const NarrowParams& P, const Emit& E, const Quad& q, unsigned long long key,
                                              const double X[4][3], double h, double root, const SegOut& S = SegOut{nullptr, nullptr})
{
    double x21[3], x43[3], x31[3], tmp[3], v1[3], v2[3], nor[3], nor_mag, dist, a, b;
    sub3(X[1], X[0], x21);
    sub3(X[3], X[2], x43);
    sub3(X[2], X[0], x31);
    cross3(x21, x43, tmp);
    if (mag3(tmp) < CLSN_ROUND_EPS) return false;
    {
        const double d2121 = dot3(x21, x21), d4343 = dot3(x43, x43), d2143 = dot3(x21, x43);
        const double d2131 = dot3(x21, x31), d4331 = dot3(x43, x31);
        const double den = d2121 * d4343 - d2143 * d2143;
        a = (d4343 * d2131 - d2143 * d4331) / den;
        b = (d2143 * d2131 - d2121 * d4331) / den;
    }
    a = stdmax(stdmin(a, 1.0), 0.0);
    b = stdmax(stdmin(b, 1.0), 0.0);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        v1[i] = X[0][i] + a * x21[i];
        v2[i] = X[2][i] + b * x43[i];
    }
    sub3(v2, v1, nor);
    nor_mag = mag3(nor);
    if (nor_mag < 1000 * CLSN_MACH_EPS) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            nor[j] = (1.0 - b) * q.xo[2][j] + b * q.xo[3][j];
            nor[j] -= (1.0 - a) * q.xo[0][j] + a * q.xo[1][j];
        }
    }
    dist = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) dist += (v1[i] - v2[i]) * (v1[i] - v2[i]);
    dist = sqrt(dist);
    if (dist > h) return false;
    nor_mag = mag3(nor);
    if (nor_mag < CLSN_MACH_EPS) {
        if (!EMIT) atomicAdd(&E.counters[CTR_ERROR], 1ull);  // reference: clean_up(ERROR)
        return false;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) nor[i] /= nor_mag;
    if (EMIT) {
        emit_contact(E, make_int4(q.id[0], q.id[1], q.id[2], q.id[3]), key, 1, root, dist, nor[0], nor[1], nor[2], a, b, 0.0);
        edge_to_edge_impulse<SEG>(P, E, S, q, key, nor, a, b, dist);
    }
    return true;
}

// positions of the four points at time t (CCD: x_old + t * avgVel, dcollid3d.cpp:338; proximity: x_old)
template <bool MOVING>
__device__ __forceinline__ void positions_at(const Quad& q, double t, double X[4][3])
{
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < 3; ++k) X[j][k] = MOVING ? q.xo[j][k] + t * q.av[j][k] : q.xo[j][k];
}

// MovingPointToTri / MovingEdgeToEdge (dcollid3d.cpp:327-369), second half: walk the sorted roots
// (invalid = -1) and then t = dt; the first time at which the static test fires wins.  Returns that time
// (>= 0) or -1.  Decision only: the contact record and the impulse are produced later, by
// feature_emit() at the returned time, with all lanes of a warp busy.
template <bool MOVING>
__device__ __forceinline__ double feature_first_hit(const NarrowParams& P, const Emit& E, const Quad& q, bool edge, double h,
                                                     double r0, double r1, double r2)
{
    double X[4][3];
    if (!MOVING) {
        positions_at<false>(q, 0.0, X);
        const bool hit = edge ? edge_to_edge<false>(P, E, q, 0ull, X, h, 0.0) : point_to_tri<false>(P, E, q, 0ull, X, h, 0.0);
        return hit ? 0.0 : -1.0;
    }
    for (int i = 0; i < 4; ++i) {
        const double t = i == 0 ? r0 : (i == 1 ? r1 : (i == 2 ? r2 : P.dt));
        if (t < 0) continue;
        positions_at<true>(q, t, X);
        const bool hit = edge ? edge_to_edge<false>(P, E, q, 0ull, X, h, t) : point_to_tri<false>(P, E, q, 0ull, X, h, t);
        if (hit) return t;
    }
    return -1.0;
}

template <bool MOVING, bool SEG = false>
__device__ __forceinline__ void feature_emit(const NarrowParams& P, const Emit& E, const Quad& q, unsigned long long key, bool edge,
                                             double h, double t, const SegOut& S = SegOut{nullptr, nullptr})
{
    double X[4][3];
    positions_at<MOVING>(q, t, X);
    if (edge) edge_to_edge<true, SEG>(P, E, q, key, X, h, t, S);
    else point_to_tri<true, SEG>(P, E, q, key, X, h, t, S);
}

} // namespace clsn
