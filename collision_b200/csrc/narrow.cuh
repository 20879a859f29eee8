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

namespace clsn {

#define CLSN_MACH_EPS 2.220446049250313e-16 /* DBL_EPSILON */
#define CLSN_ROUND_EPS 1e-10                /* collid.h:17 */

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

struct Emit {
    PointRec* prec;
    BodyRec* brec;
    Contact* contacts;           // nullptr unless debug
    unsigned long long* counters; // see CTR_* in clsn.cu
    long long cap_prec, cap_brec, cap_contacts;
    int* cnt;                    // per-point record count (this pass)
    int* cnt_rg;                 // per-body record count (this pass)
    const double* body_mass;
};

enum { CTR_PAIRS = 0, CTR_CAND = 1, CTR_PREC = 2, CTR_BREC = 3, CTR_TRUE = 4, CTR_CONTACTS = 5, CTR_ERROR = 6,
       CTR_DBG_CAND = 7, CTR_FEATS = 8, CTR_BOXSURV = 9, CTR_ROOTS = 10, CTR_FEATS_EE = 11, CTR_ROOTS_EE = 12, CTR_COUNT = 13 };

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

// the four points of one feature test
struct Quad {
    int id[4];
    int flags[4];   // CLSN_VFLAG_*
    int body[4];
    double xo[4][3];  // x_old
    double av[4][3];  // avgVel
};

__device__ __forceinline__ bool q_static(const Quad& q, int i) { return (q.flags[i] & 1) != 0; }
__device__ __forceinline__ bool q_movable(const Quad& q, int i) { return (q.flags[i] & 2) != 0; }
__device__ __forceinline__ bool q_rigid(const Quad& q, int i) { return (q.flags[i] & 3) != 0; }

// reserve n consecutive slots of a global append buffer, one atomic per converged group of lanes
__device__ __forceinline__ unsigned long long reserve(unsigned long long* ctr, int n)
{
    unsigned mask = __activemask();
    int lane = threadIdx.x & 31;
    int leader = __ffs(mask) - 1;
    // inclusive prefix of n over the active lanes
    int total = 0, before = 0;
    for (int l = 0; l < 32; ++l) {
        if (!((mask >> l) & 1u)) continue;
        int v = __shfl_sync(mask, n, l);
        if (l < lane) before += v;
        total += v;
    }
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(ctr, (unsigned long long)total);
    base = __shfl_sync(mask, base, leader);
    return base + before;
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

__device__ __noinline__ void emit_contact(const Emit& E, int4 ids, unsigned long long key, int kind, double root,
                                          double dist, double n0, double n1, double n2, double w0, double w1, double w2)
{
    unsigned long long slot = atomicAdd(&E.counters[CTR_CONTACTS], 1ull);
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
    unsigned long long slot = atomicAdd(&E.counters[CTR_BREC], 1ull);
    atomicAdd(&E.cnt_rg[body], 1);
    if ((long long)slot < E.cap_brec) {
        BodyRec r;
        r.key = key; r.body = body; r.pad = 0;
        r.v[0] = impulse * nor[0]; r.v[1] = impulse * nor[1]; r.v[2] = impulse * nor[2];
        E.brec[slot] = r;
    }
}

// PointToTriImpulse, dcollid3d.cpp:925-1107.  q: 0..2 triangle, 3 point.  w is modified as in the reference.
__device__ __forceinline__ void point_to_tri_impulse(const NarrowParams& P, const Emit& E, const Quad& q,
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
    unsigned long long slot = reserve(&E.counters[CTR_PREC], n);
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
        atomicAdd(&E.cnt[q.id[i]], 1);
        if ((long long)slot < E.cap_prec) store_prec(E.prec + slot, key, q.id[i], imp, fric);
        ++slot;
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
        atomicAdd(&E.cnt[q.id[3]], 1);
        if ((long long)slot < E.cap_prec) store_prec(E.prec + slot, key, q.id[3], imp, fric);
    }
}

// EdgeToEdgeImpulse, dcollid3d.cpp:1109-1300.  q: edge 0-1 against edge 2-3.
__device__ __forceinline__ void edge_to_edge_impulse(const NarrowParams& P, const Emit& E, const Quad& q,
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
    unsigned long long slot = reserve(&E.counters[CTR_PREC], n);
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
        atomicAdd(&E.cnt[q.id[i]], 1);
        if ((long long)slot < E.cap_prec) store_prec(E.prec + slot, key, q.id[i], imp, fric);
        ++slot;
    }
}

// PointToTri, dcollid3d.cpp:778-922.  X = positions at test time.
// EMIT = false: decision only (no records, no counters) -- used to find the first hit of a feature;
// EMIT = true: the same arithmetic followed by the contact record and the impulse.
template <bool EMIT>
__device__ __forceinline__ bool point_to_tri(const NarrowParams& P, const Emit& E, const Quad& q, unsigned long long key,
                                              const double X[4][3], double h, double root)
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
        point_to_tri_impulse(P, E, q, key, nor, w, dist);
    }
    return true;
}

// EdgeToEdge, dcollid3d.cpp:643-776
template <bool EMIT>
__device__ __forceinline__ bool edge_to_edge(const NarrowParams& P, const Emit& E, const Quad& q, unsigned long long key,
                                              const double X[4][3], double h, double root)
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
        edge_to_edge_impulse(P, E, q, key, nor, a, b, dist);
    }
    return true;
}

// Coefficients of the coplanarity cubic a t^3 + b t^2 + c t + d (dcollid3d.cpp:382-422), in the
// reference's exact operation order.
__device__ __forceinline__ void coplanar_coeffs(const Quad& q, double& a, double& b, double& c, double& d)
{
    double v[4][3], x[4][3];
#pragma unroll
    for (int i = 1; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            v[i][j] = q.av[i][j] - q.av[0][j];
            x[i][j] = q.xo[i][j] - q.xo[0][j];
        }
    double vv[3], vx[3], xx[3];
    vv[0] = v[1][1] * v[2][2] - v[1][2] * v[2][1];
    vv[1] = v[1][0] * v[2][2] - v[1][2] * v[2][0];
    vv[2] = v[1][0] * v[2][1] - v[1][1] * v[2][0];
    vx[0] = v[1][1] * x[2][2] - v[1][2] * x[2][1] - v[2][1] * x[1][2] + v[2][2] * x[1][1];
    vx[1] = v[1][0] * x[2][2] - v[1][2] * x[2][0] - v[2][0] * x[1][2] + v[2][2] * x[1][0];
    vx[2] = v[1][0] * x[2][1] - v[1][1] * x[2][0] - v[2][0] * x[1][1] + v[2][1] * x[1][0];
    xx[0] = x[1][1] * x[2][2] - x[1][2] * x[2][1];
    xx[1] = x[1][0] * x[2][2] - x[1][2] * x[2][0];
    xx[2] = x[1][0] * x[2][1] - x[1][1] * x[2][0];
    a = v[3][0] * vv[0] - v[3][1] * vv[1] + v[3][2] * vv[2];
    b = x[3][0] * vv[0] - x[3][1] * vv[1] + x[3][2] * vv[2] + v[3][0] * vx[0] - v[3][1] * vx[1] + v[3][2] * vx[2];
    c = x[3][0] * vx[0] - x[3][1] * vx[1] + x[3][2] * vx[2] + v[3][0] * xx[0] - v[3][1] * xx[1] + v[3][2] * xx[2];
    d = x[3][0] * xx[0] - x[3][1] * xx[1] + x[3][2] * xx[2];
}

// Rigorous, trig-free classifier: can isCoplanar produce a root that survives its "-MACH_EPS, keep
// [0, dt]" filter?  "false" is exact (the reference's isCoplanar returns false); "true" sends the
// feature to the correctly rounded solve.  (a, b, c, d) as returned by coplanar_coeffs.
//
// Three-real-root branch (R^2 < Q^3): the reference's roots are -2 sqrt(Q) cos(phi_k) - a/3 with
// phi_k = (acos(x) + 2 pi k)/3, x = R / sqrt(Q^3); cos(phi_k) are the three solutions of the
// triple-angle identity T3(c) = 4c^3 - 3c = x.  A computed root is valid only if its cosine lies in
// [u, v] = [-(dt + a/3)/S, -(a/3)/S] (S = 2 sqrt(Q)) up to the rounding of the root formula, so no
// root is valid when x is outside T3([u - delta, v + delta]) -- two polynomial evaluations, no acos/cos.
// delta covers: correctly rounded acos/cos (<= 1 ulp total in phi and cos), the rounding of 2*pi,
// of the sum and of /3 (<= 1e-15), and the two roundings of the root formula (eta).  DESIGN.md.
// One-real-root branch: root = A + B - a/3 with A = -sgn * pow(u, 1/3); cbrt() is within 1.5e-14
// relative of the correctly rounded pow over the whole double range, covered by a 1e-13 guard.
// Returns 0 = no valid root possible, 1 = maybe (three-real-root / trig branch), 2 = maybe (other branches).
__device__ __forceinline__ int coplanar_maybe(double a, double b, double c, double d, double dt)
{
    if (fabs(a) > CLSN_MACH_EPS) {
        b /= a; c /= a; d /= a;
        a = b; b = c; c = d;
        const double Q = (a * a - 3 * b) / 9;
        const double R = (2 * a * a * a - 9 * a * b + 27 * c) / 54;
        const double Q3 = Q * Q * Q, R2 = R * R;
        if (R2 < Q3) {
            const double S = 2 * sqrt(Q);
            const double x = R / sqrt(Q3);
            if (!(fabs(x) <= 1.0) || !(S > 0.0)) return 1;  // acos domain edge / NaN: let the exact path decide
            const double A3 = a / 3;
            const double eta = 4e-16 * (2 * S + fabs(A3) + dt) + 2 * CLSN_MACH_EPS;
            double u = -(dt + A3 + eta) / S;
            double v = -(A3 - eta) / S;
            const double delta = 4e-15 + 4e-16 * (fabs(u) + fabs(v));
            u -= delta;
            v += delta;
            if (u > 1.0 || v < -1.0) return 0;  // the cosines live in [-1, 1]
            u = fmax(u, -1.0);
            v = fmin(v, 1.0);
            const double tu = u * (4 * u * u - 3), tv = v * (4 * v * v - 3);
            double tmin = fmin(tu, tv), tmax = fmax(tu, tv);
            if (u <= -0.5 && v >= -0.5) tmax = 1.0;   // interior maximum of T3 at c = -1/2
            if (u <= 0.5 && v >= 0.5) tmin = -1.0;    // interior minimum at c = +1/2
            return !(x < tmin - 4e-14 || x > tmax + 4e-14) ? 1 : 0;
        }
        const double sgn = (R > 0) ? 1.0 : -1.0;
        const double A = -sgn * cbrt(fabs(R) + sqrt(R2 - Q3));
        if (!(fabs(fabs(A) - CLSN_ROUND_EPS) > 1e-20)) return 2;  // the |A| < 1e-10 switch could flip
        const double Bv = (fabs(A) < CLSN_ROUND_EPS) ? 0.0 : Q / A;
        const double g = (fabs(A) + fabs(Bv) + fabs(a)) * 1e-13;
        const double r0 = (A + Bv) - a / 3.0;
        bool maybe = !(r0 < -g || r0 > dt + g);
        if (!(fabs(A - Bv) > 2 * CLSN_ROUND_EPS)) {  // the double-root branch (|A-B| < 1e-10) may be taken
            const double rr = -0.5 * (A + Bv) - a / 3.0;
            maybe = maybe || !(rr < -g || rr > dt + g);
        }
        return maybe ? 2 : 0;
    }
    // quadratic / linear fall-backs use IEEE operations only: evaluate them as the reference does
    a = b; b = c; c = d;
    const double delta = b * b - 4.0 * a * c;
    double r0 = -1.0, r1 = -1.0;
    if (fabs(a) > CLSN_ROUND_EPS && delta > 0) {
        const double ds = sqrt(delta);
        r0 = (-b + ds) / (2.0 * a);
        r1 = (-b - ds) / (2.0 * a);
    } else if (fabs(a) < CLSN_ROUND_EPS && fabs(b) > CLSN_ROUND_EPS) {
        r0 = -c / b;
    }
    r0 -= CLSN_MACH_EPS;
    r1 -= CLSN_MACH_EPS;
    const bool ok0 = !(r0 < 0 || r0 > dt), ok1 = !(r1 < 0 || r1 > dt);
    return (ok0 || ok1) ? 2 : 0;
}

// isCoplanar, dcollid3d.cpp:371-482.  Returns true iff some root > MACH_EPS; roots[0..2] sorted.
// CLASSIFY = false when the caller has already run coplanar_maybe() on this feature (k_cull).
template <bool CLASSIFY>
__device__ __forceinline__ bool is_coplanar(const Quad& q, double dt, double* roots)
{
    double a, b, c, d;
    coplanar_coeffs(q, a, b, c, d);
    if (CLASSIFY && !coplanar_maybe(a, b, c, d, dt)) return false;  // every root provably outside [0, dt]
    if (fabs(a) > CLSN_MACH_EPS) {
        b /= a; c /= a; d /= a;
        a = b; b = c; c = d;
        double Q = (a * a - 3 * b) / 9;
        double R = (2 * a * a * a - 9 * a * b + 27 * c) / 54;
        double Q3 = Q * Q * Q, R2 = R * R;
        if (R2 < Q3) {
            double Qsqrt = sqrt(Q);
            const double arg = R / sqrt(Q3);
            const double two_pi = 2 * 3.14159265358979323846;
            // Which of the three roots can survive the [0, dt] filter?  Root k is -S cos(phi_k) - a/3 with
            // phi_0 in [0, pi/3], phi_1 in [2pi/3, pi], phi_2 in [-2pi/3, -pi/3], i.e. its cosine lies in
            // [1/2, 1], [-1, -1/2], [-1/2, 1/2] respectively, and a valid root has its cosine in the (tiny)
            // interval [u, v] derived in coplanar_maybe().  Only the overlapping k are evaluated (correctly
            // rounded); the others get -1, which is what the filter below would turn them into anyway.
            const double S = 2 * Qsqrt, A3 = a / 3;
            bool need0 = true, need1 = true, need2 = true;
            if (S > 0.0 && fabs(arg) <= 1.0) {
                const double eta = 4e-16 * (2 * S + fabs(A3) + dt) + 2 * CLSN_MACH_EPS;
                double u = -(dt + A3 + eta) / S;
                double v = -(A3 - eta) / S;
                const double delta = 4e-15 + 4e-16 * (fabs(u) + fabs(v));
                u -= delta;
                v += delta;
                const double e = 1e-14;
                need0 = !(v < 0.5 - e);                      // overlaps [1/2, 1]
                need1 = !(u > -0.5 + e);                     // overlaps [-1, -1/2]
                need2 = !(v < -0.5 - e || u > 0.5 + e);      // overlaps [-1/2, 1/2]
            }
            if (need0 || need1 || need2) {
                const double theta = crm::acos_cr(arg);
                if (need0) roots[0] = -2 * Qsqrt * crm::cos_cr(theta / 3) - a / 3;
                if (need1) roots[1] = -2 * Qsqrt * crm::cos_cr((theta + two_pi) / 3) - a / 3;
                if (need2) roots[2] = -2 * Qsqrt * crm::cos_cr((theta - two_pi) / 3) - a / 3;
            }
        } else {
            double sgn = (R > 0) ? 1.0 : -1.0;
            double A = -sgn * crm::pow13_cr(fabs(R) + sqrt(R2 - Q3));
            double Bv = (fabs(A) < CLSN_ROUND_EPS) ? 0.0 : Q / A;
            roots[0] = (A + Bv) - a / 3.0;
            if (fabs(A - Bv) < CLSN_ROUND_EPS) roots[1] = roots[2] = -0.5 * (A + Bv) - a / 3.0;
        }
    } else {
        a = b; b = c; c = d;
        double delta = b * b - 4.0 * a * c;
        if (fabs(a) > CLSN_ROUND_EPS && delta > 0) {
            double ds = sqrt(delta);
            roots[0] = (-b + ds) / (2.0 * a);
            roots[1] = (-b - ds) / (2.0 * a);
        } else if (fabs(a) < CLSN_ROUND_EPS && fabs(b) > CLSN_ROUND_EPS) {
            roots[0] = -c / b;
        }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        roots[i] = roots[i] - CLSN_MACH_EPS;
        if (roots[i] < 0 || roots[i] > dt) roots[i] = -1;
    }
    double t;
    if (roots[0] > roots[1]) { t = roots[0]; roots[0] = roots[1]; roots[1] = t; }
    if (roots[0] > roots[2]) { t = roots[0]; roots[0] = roots[2]; roots[2] = t; }
    if (roots[1] > roots[2]) { t = roots[1]; roots[1] = roots[2]; roots[2] = t; }
    return roots[0] > CLSN_MACH_EPS || roots[1] > CLSN_MACH_EPS || roots[2] > CLSN_MACH_EPS;
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

template <bool MOVING>
__device__ __forceinline__ void feature_emit(const NarrowParams& P, const Emit& E, const Quad& q, unsigned long long key, bool edge,
                                             double h, double t)
{
    double X[4][3];
    positions_at<MOVING>(q, t, X);
    if (edge) edge_to_edge<true>(P, E, q, key, X, h, t);
    else point_to_tri<true>(P, E, q, key, X, h, t);
}

} // namespace clsn
