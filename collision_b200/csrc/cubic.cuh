// The coplanarity cubic of the CCD tests: coefficients, the trig-free classifier and the correctly
// rounded solve (reference: isCoplanar, dcollid3d.cpp:371-482).  Pure FP64 arithmetic, no CUDA intrinsics:
// the file is __host__ __device__ so that tests/cubic_check.cpp can fuzz exactly this code on the CPU
// against the oracle (tests/test_host_cpu.py::test_cubic_path_matches_oracle_on_host).
#pragma once
#include <math.h>
#include "crmath.cuh"

#if defined(__CUDACC__)
#define CLSN_HD __host__ __device__ __forceinline__
#else
#define CLSN_HD static inline
#endif

namespace clsn {

#define CLSN_MACH_EPS 2.220446049250313e-16 /* DBL_EPSILON */
#define CLSN_ROUND_EPS 1e-10                /* collid.h:17 */

// the four points of one feature test
struct Quad {
    int id[4];
    int flags[4];   // CLSN_VFLAG_*
    int body[4];
    double xo[4][3];  // x_old
    double av[4][3];  // avgVel
};


// Coefficients of the coplanarity cubic a t^3 + b t^2 + c t + d (dcollid3d.cpp:382-422), in the
// reference's exact operation order.
CLSN_HD void coplanar_coeffs(const Quad& q, double& a, double& b, double& c, double& d)
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
CLSN_HD int coplanar_maybe(double a, double b, double c, double d, double dt)
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

// FP32 pre-filter in front of coplanar_maybe(): "true" = isCoplanar certainly finds no root that survives its
// [0, dt] filter, decided without a single FP64 multiplication (experimental, -DCULL_PREFILTER=1).
//
// With times in units of dt and lengths scaled to O(1), the coplanarity cubic is the multilinear expansion of
// p(t) = X3(t) . (X1(t) x X2(t)),  Xi(t) = xi + t wi  (positions and per-step displacements relative to point 0).
// If its four Bernstein coefficients on [0, 1] have one sign, |p| >= m = min |k_j| on [0, 1] (convex hull).  A
// root the reference would keep, t^ in [MACH_EPS, dt + MACH_EPS], comes out of closed formulas whose error is
// bounded in terms of sigma = max(|A|, sqrt|B|, cbrt|C|) (A, B, C the monic coefficients; all roots are <= 2 sigma):
//   * t^ is within g = 3e-13 sigma of an exact root of the monic cubic with (Q, R) perturbed by the rounding of
//     their evaluation (trig branch: S + |A/3| <= 1.7 sigma; Cardano: |A_c| + |B_c| + |A/3| <= 2.8 sigma),
//     i.e. |p(t*)| <= alpha (3 sigma^2 + 5 sigma^3) 1e-13 for some |t* - t^| <= g;
//   * the Cardano branch's extra double root sits at the real part of a conjugate pair whose imaginary part is
//     below 1e-10 (absolute, in time units): |p| <= alpha (1 + 2 sigma) 0.75e-20 / dt^2 there.
// So no kept root exists if  m > E32 + U64,  E32 = 4e-6 * 6 prod(|xi| + |wi|) the FP32 evaluation error of a
// Bernstein coefficient and U64 the FP64-level terms above plus g |p'|max (1e-13 = ~900 u: two orders of
// magnitude of head-room over the operation counts).  Degenerate cubics (tiny leading coefficient: sigma huge,
// or the reference's quadratic / linear branches) make U64 large or fail the alpha test and stay undecided, as
// does every NaN (unordered comparisons are false).  Fuzzed against the oracle by tests/cubic_check.cpp.
CLSN_HD bool coplanar_prefilter32(const Quad& q, double dt)
{
    double dx[3][3], dw[3][3], big = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            dx[i][k] = q.xo[i + 1][k] - q.xo[0][k];
            dw[i][k] = (q.av[i + 1][k] - q.av[0][k]) * dt;
            big = fmax(big, fmax(fabs(dx[i][k]), fabs(dw[i][k])));
        }
    if (!(big > 1e-300 && big < 1e300 && dt > 1e-30)) return false;
    const double sc = 1.0 / big;
    float x[3][3], w[3][3], X[3], W[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        X[i] = 0.f; W[i] = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            x[i][k] = (float)(dx[i][k] * sc);
            w[i][k] = (float)(dw[i][k] * sc);
            X[i] = fmaxf(X[i], fabsf(x[i][k]));
            W[i] = fmaxf(W[i], fabsf(w[i][k]));
        }
    }
    // s . (u x v)
#define CLSN_TRIPLE(u, v, s) ((s)[0] * ((u)[1] * (v)[2] - (u)[2] * (v)[1]) + (s)[1] * ((u)[2] * (v)[0] - (u)[0] * (v)[2]) + \
                              (s)[2] * ((u)[0] * (v)[1] - (u)[1] * (v)[0]))
    const float d = CLSN_TRIPLE(x[0], x[1], x[2]);
    const float c = CLSN_TRIPLE(w[0], x[1], x[2]) + CLSN_TRIPLE(x[0], w[1], x[2]) + CLSN_TRIPLE(x[0], x[1], w[2]);
    const float b = CLSN_TRIPLE(x[0], w[1], w[2]) + CLSN_TRIPLE(w[0], x[1], w[2]) + CLSN_TRIPLE(w[0], w[1], x[2]);
    const float a = CLSN_TRIPLE(w[0], w[1], w[2]);
#undef CLSN_TRIPLE
    const float k0 = d, k1 = d + c * (1.f / 3.f), k2 = d + (2.f * c + b) * (1.f / 3.f), k3 = d + c + b + a;
    const bool pos = k0 > 0.f && k1 > 0.f && k2 > 0.f && k3 > 0.f;
    const bool neg = k0 < 0.f && k1 < 0.f && k2 < 0.f && k3 < 0.f;
    if (!(pos || neg)) return false;
    const float m = fminf(fminf(fabsf(k0), fabsf(k1)), fminf(fabsf(k2), fabsf(k3)));
    const float mtot = 6.f * (X[0] + W[0]) * (X[1] + W[1]) * (X[2] + W[2]);
    const float e32 = 4e-6f * mtot;
    if (!(m > 2.f * e32)) return false;
    // FP64-level terms; they matter only for nearly degenerate cubics, which must stay undecided
    const float al = fabsf(a);
    if (!(al > 64e-6f * W[0] * W[1] * W[2])) return false;   // leading coefficient known to ~10 %: else undecided
    const float A = fabsf(b) / al, B = fabsf(c) / al, C = fabsf(d) / al;
    const float sg = 1.2f * fmaxf(A, fmaxf(sqrtf(B), cbrtf(C)));
    const float dtf = (float)dt;
    const float g = 3e-13f * sg + 4e-16f / dtf;
    if (!(g < 0.01f)) return false;
    const float dmax = 1.1f * (fabsf(c) + 2.f * fabsf(b) + 3.f * al);
    const float u64 = al * (3.f * sg * sg + 5.f * sg * sg * sg) * 1e-13f + g * dmax +
                      al * (1.f + 2.f * sg) * (1e-20f / (dtf * dtf)) + 1e-13f * mtot;
    return m > 2.f * (e32 + u64);
}

// Separating-axis pre-filter along the feature's own normal, FP32.  "true" = the feature test certainly returns false.
//
// The swept-box cull (k_cull, boxes_far) rejects a feature when the two sub-features are separated along a coordinate
// axis by more than margin = 1.001 h + (3 eps + 2e-3) extent over the whole step.  The same argument holds for ANY fixed
// direction n, because positions are linear in time (x_old + t avgVel, dcollid3d.cpp:338): the projections n.x(t) are
// linear in t, so if  n.(P(t) - Q(t)) > margin |n|  holds at t = 0 and at t = dt for every vertex P of one sub-feature
// and Q of the other, it holds in between, the sub-features are further apart than the contact distance at every time
// the reference could test (its roots and dt), and PointToTri / EdgeToEdge return false there (dist > h, or a barycentric
// coordinate outside its eps-slack; the 2e-3 term covers the reference's own rounding on slivers, see boxes_far).
// n = the triangle normal (point-triangle) resp. the common normal of the two edges (edge-edge) at t = 0: layered cloth
// is separated along exactly this direction, not along a coordinate axis, and 43 % of the features that survive the boxes
// on config 4 end here -- before a single FP64 multiplication of the cubic.  Everything is relative to point 0 (FP64
// differences, then FP32: |rounding| <= 1e-6 extent, far inside the 2e-3 extent slack); a degenerate normal (parallel
// edges, sliver) never rejects.  moving = false: the static test of the proximity pass (one time only).
CLSN_HD bool sat_normal_far(const Quad& q, bool edge, bool moving, double dt, double h, double eps)
{
    float r[4][3], w[4][3];   // positions at t = 0 relative to point 0, displacements over the step
    float ext = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            r[i][k] = (float)(q.xo[i][k] - q.xo[0][k]);
            w[i][k] = moving ? (float)(q.av[i][k] * dt) : 0.f;
            ext = fmaxf(ext, fmaxf(fabsf(r[i][k]), fabsf(w[i][k])));
        }
    if (!(ext < 1e18f)) return false;
    // the two edge vectors spanning the normal: triangle (0,1),(0,2) resp. edges (0,1),(2,3)
    const float* e1 = r[1];
    float e2[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) e2[k] = edge ? r[3][k] - r[2][k] : r[2][k];
    const float n[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
    const float nn = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
    const float l1 = e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2], l2 = e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2];
    if (!(nn > 1e-4f * l1 * l2) || !(nn > 1e-30f)) return false;   // sin^2 < 1e-4: the normal is not trustworthy in FP32
    float rel = (float)(3.0 * eps + 2e-3);
    if (ext > 1.f) rel *= (ext * ext) * (ext * ext);                 // unit-scale figure, see boxes_far
    const float margin = (1.002f * (float)h + rel * 2.f * ext) * sqrtf(nn) * 1.0001f + 64.f * 1.2e-7f * ext * sqrtf(l1 * l2);
    // signed separations n.(P - Q): first sub-feature = points [0, na), second = [na, 4)
    const int na = edge ? 2 : 3;
    float smin = 3.0e38f, smax = -3.0e38f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (i >= na) break;
#pragma unroll
        for (int j = 2; j < 4; ++j) {
            if (j < na) continue;
            float d0 = 0.f, d1 = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float dr = r[j][k] - r[i][k];
                d0 += n[k] * dr;
                d1 += n[k] * (dr + (w[j][k] - w[i][k]));
            }
            smin = fminf(smin, fminf(d0, d1));
            smax = fmaxf(smax, fmaxf(d0, d1));
        }
    }
    return smin > margin || smax < -margin;
}

// isCoplanar, dcollid3d.cpp:371-482.  Returns true iff some root > MACH_EPS; roots[0..2] sorted.
// CLASSIFY = false when the caller has already run coplanar_maybe() on this feature (k_cull).
template <bool CLASSIFY>
CLSN_HD bool is_coplanar(const Quad& q, double dt, double* roots)
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

} // namespace clsn
