// Fast path of the CCD feature test (MovingPointToTri / MovingEdgeToEdge, dcollid3d.cpp:327-369).
//
// The reference solves the coplanarity cubic with libm acos/cos/pow and then runs the static test at
// every valid root and at t = dt; the first hit wins.  Reproducing the root BITS needs the correctly
// rounded double-double evaluation of crmath.cuh (~1000 FP64 instructions), but four features out of
// five do not fire at any root, and for those the bits of the root never reach the result.  This file
// decides, in plain FP64 and with proven margins, between
//   FAST_MISS      isCoplanar returns false: no root survives the [0, dt] filter -> no test at all;
//   FAST_DT_ONLY   isCoplanar returns true and the static test misses at every valid root, so the
//                  outcome of the feature is the outcome of the static test at t = dt -- a time that is
//                  known exactly (the caller runs that test with the reference's own arithmetic);
//   FAST_UNCERTAIN anything else: the caller runs the correctly rounded solve and the exact tests.
// Nothing here produces a value that is stored: the fast path only ever REMOVES work whose outcome is
// certain, so results stay bit-identical (tests/fastpath_check.cpp fuzzes exactly this code on the
// host against the oracle, with a deliberately perturbed libm).
//
// Error model.  Approximate and reference roots come from the same IEEE expression tree and differ
// only through acos / cos / pow(u, 1/3):
//   trig branch     root = -2 sqrt(Q) cos((acos(x) + 2 pi k)/3) - a/3.  A libm within 4 ulp of the correctly
//                   rounded one moves cos(...) by < 2e-15, the root by < 2e-15 (S + |a/3|), S = 2 sqrt(Q);
//                   tol = 2e-14 (S + |a/3|).
//   Cardano branch  A = -sgn pow(u, 1/3): cbrt(u) is within 1.5e-14 relative of pow(u, (double)(1/3)) over the
//                   whole double range; tol = 4e-14 (|A| + |B| + |a/3|); the |A| < 1e-10 and |A - B| < 1e-10
//                   switches are "uncertain" inside that band.
//   quadratic / linear branches use IEEE operations only: tol = 0, roots identical.
// A root at distance > tol from 0, MACH_EPS and dt is classified like the reference's.  The static test at
// the approximate time sees every point moved by at most delta = tol * max|avgVel| + 4 MACH_EPS * max|x|;
// the barycentric coordinates (a, b of the edge pair) are the solution of a 2x2 least-squares system
// whose perturbation bound is  err = 16 eta (P + L (1 + |w0| + |w1|)) / (l^2 sin^2)  with eta = 4 delta + 32 u L,
// L / l the longer / shorter edge, P the length of the right-hand side, sin^2 = det / (G11 G22)
// (normal-equation perturbation, lambda_min(G) >= l^2 sin^2 / 2, both evaluations' rounding included).
// "Certain miss" = a barycentric coordinate outside [-eps, 1 + eps] by more than err (point-triangle), or
// the distance of the clamped closest points above h by more than its bound (edge-edge).  A degenerate
// configuration has a huge err and lands in FAST_UNCERTAIN by itself; so does every NaN (all tests are
// written so that an unordered comparison means "uncertain").
#pragma once
#include <math.h>
#include "cubic.cuh"

#ifndef CLSN_FAST_ACOS
#define CLSN_FAST_ACOS(x) acos(x)
#define CLSN_FAST_COS(x) cos(x)
#define CLSN_FAST_CBRT(x) cbrt(x)
#endif

namespace clsn {

enum { FAST_MISS = 0, FAST_DT_ONLY = 1, FAST_UNCERTAIN = 2 };

CLSN_HD double fp_dot(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// Point (X[3]) against triangle (X[0..2]) at an approximate root time: true = the reference's PointToTri
// (dcollid3d.cpp:778-922) certainly returns false at the exact time.  delta = bound on the displacement of
// every point between the two times (incl. rounding of the positions).
CLSN_HD bool pt_certain_miss(const double X[4][3], double delta, double eps)
{
    double e1[3], e2[3], p[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        e1[k] = X[0][k] - X[2][k];
        e2[k] = X[1][k] - X[2][k];
        p[k] = X[3][k] - X[2][k];
    }
    const double g11 = fp_dot(e1, e1), g22 = fp_dot(e2, e2), g12 = fp_dot(e1, e2);
    const double b1 = fp_dot(e1, p), b2 = fp_dot(e2, p), pp = fp_dot(p, p);
    const double det = g11 * g22 - g12 * g12;
    if (!(g11 > 1e-60 && g22 > 1e-60 && g11 < 1e60 && g22 < 1e60 && pp < 1e60 && det > 0.0)) return false;
    const double w0 = (b1 * g22 - b2 * g12) / det;
    const double w1 = (g11 * b2 - g12 * b1) / det;
    const double w2 = 1.0 - w0 - w1;
    const double gmax = fmax(g11, g22), gmin = fmin(g11, g22);
    const double L = sqrt(gmax), P = sqrt(pp);
    const double eta = 4.0 * delta + 32.0 * CLSN_MACH_EPS * L;
    // l^2 sin^2 = gmin * det / (g11 g22) = det / gmax
    const double err = 16.0 * eta * (P + L * (1.0 + fabs(w0) + fabs(w1))) * gmax / det;
    if (!(err < 0.25)) return false;
    const double lo = -eps - err, hi = 1.0 + eps + err;
    const double lo2 = -eps - 2.0 * err - 1e-15, hi2 = 1.0 + eps + 2.0 * err + 1e-15;
    return w0 < lo || w0 > hi || w1 < lo || w1 > hi || w2 < lo2 || w2 > hi2;
}

// Edge X[0]-X[1] against edge X[2]-X[3] at an approximate root time: true = the reference's EdgeToEdge
// (dcollid3d.cpp:643-776) certainly returns false at the exact time (the distance of its clamped closest
// points is certainly above h; parallel edges return false there as well).
CLSN_HD bool ee_certain_miss(const double X[4][3], double delta, double h, double xmax)
{
    double x21[3], x43[3], x31[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        x21[k] = X[1][k] - X[0][k];
        x43[k] = X[3][k] - X[2][k];
        x31[k] = X[2][k] - X[0][k];
    }
    const double g11 = fp_dot(x21, x21), g22 = fp_dot(x43, x43), g12 = fp_dot(x21, x43);
    const double b1 = fp_dot(x21, x31), b2 = fp_dot(x43, x31), pp = fp_dot(x31, x31);
    const double det = g11 * g22 - g12 * g12;
    if (!(g11 > 1e-60 && g22 > 1e-60 && g11 < 1e60 && g22 < 1e60 && pp < 1e60 && det > 0.0)) return false;
    double a = (g22 * b1 - g12 * b2) / det;
    double b = (g12 * b1 - g11 * b2) / det;
    const double gmax = fmax(g11, g22);
    const double L = sqrt(gmax), P = sqrt(pp);
    const double eta = 4.0 * delta + 32.0 * CLSN_MACH_EPS * L;
    const double err = 16.0 * eta * (P + L * (1.0 + fabs(a) + fabs(b))) * gmax / det;
    if (!(err < 0.25)) return false;
    a = fmax(fmin(a, 1.0), 0.0);
    b = fmax(fmin(b, 1.0), 0.0);
    double d2 = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double v = (a * x21[k] - b * x43[k]) - x31[k];
        d2 += v * v;
    }
    const double dist = sqrt(d2);
    // movement of the four points, of a and b (clamping is 1-Lipschitz), rounding in absolute coordinates
    const double err_d = 12.0 * delta + 2.0 * L * err + 64.0 * CLSN_MACH_EPS * (xmax + L + P);
    return dist - err_d > h * (1.0 + 1e-9);
}

// See the header comment.  q = the four points of the feature in the order PointToTri / EdgeToEdge take them.
CLSN_HD int feature_fast(const Quad& q, bool edge, double dt, double h, double eps)
{
    double a, b, c, d;
    coplanar_coeffs(q, a, b, c, d);
    double r[3] = {-1.0, -1.0, -1.0};
    unsigned set = 0;
    double tol = 0.0;
    if (fabs(a) > CLSN_MACH_EPS) {
        b /= a; c /= a; d /= a;
        a = b; b = c; c = d;
        const double Q = (a * a - 3 * b) / 9;
        const double R = (2 * a * a * a - 9 * a * b + 27 * c) / 54;
        const double Q3 = Q * Q * Q, R2 = R * R;
        if (R2 < Q3) {
            const double Qsqrt = sqrt(Q);
            const double arg = R / sqrt(Q3);
            const double S = 2 * Qsqrt, A3 = a / 3;
            if (!(S > 0.0 && fabs(arg) <= 1.0)) return FAST_UNCERTAIN;
            // which k can survive the filter: same interval logic as is_coplanar()
            const double eta = 4e-16 * (2 * S + fabs(A3) + dt) + 2 * CLSN_MACH_EPS;
            double u = -(dt + A3 + eta) / S;
            double v = -(A3 - eta) / S;
            const double delta = 4e-15 + 4e-16 * (fabs(u) + fabs(v));
            u -= delta;
            v += delta;
            const double e = 1e-14;
            const bool need0 = !(v < 0.5 - e), need1 = !(u > -0.5 + e), need2 = !(v < -0.5 - e || u > 0.5 + e);
            const double two_pi = 2 * 3.14159265358979323846;
            if (need0 || need1 || need2) {
                const double theta = CLSN_FAST_ACOS(arg);
                if (need0) { r[0] = -2 * Qsqrt * CLSN_FAST_COS(theta / 3) - a / 3; set |= 1u; }
                if (need1) { r[1] = -2 * Qsqrt * CLSN_FAST_COS((theta + two_pi) / 3) - a / 3; set |= 2u; }
                if (need2) { r[2] = -2 * Qsqrt * CLSN_FAST_COS((theta - two_pi) / 3) - a / 3; set |= 4u; }
            }
            tol = 2e-14 * (S + fabs(A3));
        } else {
            const double sgn = (R > 0) ? 1.0 : -1.0;
            const double A = -sgn * CLSN_FAST_CBRT(fabs(R) + sqrt(R2 - Q3));
            if (!(fabs(fabs(A) - CLSN_ROUND_EPS) > 1e-20)) return FAST_UNCERTAIN;  // the |A| < 1e-10 switch could flip
            const double Bv = (fabs(A) < CLSN_ROUND_EPS) ? 0.0 : Q / A;
            const double A3 = a / 3.0;
            r[0] = (A + Bv) - A3; set |= 1u;
            const double eab = 4e-14 * (fabs(A) + fabs(Bv));
            const double gap = fabs(A - Bv);
            if (gap < CLSN_ROUND_EPS - eab) { r[1] = r[2] = -0.5 * (A + Bv) - A3; set |= 6u; }
            else if (!(gap > CLSN_ROUND_EPS + eab)) return FAST_UNCERTAIN;
            tol = 4e-14 * (fabs(A) + fabs(Bv) + fabs(A3));
        }
    } else {
        a = b; b = c; c = d;
        const double disc = b * b - 4.0 * a * c;
        if (fabs(a) > CLSN_ROUND_EPS && disc > 0) {
            const double ds = sqrt(disc);
            r[0] = (-b + ds) / (2.0 * a);
            r[1] = (-b - ds) / (2.0 * a);
            set |= 3u;
        } else if (fabs(a) < CLSN_ROUND_EPS && fabs(b) > CLSN_ROUND_EPS) {
            r[0] = -c / b;
            set |= 1u;
        }
        tol = 0.0;
    }
    if (!(tol < 0.25 * dt)) return FAST_UNCERTAIN;
    // classify the roots like the reference's "-= MACH_EPS; outside [0, dt] -> -1; any > MACH_EPS"
    double tv[3];
    int nv = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (!((set >> i) & 1u)) continue;          // the reference keeps -1 there
        const double t = r[i] - CLSN_MACH_EPS;
        if (t < -tol || t > dt + tol) continue;    // certainly filtered out
        if (!(t > 2 * CLSN_MACH_EPS + tol && t < dt - tol)) return FAST_UNCERTAIN;
        tv[nv++] = t;
    }
    if (nv == 0) return FAST_MISS;
    double vmax = 0.0, xmax = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            vmax = fmax(vmax, fabs(q.av[i][k]));
            xmax = fmax(xmax, fabs(q.xo[i][k]));
        }
    xmax += dt * vmax;
    if (!(vmax < 1e100 && xmax < 1e100)) return FAST_UNCERTAIN;
    const double delta = tol * vmax + 4.0 * CLSN_MACH_EPS * xmax;
#pragma unroll 1
    for (int i = 0; i < nv; ++i) {
        double X[4][3];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int k = 0; k < 3; ++k) X[j][k] = q.xo[j][k] + tv[i] * q.av[j][k];
        const bool miss = edge ? ee_certain_miss(X, delta, h, xmax) : pt_certain_miss(X, delta, eps);
        if (!miss) return FAST_UNCERTAIN;
    }
    return FAST_DT_ONLY;
}

} // namespace clsn
