// Host-side fuzz of collision_b200/csrc/fastpath.cuh (the CCD feature fast path) against the oracle's
// MovingPointToTri / MovingEdgeToEdge (orc_feature kinds 3 / 4, correctly rounded libm flavour).
// The fast path's libm is deliberately perturbed by up to +-4 ulp per call, so the check does not depend
// on which libm (host glibc, CUDA) evaluates it.
// Built and run by tests/test_host_cpu.py.  Usage: fastpath_check <liboracle.so> <cases> [seed]
// Prints: cases  wrong  n_miss  n_dt_only  n_uncertain  oracle_hits  oracle_hits_at_root
//   wrong = FAST_MISS although the oracle's isCoplanar is true, or FAST_DT_ONLY although the oracle's
//           isCoplanar is false or it fires before dt.  Must be 0.
#include <dlfcn.h>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

static double perturb_ulps(double y)
{
    int k = (int)(lrand48() % 9) - 4;
    for (; k > 0; --k) y = nextafter(y, INFINITY);
    for (; k < 0; ++k) y = nextafter(y, -INFINITY);
    return y;
}
#define CLSN_FAST_ACOS(x) perturb_ulps(acos(x))
#define CLSN_FAST_COS(x) perturb_ulps(cos(x))
#define CLSN_FAST_CBRT(x) perturb_ulps(cbrt(x))
#include "fastpath.cuh"

typedef int (*orc_feature_t)(int, const double*, const double*, const double*, const unsigned char*, const double*, double,
                             double, const double*, double*, double*, double*);
typedef void (*orc_set_libm_t)(int);

static double U() { return drand48(); }
static double S() { return 2 * drand48() - 1; }
static double sgn() { return drand48() < 0.5 ? -1.0 : 1.0; }

static void cross(const double* a, const double* b, double* r)
{
    r[0] = a[1] * b[2] - a[2] * b[1]; r[1] = a[2] * b[0] - a[0] * b[2]; r[2] = a[0] * b[1] - a[1] * b[0];
}

// Scene-level check (called through ctypes by tests/test_host_cpu.py): every CCD feature test of every
// non-adjacent candidate pair of a pass, on the state the oracle had before that pass.
// elem: 4 ints per element (p0, p1, p2 or -1 for a bond, unused).  counts: miss, dt_only, uncertain, oracle hits,
// classifier rejections (coplanar_maybe == 0), FP32 pre-filter rejections; both kinds of rejection are checked against
// the oracle's roots as well.
extern "C" long fastpath_scene_check(const char* oracle_so, long npairs, const int* pairs, const int* elem, const double* xo,
                                     const double* av, double dt, double eps, long* counts)
{
    void* hnd = dlopen(oracle_so, RTLD_NOW);
    if (!hnd) return -1;
    orc_feature_t orc_feature = (orc_feature_t)dlsym(hnd, "orc_feature");
    orc_set_libm_t orc_set_libm = (orc_set_libm_t)dlsym(hnd, "orc_set_libm");
    orc_set_libm(1);
    static const int tt[15][4] = {{0, 1, 2, 3}, {0, 1, 2, 4}, {0, 1, 2, 5}, {3, 4, 5, 0}, {3, 4, 5, 1}, {3, 4, 5, 2},
                                  {0, 1, 3, 4}, {0, 1, 4, 5}, {0, 1, 5, 3}, {1, 2, 3, 4}, {1, 2, 4, 5}, {1, 2, 5, 3},
                                  {2, 0, 3, 4}, {2, 0, 4, 5}, {2, 0, 5, 3}};
    static const int tb[5][4] = {{0, 1, 2, 3}, {0, 1, 2, 4}, {0, 1, 3, 4}, {1, 2, 3, 4}, {2, 0, 3, 4}};
    const double params[6] = {eps, 1e-4, 1000, 0.01, 0.02, 0};
    const unsigned char flags[4] = {0, 0, 0, 0};
    const double mass[4] = {1, 1, 1, 1};
    long wrong = 0;
    for (int i = 0; i < 6; ++i) counts[i] = 0;
    for (long pi = 0; pi < npairs; ++pi) {
        int ea = pairs[2 * pi], eb = pairs[2 * pi + 1];
        if (ea > eb) { int t = ea; ea = eb; eb = t; }
        const int* A = elem + 4 * ea;
        const int* B = elem + 4 * eb;
        if (A[2] < 0 && B[2] >= 0) { const int* t = A; A = B; B = t; }   // (tri, bond)
        const int ids[6] = {A[0], A[1], A[2], B[0], B[1], B[2]};
        bool adjacent = false;
        for (int i = 0; i < 3; ++i)
            for (int j = 3; j < 6; ++j)
                if (ids[i] >= 0 && ids[j] >= 0 && ids[i] == ids[j]) adjacent = true;
        if (adjacent) continue;
        const int nf = (A[2] >= 0 && B[2] >= 0) ? 15 : (A[2] >= 0 ? 5 : 1);
        for (int f = 0; f < nf; ++f) {
            int sl[4];
            bool edge;
            if (nf == 15) { for (int q = 0; q < 4; ++q) sl[q] = tt[f][q]; edge = f >= 6; }
            else if (nf == 5) { for (int q = 0; q < 4; ++q) sl[q] = tb[f][q]; edge = f >= 2; }
            else { sl[0] = 0; sl[1] = 1; sl[2] = 3; sl[3] = 4; edge = true; }
            clsn::Quad q;
            double qx[12], qv[12];
            for (int i = 0; i < 4; ++i)
                for (int k = 0; k < 3; ++k) {
                    q.xo[i][k] = qx[3 * i + k] = xo[3 * ids[sl[i]] + k];
                    q.av[i][k] = qv[3 * i + k] = av[3 * ids[sl[i]] + k];
                }
            double r_or[4] = {-1, -1, -1, dt}, acc[40], hit_root = -1;
            const int ret_or = orc_feature(edge ? 4 : 3, qx, qx, qv, flags, mass, eps, dt, params, r_or, acc, &hit_root);
            const bool cop = r_or[0] > DBL_EPSILON || r_or[1] > DBL_EPSILON || r_or[2] > DBL_EPSILON;
            if (ret_or > 0) ++counts[3];
            const bool any_root = r_or[0] >= 0 || r_or[1] >= 0 || r_or[2] >= 0;
            double ca, cb, cc, cd;
            clsn::coplanar_coeffs(q, ca, cb, cc, cd);
            if (clsn::coplanar_maybe(ca, cb, cc, cd, dt) == 0) { ++counts[4]; if (any_root) ++wrong; }
            if (clsn::coplanar_prefilter32(q, dt)) { ++counts[5]; if (any_root) ++wrong; }
            const int st = clsn::feature_fast(q, edge, dt, eps, eps);
            ++counts[st];
            if (st == clsn::FAST_MISS && (cop || ret_or != 0)) ++wrong;
            if (st == clsn::FAST_DT_ONLY && (!cop || ret_or < 0 || (ret_or > 0 && hit_root != dt))) ++wrong;
        }
    }
    return wrong;
}

int main(int argc, char** argv)
{
    if (argc < 3) return 2;
    void* hnd = dlopen(argv[1], RTLD_NOW);
    if (!hnd) { fprintf(stderr, "%s\n", dlerror()); return 2; }
    orc_feature_t orc_feature = (orc_feature_t)dlsym(hnd, "orc_feature");
    orc_set_libm_t orc_set_libm = (orc_set_libm_t)dlsym(hnd, "orc_set_libm");
    orc_set_libm(1);  // ORC_LIBM_CR
    const long N = atol(argv[2]);
    srand48(argc > 3 ? atol(argv[3]) : 99);
    const double eps = 1e-6;
    const double params[6] = {eps, 1e-4, 1000, 0.01, 0.02, 0};
    const unsigned char flags[4] = {0, 0, 0, 0};
    const double mass[4] = {1, 1, 1, 1};
    long wrong = 0, n_miss = 0, n_dt = 0, n_unc = 0, hits = 0, hits_root = 0;
    for (long it = 0; it < N; ++it) {
        const bool edge = (it & 1) != 0;
        const double L = 4e-3 * pow(10.0, (it % 5 == 0) ? 2 * S() : 0.0);
        const double dt = (it % 7 == 0) ? pow(10.0, -3 + 2 * U()) : 1e-3;
        double x[4][3], v[4][3];
        const double base[3] = {U(), U(), U()};
        const double speed = L / dt * pow(10.0, -2.5 + 2.5 * U());
        for (int i = 0; i < 4; ++i)
            for (int k = 0; k < 3; ++k) { x[i][k] = base[k] + L * S(); v[i][k] = speed * S(); }
        // configuration at a chosen time tc at which the four points are (nearly) coplanar, with chosen
        // barycentric coordinates (point-triangle) or line parameters (edge-edge) -- often right at the
        // accept/reject boundary of the static test
        const double tc = dt * ((it % 13 == 0) ? 1.2 * U() : U());
        double Y[4][3];
        for (int i = 0; i < 4; ++i)
            for (int k = 0; k < 3; ++k) Y[i][k] = x[i][k] + tc * v[i][k];
        const int mode = (int)(it % 6);
        double gapmag = (mode == 5) ? eps * pow(10.0, 1.5 * S()) : 0.0;   // off-plane offset at tc
        if (!edge) {
            double e1[3], e2[3], n[3];
            for (int k = 0; k < 3; ++k) { e1[k] = Y[0][k] - Y[2][k]; e2[k] = Y[1][k] - Y[2][k]; }
            cross(e1, e2, n);
            const double nm = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]) + 1e-300;
            double w0 = 1.4 * U() - 0.2, w1 = 1.4 * U() - 0.2;
            if (mode == 1) w0 = -eps + sgn() * pow(10.0, -6 - 8 * U());            // at the lower boundary
            if (mode == 2) w1 = 1.0 + eps + sgn() * pow(10.0, -6 - 8 * U()) - w0;   // w2 at the lower boundary... (w2 = 1-w0-w1)
            if (mode == 3) { w0 = 1.0 + eps + sgn() * pow(10.0, -6 - 8 * U()); w1 = 1e-3 * S(); }
            if (mode == 4) { w0 = 0.3; w1 = -eps + sgn() * pow(10.0, -9 - 6 * U()); }
            for (int k = 0; k < 3; ++k) {
                const double Pk = Y[2][k] + w0 * e1[k] + w1 * e2[k] + gapmag * n[k] / nm;
                x[3][k] = Pk - tc * v[3][k];
            }
        } else {
            double d1[3], d2[3], n[3];
            for (int k = 0; k < 3; ++k) { d1[k] = Y[1][k] - Y[0][k]; d2[k] = Y[3][k] - Y[2][k]; }
            cross(d1, d2, n);
            const double nm = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]) + 1e-300;
            const double l1 = sqrt(d1[0] * d1[0] + d1[1] * d1[1] + d1[2] * d1[2]) + 1e-300;
            double a = 1.4 * U() - 0.2, b = 1.4 * U() - 0.2;
            // crossing just outside the segment: the clamped distance is (a - 1) * |d1| ~ eps
            if (mode == 1) a = 1.0 + eps * (1.0 + sgn() * pow(10.0, -8 * U())) / l1;
            if (mode == 2) a = -eps * (1.0 + sgn() * pow(10.0, -8 * U())) / l1;
            if (mode == 3) { b = 1.0 + eps * (1.0 + sgn() * pow(10.0, -8 * U())) / (sqrt(d2[0] * d2[0] + d2[1] * d2[1] + d2[2] * d2[2]) + 1e-300); }
            if (mode == 4) { a = U(); b = U(); gapmag = eps * (1.0 + sgn() * pow(10.0, -8 * U())); }
            // move edge 2-3 rigidly so that Y2 + b d2 = Y0 + a d1 + gap n
            double shift[3];
            for (int k = 0; k < 3; ++k) shift[k] = (Y[0][k] + a * d1[k] + gapmag * n[k] / nm) - (Y[2][k] + b * d2[k]);
            for (int k = 0; k < 3; ++k) { x[2][k] += shift[k]; x[3][k] += shift[k]; }
        }
        switch (it % 11) {
        case 1: for (int i = 0; i < 4; ++i) for (int k = 0; k < 3; ++k) v[i][k] *= 1e2; break;   // fast (roots move)
        case 2: for (int k = 0; k < 3; ++k) { v[1][k] = v[0][k]; v[2][k] = v[0][k]; } break;     // rigid first three
        case 3: for (int i = 0; i < 4; ++i) for (int k = 0; k < 3; ++k) v[i][k] *= 1e-3; break;  // slow
        case 4: for (int i = 1; i < 4; ++i) for (int k = 0; k < 3; ++k) v[i][k] = v[0][k] + 1e-7 * speed * S(); break; // nearly rigid
        default: break;
        }
        clsn::Quad q;
        double xo[12], av[12];
        for (int i = 0; i < 4; ++i)
            for (int k = 0; k < 3; ++k) { q.xo[i][k] = xo[3 * i + k] = x[i][k]; q.av[i][k] = av[3 * i + k] = v[i][k]; }
        double r_or[4] = {-1, -1, -1, dt}, acc[40], hit_root = -1;
        const int ret_or = orc_feature(edge ? 4 : 3, xo, xo, av, flags, mass, eps, dt, params, r_or, acc, &hit_root);
        const bool cop = r_or[0] > DBL_EPSILON || r_or[1] > DBL_EPSILON || r_or[2] > DBL_EPSILON;
        if (ret_or > 0) { ++hits; if (hit_root != dt) ++hits_root; }
        const int st = clsn::feature_fast(q, edge, dt, eps, eps);
        bool bad = false;
        if (st == clsn::FAST_MISS) { ++n_miss; bad = cop || ret_or != 0; }
        else if (st == clsn::FAST_DT_ONLY) { ++n_dt; bad = !cop || ret_or < 0 || (ret_or > 0 && hit_root != dt); }
        else ++n_unc;
        if (bad) {
            if (wrong < 5)
                fprintf(stderr, "case %ld (%s): fast %d, oracle ret %d cop %d hit_root %.17g roots %.17g %.17g %.17g dt %.17g\n", it,
                        edge ? "ee" : "pt", st, ret_or, (int)cop, hit_root, r_or[0], r_or[1], r_or[2], dt);
            ++wrong;
        }
    }
    printf("%ld %ld %ld %ld %ld %ld %ld\n", N, wrong, n_miss, n_dt, n_unc, hits, hits_root);
    return 0;
}
