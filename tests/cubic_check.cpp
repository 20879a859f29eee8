// Host-side fuzz of collision_b200/csrc/cubic.cuh (the CUDA path's isCoplanar: coefficients, trig-free
// classifier, selective correctly rounded solve) against the oracle's is_coplanar (binary128 libm flavour).
// Built and run by tests/test_host_cpu.py.  Usage: cubic_check <liboracle.so> <cases> [seed]
// Prints: cases, mismatches (ret or any root bit), classifier rejections, of which oracle-true (must be 0), oracle-true
// cases, FP32 pre-filter rejections, of which the oracle keeps a root (must be 0).
#include <dlfcn.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "cubic.cuh"

typedef int (*orc_feature_t)(int, const double*, const double*, const double*, const unsigned char*, const double*, double,
                             double, const double*, double*, double*, double*);
typedef void (*orc_set_libm_t)(int);

static double U() { return drand48(); }
static double S() { return 2 * drand48() - 1; }

int main(int argc, char** argv)
{
    if (argc < 3) return 2;
    void* h = dlopen(argv[1], RTLD_NOW);
    if (!h) { fprintf(stderr, "%s\n", dlerror()); return 2; }
    orc_feature_t orc_feature = (orc_feature_t)dlsym(h, "orc_feature");
    orc_set_libm_t orc_set_libm = (orc_set_libm_t)dlsym(h, "orc_set_libm");
    orc_set_libm(1);  // ORC_LIBM_CR
    const long N = atol(argv[2]);
    srand48(argc > 3 ? atol(argv[3]) : 4242);
    const double params[6] = {1e-6, 1e-4, 1000, 0.01, 0.02, 0};
    const unsigned char flags[4] = {0, 0, 0, 0};
    const double mass[4] = {1, 1, 1, 1};
    long bad = 0, rejected = 0, rejected_wrong = 0, oracle_true = 0, pre = 0, pre_wrong = 0;
    for (long it = 0; it < N; ++it) {
        double x[4][3], v[4][3];
        const double L = 4e-3 * pow(10.0, (it % 5 == 0) ? 3 * S() : 0.0);   // element size, sometimes rescaled
        const double dt = (it % 7 == 0) ? pow(10.0, -3 + 3 * U()) : 1e-3;
        const double base[3] = {U(), U(), U()};
        for (int i = 0; i < 3; ++i)
            for (int k = 0; k < 3; ++k) x[i][k] = base[k] + L * S();
        // fourth point near the plane of the first three (or near their edge for the edge-edge flavour)
        double e1[3], e2[3], n[3];
        for (int k = 0; k < 3; ++k) { e1[k] = x[1][k] - x[0][k]; e2[k] = x[2][k] - x[0][k]; }
        n[0] = e1[1] * e2[2] - e1[2] * e2[1]; n[1] = e1[2] * e2[0] - e1[0] * e2[2]; n[2] = e1[0] * e2[1] - e1[1] * e2[0];
        double nm = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]) + 1e-300;
        const double w0 = 1.4 * U() - 0.2, w1 = 1.4 * U() - 0.2;
        const double gap = L * pow(10.0, -4 * U()) * (S() > 0 ? 1 : -1) * 0.25;
        for (int k = 0; k < 3; ++k) x[3][k] = x[0][k] + w0 * e1[k] + w1 * e2[k] + gap * n[k] / nm;
        const double speed = fabs(gap) / dt * pow(10.0, 1.5 * S());        // reaches the plane around t ~ dt
        const double jit = speed * pow(10.0, -6 * U());
        for (int i = 0; i < 4; ++i)
            for (int k = 0; k < 3; ++k) v[i][k] = jit * S();
        for (int k = 0; k < 3; ++k) v[3][k] -= (gap > 0 ? 1 : -1) * speed * n[k] / nm;
        switch (it % 11) {
        case 1: for (int i = 1; i < 4; ++i) for (int k = 0; k < 3; ++k) v[i][k] = v[0][k]; break;        // rigid translation
        case 2: for (int k = 0; k < 3; ++k) { v[1][k] = v[0][k]; v[2][k] = v[0][k]; } break;             // rigid triangle
        case 3: for (int k = 0; k < 3; ++k) x[3][k] = x[0][k] + w0 * e1[k] + w1 * e2[k]; break;          // starts coplanar
        case 4: for (int i = 0; i < 4; ++i) for (int k = 0; k < 3; ++k) v[i][k] *= 1e3; break;           // fast
        case 5: for (int i = 0; i < 4; ++i) for (int k = 0; k < 3; ++k) v[i][k] *= 1e-4; break;          // slow
        case 6: for (int k = 0; k < 3; ++k) v[3][k] = -v[3][k]; break;                                   // separating
        case 7: for (int k = 0; k < 3; ++k) { v[1][k] = -v[3][k]; v[2][k] = 0.5 * v[3][k]; } break;      // tumbling triangle
        default: break;
        }
        clsn::Quad q;
        double xo[12], av[12];
        for (int i = 0; i < 4; ++i)
            for (int k = 0; k < 3; ++k) { q.xo[i][k] = xo[3 * i + k] = x[i][k]; q.av[i][k] = av[3 * i + k] = v[i][k]; }
        double r_or[4] = {-1, -1, -1, dt}, acc[40], hit;
        const int ret_or = orc_feature(0, xo, xo, av, flags, mass, 1e-6, dt, params, r_or, acc, &hit);
        double r_me[3] = {-1, -1, -1};
        const int ret_me = clsn::is_coplanar<true>(q, dt, r_me) ? 1 : 0;
        double a, b, c, d;
        clsn::coplanar_coeffs(q, a, b, c, d);
        const int kind = clsn::coplanar_maybe(a, b, c, d, dt);
        oracle_true += ret_or;
        if (clsn::coplanar_prefilter32(q, dt)) {
            // the FP32 pre-filter may only reject what the oracle rejects
            ++pre;
            if (ret_or || r_or[0] >= 0 || r_or[1] >= 0 || r_or[2] >= 0) ++pre_wrong;
        }
        if (kind == 0) {
            ++rejected;
            // a rejected cubic must have NO root surviving the [0, dt] filter in the oracle
            if (ret_or || r_or[0] >= 0 || r_or[1] >= 0 || r_or[2] >= 0) ++rejected_wrong;
        } else {
            // roots of the selective correctly rounded solve == the oracle's, bit for bit
            if (ret_me != ret_or || memcmp(r_me, r_or, 3 * sizeof(double)) != 0) {
                if (bad < 5)
                    fprintf(stderr, "case %ld: ret %d/%d roots %.17g %.17g %.17g | %.17g %.17g %.17g\n", it, ret_me, ret_or, r_me[0],
                            r_me[1], r_me[2], r_or[0], r_or[1], r_or[2]);
                ++bad;
            }
        }
    }
    printf("%ld %ld %ld %ld %ld %ld %ld\n", N, bad, rejected, rejected_wrong, oracle_true, pre, pre_wrong);
    return 0;
}
