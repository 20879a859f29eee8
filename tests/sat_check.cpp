// Host-side fuzz of sat_normal_far() (collision_b200/csrc/cubic.cuh: the FP32 separating-axis test along the feature's
// own normal that k_cull puts in front of the proximity narrow phase) against the oracle's PointToTri / EdgeToEdge
// (static, proximity thickness) and MovingPointToTri / MovingEdgeToEdge (moving, rounding tolerance): whatever the test
// rejects, the oracle must not report.  Half of the cases sit within a few contact distances of the accept/reject
// boundary.  Built and run by tests/test_host_cpu.py.  Usage: sat_check <liboracle.so> <cases> [seed]
// Prints: cases, rejected (static), of which oracle hits (must be 0), oracle hits (static), rejected (moving), of which
// oracle hits (must be 0), oracle hits (moving).
#include <dlfcn.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
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
    orc_set_libm(1);
    const long N = atol(argv[2]);
    srand48(argc > 3 ? atol(argv[3]) : 99);
    const double eps = 1e-6, thickness = 1e-4, dt = 1e-3;
    const double params[6] = {eps, thickness, 1000, 0.01, 0.02, 0};
    const unsigned char flags[4] = {0, 0, 0, 0};
    const double mass[4] = {1, 1, 1, 1};
    long rej_s = 0, rej_s_wrong = 0, hit_s = 0, rej_m = 0, rej_m_wrong = 0, hit_m = 0;
    for (long it = 0; it < N; ++it) {
        const bool edge = it & 1;
        const double L = 4e-3 * pow(10.0, (it % 9 == 0) ? 2.5 * S() : 0.0);
        double x[4][3], v[4][3];
        const double base[3] = {U(), U(), U()};
        for (int i = 0; i < 3; ++i)
            for (int k = 0; k < 3; ++k) x[i][k] = base[k] + L * S();
        if (it % 13 == 0)   // sliver triangle / nearly parallel edges
            for (int k = 0; k < 3; ++k) x[2][k] = x[0][k] + (x[1][k] - x[0][k]) * U() + 1e-6 * L * S();
        double e1[3], e2[3], n[3];
        for (int k = 0; k < 3; ++k) { e1[k] = x[1][k] - x[0][k]; e2[k] = x[2][k] - x[0][k]; }
        n[0] = e1[1] * e2[2] - e1[2] * e2[1]; n[1] = e1[2] * e2[0] - e1[0] * e2[2]; n[2] = e1[0] * e2[1] - e1[1] * e2[0];
        const double nm = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]) + 1e-300;
        // distance of the second sub-feature from the first: around the contact distance in half of the cases
        const double hh = (it % 4 < 2) ? thickness : eps;
        const double gap = (it % 2 ? 1 : -1) * ((it % 3 == 0) ? hh * (0.5 + 1.5 * U()) : hh * pow(10.0, 3 * U()));
        const double w0 = 1.3 * U() - 0.15, w1 = 1.3 * U() - 0.15;
        double foot[3];
        for (int k = 0; k < 3; ++k) foot[k] = x[0][k] + w0 * e1[k] + w1 * e2[k] + gap * n[k] / nm;
        if (!edge) {
            for (int k = 0; k < 3; ++k) x[3][k] = foot[k];
        } else {
            // second edge through `foot`, roughly in a plane parallel to edge 0-1: points 2, 3
            double d[3] = {S(), S(), S()};
            for (int k = 0; k < 3; ++k) { x[2][k] = foot[k] - 0.5 * L * d[k]; x[3][k] = foot[k] + 0.5 * L * d[k]; }
        }
        const double speed = fabs(gap) / dt * pow(10.0, 1.5 * S());
        for (int i = 0; i < 4; ++i)
            for (int k = 0; k < 3; ++k) v[i][k] = speed * 0.05 * S();
        const int mov0 = edge ? 2 : 3;
        for (int i = mov0; i < 4; ++i)
            for (int k = 0; k < 3; ++k) v[i][k] -= (gap > 0 ? 1 : -1) * speed * n[k] / nm * (it % 5 == 0 ? -1 : 1);
        clsn::Quad q;
        double xo[12], av[12];
        for (int i = 0; i < 4; ++i)
            for (int k = 0; k < 3; ++k) { q.xo[i][k] = xo[3 * i + k] = x[i][k]; q.av[i][k] = av[3 * i + k] = v[i][k]; }
        double roots[4], acc[40], hit;
        const int ret_s = orc_feature(edge ? 2 : 1, xo, xo, av, flags, mass, thickness, dt, params, roots, acc, &hit);
        const int ret_m = orc_feature(edge ? 4 : 3, xo, xo, av, flags, mass, eps, dt, params, roots, acc, &hit);
        hit_s += ret_s > 0;
        hit_m += ret_m > 0;
        if (clsn::sat_normal_far(q, edge, false, dt, thickness, eps)) { ++rej_s; if (ret_s > 0) ++rej_s_wrong; }
        if (clsn::sat_normal_far(q, edge, true, dt, eps, eps)) { ++rej_m; if (ret_m > 0) ++rej_m_wrong; }
    }
    printf("%ld %ld %ld %ld %ld %ld %ld\n", N, rej_s, rej_s_wrong, hit_s, rej_m, rej_m_wrong, hit_m);
    return 0;
}
