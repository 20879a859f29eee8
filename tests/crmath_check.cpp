// Host-side check of collision_b200/csrc/crmath.cuh against binary128 (libquadmath).
// Built and run by tests/test_crmath.py:  prints "<name> <samples> <mismatches>" per function.
#include <quadmath.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include "crmath.cuh"

int main(int argc, char** argv)
{
    long N = argc > 1 ? atol(argv[1]) : 1000000;
    srand48(argc > 2 ? atol(argv[2]) : 20241017);
    long bad_cos = 0, bad_sin = 0, bad_acos = 0, bad_pow = 0, bad_edge = 0;
    for (long i = 0; i < N; i++) {
        double y = (drand48() * 2 - 1) * 3.3;
        if (i % 7 == 0) y = (drand48() * 2 - 1) * 6.4;           // the whole supported range
        if (i % 11 == 0) y = M_PI / 2 * (double)(lrand48() % 9 - 4) + (drand48() - 0.5) * 1e-9; // near zeros
        double x = drand48() * 2 - 1;
        if (i % 5 == 0) x = (lrand48() & 1 ? 1 : -1) * (1.0 - drand48() * 1e-9);                // near +-1
        if (i % 13 == 0) x = (drand48() - 0.5) * 1e-9;                                           // near 0
        double u = exp((drand48() * 2 - 1) * 80);
        if (crm::cos_cr(y) != (double)cosq((__float128)y)) bad_cos++;
        if (crm::sin_cr(y) != (double)sinq((__float128)y)) bad_sin++;
        if (crm::acos_cr(x) != (double)acosq((__float128)x)) bad_acos++;
        if (crm::pow13_cr(u) != (double)powq((__float128)u, (__float128)(1.0 / 3.0))) bad_pow++;
    }
    const double edge[] = {1.0, -1.0, 0.0, -0.0, 0.5, -0.5, 0x1.fffffffffffffp-1, -0x1.fffffffffffffp-1};
    for (double x : edge)
        if (crm::acos_cr(x) != (double)acosq((__float128)x)) bad_edge++;
    const double uedge[] = {0.0, 1.0, 8.0, 27.0, 1e-289, 9e299, 0.001};
    for (double u : uedge)
        if (crm::pow13_cr(u) != (double)powq((__float128)u, (__float128)(1.0 / 3.0))) bad_edge++;
    printf("cos %ld %ld\nsin %ld %ld\nacos %ld %ld\npow13 %ld %ld\nedge 15 %ld\n", N, bad_cos, N, bad_sin, N, bad_acos, N,
           bad_pow, bad_edge);
    return 0;
}
