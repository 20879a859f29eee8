// Correctly rounded acos / cos / pow(x, 1.0/3.0) for the coplanarity cubic.
//
// The reference's isCoplanar (dcollid3d.cpp:435-444) takes these three from libm.  An edge-edge
// contact normal evaluated at a coplanarity root amplifies a 1-ulp change of the root to an O(1)
// change of the impulse, so "same result as the reference" needs these functions bit-identical,
// and no two libms agree bit for bit (glibc itself differs between its FMA and non-FMA ifunc
// variants).  The contract of this library is therefore the correctly rounded value: evaluated
// here in double-double arithmetic (error <= ~2^-90 relative, i.e. correct rounding except with
// probability ~2^-37 per call) and checked against binary128 in tests/test_crmath.py.
// glibc 2.39 agrees with the correctly rounded value on 99.9 % of arguments.
//
// The file is __host__ __device__ so the same code is unit-tested on the CPU; FMAs are explicit
// (the translation units that include it are built with --fmad=false / -ffp-contract=off).
#pragma once
#include <math.h>
#include "crmath_constants.inc"

#if defined(__CUDACC__)
#define CRM_HD __host__ __device__ __forceinline__
#else
#define CRM_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define CRM_FMA(a, b, c) __fma_rn((a), (b), (c))
#else
#define CRM_FMA(a, b, c) __builtin_fma((a), (b), (c))
#endif

namespace crm {

// sin(j/16), cos(j/16), j = 0..13, as double-doubles.  A namespace-scope table (global memory on the
// device, read through the L1) -- a function-local array would be rebuilt on the stack per call.
#if defined(__CUDACC__)
static __device__ const double d_sincos_tab[14][4] = {CRM_SINCOS_TABLE};
#endif
static const double h_sincos_tab[14][4] = {CRM_SINCOS_TABLE};

struct dd {
    double hi, lo;
};

CRM_HD dd two_sum(double a, double b)
{
    double s = a + b;
    double bb = s - a;
    double e = (a - (s - bb)) + (b - bb);
    return dd{s, e};
}
CRM_HD dd quick_two_sum(double a, double b) // |a| >= |b|
{
    double s = a + b;
    return dd{s, b - (s - a)};
}
CRM_HD dd two_prod(double a, double b)
{
    double p = a * b;
    return dd{p, CRM_FMA(a, b, -p)};
}
CRM_HD dd add(dd a, dd b)
{
    dd s = two_sum(a.hi, b.hi);
    dd t = two_sum(a.lo, b.lo);
    s.lo += t.hi;
    s = quick_two_sum(s.hi, s.lo);
    s.lo += t.lo;
    return quick_two_sum(s.hi, s.lo);
}
CRM_HD dd add_d(dd a, double b)
{
    dd s = two_sum(a.hi, b);
    s.lo += a.lo;
    return quick_two_sum(s.hi, s.lo);
}
CRM_HD dd mul(dd a, dd b)
{
    dd p = two_prod(a.hi, b.hi);
    p.lo += a.hi * b.lo + a.lo * b.hi;
    return quick_two_sum(p.hi, p.lo);
}
CRM_HD dd mul_d(dd a, double b)
{
    dd p = two_prod(a.hi, b);
    p.lo += a.lo * b;
    return quick_two_sum(p.hi, p.lo);
}
CRM_HD dd neg(dd a) { return dd{-a.hi, -a.lo}; }

// sin and cos of a double argument, |y| <= 6.5 (so that k*PIO2_1 is exact), as double-doubles.
// The cubic only needs [-2pi/3, pi].
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#else
static
#endif
void sincos_dd(double y, dd& sn, dd& cs)
{
#if defined(__CUDA_ARCH__)
    const double (*tab)[4] = d_sincos_tab;
#else
    const double (*tab)[4] = h_sincos_tab;
#endif
    // y = k*pi/2 + r, |r| <= pi/4 (+ a hair); pi/2 carried to ~155 bits
    double k = rint(y * CRM_2_OVER_PI);
    dd r = two_sum(y, -(k * CRM_PIO2_1)); // k*PIO2_1 is exact (51-bit constant, |k| small)
    r = add_d(r, -(k * CRM_PIO2_2));
    dd p3 = two_prod(k, CRM_PIO2_3);
    r = add(r, neg(p3));
    // r = j/16 + s, |s| <= 1/32
    double jf = rint(r.hi * 16.0);
    dd s = add_d(r, -(jf * 0.0625));
    int j = (int)jf;
    int ja = j < 0 ? -j : j;
    dd Sj = dd{tab[ja][0], tab[ja][1]};
    dd Cj = dd{tab[ja][2], tab[ja][3]};
    if (j < 0) Sj = neg(Sj);
    // sin(s), cos(s): leading terms in double-double, tails in double
    dd s2 = mul(s, s);
    dd s3 = mul(s2, s);
    dd s4 = mul(s2, s2);
    dd s5 = mul(s3, s2);
    double t = s2.hi;
    double s7 = s5.hi * t;
    double sin_tail = s7 * (-1.0 / 5040.0 + t * (1.0 / 362880.0 + t * (-1.0 / 39916800.0 + t * (1.0 / 6227020800.0))));
    double t3 = s4.hi * t;
    double cos_tail = t3 * (-1.0 / 720.0 + t * (1.0 / 40320.0 + t * (-1.0 / 3628800.0 + t * (1.0 / 479001600.0))));
    dd sin_s = add(s, add(mul(s3, dd{CRM_C3_HI, CRM_C3_LO}), add_d(mul(s5, dd{CRM_C5_HI, CRM_C5_LO}), sin_tail)));
    dd cos_s = add(dd{1.0, 0.0}, add(mul_d(s2, -0.5), add_d(mul(s4, dd{CRM_C4_HI, CRM_C4_LO}), cos_tail)));
    // angle addition
    dd sin_r = add(mul(Sj, cos_s), mul(Cj, sin_s));
    dd cos_r = add(mul(Cj, cos_s), neg(mul(Sj, sin_s)));
    int q = ((int)k) & 3;
    if (q == 0) { sn = sin_r; cs = cos_r; }
    else if (q == 1) { sn = cos_r; cs = neg(sin_r); }
    else if (q == 2) { sn = neg(sin_r); cs = neg(cos_r); }
    else { sn = neg(cos_r); cs = sin_r; }
}

// correctly rounded cos(y) for |y| <= 6.5.  Beyond that the host build falls back to libm; the device
// build has no fallback (keeps the kernels small): its callers -- the cubic (|y| <= pi) and the rigid
// body rotation angle dt*|w| -- stay far inside the range.
CRM_HD double cos_cr(double y)
{
#if !defined(__CUDA_ARCH__)
    if (!(fabs(y) <= 6.5)) return cos(y);
#endif
    dd s, c;
    sincos_dd(y, s, c);
    return c.hi + c.lo;
}
CRM_HD double sin_cr(double y)
{
#if !defined(__CUDA_ARCH__)
    if (!(fabs(y) <= 6.5)) return sin(y);
#endif
    dd s, c;
    sincos_dd(y, s, c);
    return s.hi + s.lo;
}

// correctly rounded acos(x), |x| <= 1: one Newton step on cos(theta) = x from the libm estimate
CRM_HD double acos_cr(double x)
{
    if (!(fabs(x) < 1.0)) {
        if (x == 1.0) return 0.0;
        if (x == -1.0) return 0x1.921fb54442d18p+1;
        return acos(x); // NaN for |x| > 1, as libm
    }
    double th0 = acos(x);
    dd s, c;
    sincos_dd(th0, s, c);
    dd e = add_d(c, -x);
    double corr = (e.hi + e.lo) / s.hi;
    return th0 + corr;
}

// correctly rounded pow(u, 1.0/3.0) for u >= 0 -- the exponent is the DOUBLE 1.0/3.0
// (= 1/3 - 2^-54/3), so the result is cbrt(u) * exp(delta * ln u), delta = CRM_POW13_DELTA
CRM_HD double pow13_cr(double u)
{
    // outside (1e-290, 1e300) and for 0 / inf / NaN: plain cbrt (exact for 0, inf, NaN; elsewhere within
    // 1.5e-14 relative -- such magnitudes do not occur for the cubic of a real mesh)
    if (!(u > 1.0e-290) || !(u < 1.0e300)) return cbrt(u);
    double c0 = cbrt(u);
    dd c2 = two_prod(c0, c0);
    dd c3 = mul_d(c2, c0);
    dd e = add_d(c3, -u);
    double corr = (e.hi + e.lo) / (3.0 * c2.hi); // Newton: cbrt(u) = c0 - corr
    double z = CRM_POW13_DELTA * log(u);
    double m = z + 0.5 * z * z; // exp(z) - 1
    return c0 + (c0 * m - corr);
}

} // namespace crm
