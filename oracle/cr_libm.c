/* Correctly rounded acos / cos / sin / pow for the SECOND build of the reference -- TEST INFRASTRUCTURE.
 *
 * oracle/_ref/libcollision_ref_cr.so is the same unmodified /root/reference sources as libcollision_ref.so, linked
 * with -Wl,-Bsymbolic against these definitions instead of the host libm's, i.e. "the reference on a platform whose
 * libm is correctly rounded".  The reference takes acos/cos/pow (isCoplanar, dcollid3d.cpp:436-444) and sin/cos
 * (updateImpactListVelocity, :163-169) from whatever libm it is linked with; glibc's are within 1 ulp but not
 * correctly rounded (~0.1 % of arguments differ) and vary with the CPU (ifunc FMA variants).  The CUDA path is held
 * to the correctly rounded values (collision_b200/csrc/crmath.cuh), so against THIS build its times of impact can be
 * compared bit for bit; against the native-libm build the libm flavour shows up as <= 1e-10 relative noise on the
 * ill-conditioned trig form of the cubic root.  Values: binary128 libquadmath, rounded once to double. */
#include <quadmath.h>

double acos(double x) { return (double)acosq((__float128)x); }
double cos(double x) { return (double)cosq((__float128)x); }
double sin(double x) { return (double)sinq((__float128)x); }
double pow(double x, double y) { return (double)powq((__float128)x, (__float128)y); }
void sincos(double x, double* s, double* c)
{
    *s = (double)sinq((__float128)x);
    *c = (double)cosq((__float128)x);
}
