/* CPU restatement of antdvid/Collision's per-step pipeline -- TEST INFRASTRUCTURE.
 * See collision_oracle.h for the rules (who may call this, canonical order, parity pinning).
 * All citations are file:line into the reference.  Arithmetic keeps the reference's operation
 * order exactly (Dot3d = a0*b0 + a1*b1 + a2*b2 left to right, no FMA contraction: build with
 * -ffp-contract=off and no -march), because contact decisions sit on thresholds.
 */
#include "collision_oracle.h"

#include <float.h>
#include <math.h>
#include <quadmath.h>
#include <stdlib.h>
#include <string.h>

#define MACH_EPS DBL_EPSILON /* FronTier's MACH_EPS, SURVEY 8(c) */
#define ROUND_EPS 1e-10      /* collid.h:17 */
#define BOX_PAD 1e-6         /* AABB.cpp:10-16, fixed regardless of setRoundingTolerance */
#define MAX_ITER 5           /* dcollid.cpp:433 */

#define F_FIXED 1
#define F_MOVABLE 2

struct orc_ctx {
    int V, T, B, N, nhs;
    int *tri, *tri_surf, *bond;
    unsigned char* flags;
    int* vhs;
    double* hs_mass;
    double eps, thickness, k, m, lambda, cr, dt;
    double lo[3], hi[3];
    double *xo, *x, *av;   /* x_old, Coords (candidate / final), avgVel: 3V each */
    double *imp, *fric;    /* 3V */
    int* cnt;              /* V */
    unsigned char* has;    /* V */
    double* imp_rg;        /* 3*nhs: collsnImpulse_RG, identical on every point of a body */
    int* cnt_rg;           /* nhs */
    /* union-find lists of createImpZoneForRG (dcollid3d.cpp:54-68), topology only */
    int *uf_root, *uf_next, *uf_tail, *uf_weight;
    int uf_ready;
    int det_imp_zone;  /* s_detImpZone, dcollid.cpp:28 */
    int impact_zones;  /* run computeImpactZone from orc_resolve when passes are exhausted */
    unsigned char* sorted;
    double* tri_len0;  /* TRI::side_length0[3] */
    double* bond_len0; /* BOND::length0 */
    int have_len0, strain_limiting;
    /* results of the last detect */
    double* mrg_com;          /* 3*nhs: CollisionSolver::mrg_com (collid.h:174), lives across steps */
    unsigned char* mrg_valid; /* nhs */
    int* cand; long n_cand, cap_cand;
    int* truep; long n_true, cap_true;
    orc_contact* con; long n_con, cap_con;
    int error;
};

/* ------------------------------------------------------------------ libm flavour
 * isCoplanar calls acos, cos and pow(x, 1.0/3.0) (dcollid3d.cpp:435-444).  The reference gets them
 * from whatever libm it is linked with; glibc's results are within 1 ulp but not always correctly
 * rounded (measured here: ~0.1 % of arguments) and depend on the CPU (ifunc FMA variants), and the
 * edge-edge contact normal at a coplanarity root amplifies a 1-ulp root change to O(1).
 *   ORC_LIBM_NATIVE  : the host libm, exactly as the reference -- used to pin this restatement
 *                      bit for bit against the compiled reference (oracle/_ref);
 *   ORC_LIBM_CR      : correctly rounded values (binary128 libquadmath, rounded once) -- the
 *                      platform-independent definition the CUDA path is held to bit for bit. */
static int g_libm = 0;
void orc_set_libm(int mode) { g_libm = mode; }
int orc_get_libm(void) { return g_libm; }
static double m_acos(double x) { return g_libm ? (double)acosq((__float128)x) : acos(x); }
static double m_cos(double x) { return g_libm ? (double)cosq((__float128)x) : cos(x); }
static double m_sin(double x) { return g_libm ? (double)sinq((__float128)x) : sin(x); }
static double m_pow13(double x)
{
    return g_libm ? (double)powq((__float128)x, (__float128)(1.0 / 3.0)) : pow(x, 1.0 / 3.0);
}

/* ------------------------------------------------------------------ small vector helpers */
static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static double mag3(const double* a) { return sqrt(dot3(a, a)); }
static void cross3(const double* b, const double* c, double* r)
{
    r[0] = b[1] * c[2] - b[2] * c[1];
    r[1] = b[2] * c[0] - b[0] * c[2];
    r[2] = b[0] * c[1] - b[1] * c[0];
}
static void sub3(const double* a, const double* b, double* r)
{
    r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2];
}
static double dmin(double a, double b) { return b < a ? b : a; } /* std::min(a,b) */
static double dmax(double a, double b) { return a < b ? b : a; } /* std::max(a,b) */

static int is_static(const orc_ctx* c, int p) { return (c->flags[p] & F_FIXED) != 0; }    /* dcollid.cpp:1069 */
static int is_movable(const orc_ctx* c, int p) { return (c->flags[p] & F_MOVABLE) != 0; } /* dcollid.cpp:1081 */
static int is_rigid(const orc_ctx* c, int p) { return (c->flags[p] & (F_FIXED | F_MOVABLE)) != 0; }

static void push_contact(orc_ctx* c, const orc_contact* k)
{
    if (c->n_con == c->cap_con) {
        c->cap_con = c->cap_con ? 2 * c->cap_con : 1024;
        c->con = (orc_contact*)realloc(c->con, (size_t)c->cap_con * sizeof(orc_contact));
    }
    c->con[c->n_con++] = *k;
}

/* SpreadImpactZoneImpulse, dcollid.cpp:1101-1115.  Every point of a movable body is in one
 * union-find list (createImpZoneForRG), so "add to each point of the list" is one per-body add. */
static void spread_rg(orc_ctx* c, int p, double impulse, const double* nor)
{
    int b = c->vhs[p];
    for (int i = 0; i < 3; ++i) c->imp_rg[3 * b + i] += impulse * nor[i];
    c->cnt_rg[b] += 1;
}

/* ------------------------------------------------------------------ impulses */
/* PointToTriImpulse, dcollid3d.cpp:925-1107.  p[0..2] triangle, p[3] point. */
static void point_to_tri_impulse(orc_ctx* c, const int* p, const double* nor, double* w, double dist)
{
    const double* v[4];
    for (int i = 0; i < 4; ++i) v[i] = c->av + 3 * p[i];
    double v_rel[3] = {0.0, 0.0, 0.0}, vn, vt;
    double impulse = 0.0, m_impulse, sum_w = 0.0;
    double rigid_impulse[2] = {0.0, 0.0};
    double k = c->k, m = c->m, dt = c->dt, lambda = c->lambda, h = c->thickness, cr = c->cr;
    dist = h - dist;                                               /* :942 */
    for (int i = 0; i < 3; ++i) {                                  /* :946-951 */
        v_rel[i] += v[3][i];
        for (int j = 0; j < 3; ++j) v_rel[i] -= w[j] * v[j][i];
    }
    vn = dot3(v_rel, nor);
    if (dot3(v_rel, v_rel) > vn * vn) vt = sqrt(dot3(v_rel, v_rel) - vn * vn);
    else vt = 0.0;
    if (vn < 0) {                                                  /* :957-995 */
        if (is_static(c, p[3]) || (is_static(c, p[0]) && is_static(c, p[1]) && is_static(c, p[2]))) {
            impulse = vn; rigid_impulse[0] = vn; rigid_impulse[1] = vn;
        } else if (is_movable(c, p[0]) && is_movable(c, p[1]) && is_movable(c, p[2]) && is_movable(c, p[3])) {
            double m1 = c->hs_mass[c->vhs[p[0]]], m2 = c->hs_mass[c->vhs[p[3]]];
            rigid_impulse[0] = vn * m2 / (m1 + m2);
            rigid_impulse[1] = vn * m1 / (m1 + m2);
        } else if (is_movable(c, p[0]) && is_movable(c, p[1]) && is_movable(c, p[2])) {
            rigid_impulse[0] = 0.5 * vn; impulse = 0.5 * vn;
        } else if (is_movable(c, p[3])) {
            impulse = 0.5 * vn; rigid_impulse[1] = 0.5 * vn;
        } else
            impulse = vn * 0.5;
        for (int i = 0; i < 3; ++i) {
            if (is_static(c, p[i])) w[i] = 0.0;
            sum_w += w[i];
        }
        if (fabs(sum_w) > MACH_EPS) {
            double s = 1.0 / sum_w;
            for (int i = 0; i < 3; ++i) w[i] = s * w[i];
        }
    }
    int all_rigid = is_rigid(c, p[0]) && is_rigid(c, p[1]) && is_rigid(c, p[2]) && is_rigid(c, p[3]);
    if (vn * dt < 0.1 * dist) {                                    /* :996-1011 */
        if (all_rigid) {
            rigid_impulse[0] *= 1.0 + cr; rigid_impulse[1] *= 1.0 + cr;
        } else {
            double tmp = -dmin(dt * k * dist / m, (0.1 * dist / dt - vn));
            impulse += tmp; rigid_impulse[0] += tmp; rigid_impulse[1] += tmp;
        }
    }
    if (fabs(sum_w) < MACH_EPS) m_impulse = impulse;               /* :1012-1015 */
    else m_impulse = 2.0 * impulse / (1.0 + dot3(w, w));
    if (all_rigid) {                                               /* :1044-1052 */
        if (is_movable(c, p[0])) spread_rg(c, p[0], rigid_impulse[0], nor);
        if (is_movable(c, p[3])) spread_rg(c, p[3], -1.0 * rigid_impulse[1], nor);
        return;
    }
    for (int i = 0; i < 3; ++i) {                                  /* :1053-1066 */
        if (is_static(c, p[i])) continue;
        double t_impulse = m_impulse;
        if (is_movable(c, p[i])) t_impulse = rigid_impulse[0];
        for (int j = 0; j < 3; ++j) {
            c->imp[3 * p[i] + j] += w[i] * t_impulse * nor[j];
            if (fabs(vt) > ROUND_EPS)
                c->fric[3 * p[i] + j] += dmax(-fabs(lambda * w[i] * t_impulse / vt), -1.0) * (v_rel[j] - vn * nor[j]);
        }
        c->cnt[p[i]] += 1;
    }
    if (!is_static(c, p[3])) {                                     /* :1067-1079 */
        double t_impulse = m_impulse;
        if (is_movable(c, p[3])) t_impulse = rigid_impulse[1];
        for (int j = 0; j < 3; ++j) {
            c->imp[3 * p[3] + j] -= t_impulse * nor[j];
            if (fabs(vt) > ROUND_EPS)
                c->fric[3 * p[3] + j] += dmax(-fabs(lambda * t_impulse / vt), -1.0) * (v_rel[j] - vn * nor[j]);
        }
        c->cnt[p[3]] += 1;
    }
    /* :1080-1082 zeroes collsnImpulse of static points; they never accumulate, so a no-op */
}

/* EdgeToEdgeImpulse, dcollid3d.cpp:1109-1300.  Edge p[0]-p[1] against edge p[2]-p[3]. */
static void edge_to_edge_impulse(orc_ctx* c, const int* p, const double* nor, double a, double b, double dist)
{
    const double* v[4];
    for (int i = 0; i < 4; ++i) v[i] = c->av + 3 * p[i];
    double v_rel[3], vn, vt;
    double impulse = 0.0, m_impulse;
    double rigid_impulse[2] = {0.0, 0.0};
    double wa[2] = {1.0 - a, a}, wb[2] = {1.0 - b, b};
    double k = c->k, m = c->m, dt = c->dt, lambda = c->lambda, h = c->thickness, cr = c->cr;
    dist = h - dist;
    for (int j = 0; j < 3; ++j) {                                  /* :1132-1136 */
        v_rel[j] = (1.0 - b) * v[2][j] + b * v[3][j];
        v_rel[j] -= (1.0 - a) * v[0][j] + a * v[1][j];
    }
    vn = dot3(v_rel, nor);
    if (dot3(v_rel, v_rel) > vn * vn) vt = sqrt(dot3(v_rel, v_rel) - vn * vn);
    else vt = 0.0;
    if (vn < 0.0) {                                                /* :1143-1176 */
        if ((is_static(c, p[0]) && is_static(c, p[1])) || (is_static(c, p[2]) && is_static(c, p[3]))) {
            impulse = vn; rigid_impulse[0] = vn; rigid_impulse[1] = vn;
        } else if (is_movable(c, p[0]) && is_movable(c, p[1]) && is_movable(c, p[2]) && is_movable(c, p[3])) {
            double m1 = c->hs_mass[c->vhs[p[0]]], m2 = c->hs_mass[c->vhs[p[2]]];
            rigid_impulse[0] = vn * m2 / (m1 + m2);
            rigid_impulse[1] = vn * m1 / (m1 + m2);
        } else if (is_movable(c, p[0]) && is_movable(c, p[1])) {
            rigid_impulse[0] = 0.5 * vn; impulse = 0.5 * vn;
        } else if (is_movable(c, p[2]) && is_movable(c, p[3])) {
            impulse = 0.5 * vn; rigid_impulse[1] = 0.5 * vn;
        } else
            impulse = vn * 0.5;
        if (is_static(c, p[0])) wa[0] = 0.0;
        if (is_static(c, p[1])) wa[1] = 0.0;
        if (is_static(c, p[2])) wb[0] = 0.0;
        if (is_static(c, p[3])) wb[1] = 0.0;
    }
    int all_rigid = is_rigid(c, p[0]) && is_rigid(c, p[1]) && is_rigid(c, p[2]) && is_rigid(c, p[3]);
    if (vn * dt < 0.1 * dist) {                                    /* :1177-1192 */
        if (all_rigid) {
            rigid_impulse[0] *= 1.0 + cr; rigid_impulse[1] *= 1.0 + cr;
        } else {
            double tmp = -dmin(dt * k * dist / m, (0.1 * dist / dt - vn));
            impulse += tmp; rigid_impulse[0] += tmp; rigid_impulse[1] += tmp;
        }
    }
    if (wa[0] + wa[1] < MACH_EPS || wb[0] + wb[1] < MACH_EPS) m_impulse = impulse; /* :1193-1197 */
    else m_impulse = 2.0 * impulse / (wa[0] * wa[0] + wa[1] * wa[1] + wb[0] * wb[0] + wb[1] * wb[1]);
    if (all_rigid) {                                               /* :1226-1234 */
        if (is_movable(c, p[0])) spread_rg(c, p[0], rigid_impulse[0], nor);
        if (is_movable(c, p[2])) spread_rg(c, p[2], -1.0 * rigid_impulse[1], nor);
        return;
    }
    const double wgt[4] = {wa[0], wa[1], wb[0], wb[1]};
    for (int j = 0; j < 3; ++j) {                                  /* :1235-1284 */
        for (int q = 0; q < 4; ++q) {
            if (is_static(c, p[q])) continue;
            double t_impulse = m_impulse;
            if (is_movable(c, p[q])) t_impulse = rigid_impulse[q < 2 ? 0 : 1];
            if (q < 2) c->imp[3 * p[q] + j] += wgt[q] * t_impulse * nor[j];
            else c->imp[3 * p[q] + j] -= wgt[q] * t_impulse * nor[j];
            if (fabs(vt) > ROUND_EPS)
                c->fric[3 * p[q] + j] += dmax(-fabs(lambda * wgt[q] * t_impulse / vt), -1.0) * (v_rel[j] - vn * nor[j]);
            if (j == 0) c->cnt[p[q]] += 1;
        }
    }
}

/* ------------------------------------------------------------------ static tests */
typedef struct { int ea, eb, feature; } tag_t;

/* PointToTri, dcollid3d.cpp:778-922.  X = Coords of the four points at test time. */
static int point_to_tri(orc_ctx* c, const int* p, double X[4][3], double h, double root, const tag_t* tag)
{
    double w[3] = {0.0, 0.0, 0.0};
    double x13[3], x23[3], x43[3], nor[3], nor_mag, dist, det;
    sub3(X[0], X[2], x13);                                         /* :793-795 */
    sub3(X[1], X[2], x23);
    sub3(X[3], X[2], x43);
    det = dot3(x13, x13) * dot3(x23, x23) - dot3(x13, x23) * dot3(x13, x23);
    if (fabs(det) < 1000 * MACH_EPS) return 0;                     /* :798-799 degenerate tri ignored */
    cross3(x13, x23, nor);                                         /* :849-868 */
    nor_mag = mag3(nor);
    double x43_old[3];
    sub3(c->xo + 3 * p[3], c->xo + 3 * p[2], x43_old);
    dist = dot3(x43_old, nor);
    for (int i = 0; i < 3; ++i) nor[i] /= nor_mag * ((dist >= 0) ? 1.0 : -1.0);
    dist = fabs(dot3(x43, nor));
    w[0] = (dot3(x13, x43) * dot3(x23, x23) - dot3(x23, x43) * dot3(x13, x23)) / det;
    w[1] = (dot3(x13, x13) * dot3(x23, x43) - dot3(x13, x23) * dot3(x13, x43)) / det;
    w[2] = 1 - w[0] - w[1];
    if (fabs(w[0]) < ROUND_EPS || fabs(w[1]) < ROUND_EPS || fabs(w[2]) < ROUND_EPS) { /* :872-891 */
        double vec[3];
        for (int j = 0; j < 3; ++j) vec[j] = c->xo[3 * p[3] + j];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) vec[j] -= w[i] * c->xo[3 * p[i] + j];
        if (mag3(vec) > ROUND_EPS)
            for (int j = 0; j < 3; ++j) nor[j] = vec[j];
    }
    nor_mag = mag3(nor);                                           /* :894-902 */
    if (nor_mag > ROUND_EPS)
        for (int i = 0; i < 3; ++i) nor[i] /= nor_mag;
    else {
        c->error = 1; /* reference: clean_up(ERROR) */
        return 0;
    }
    if (dist > h) return 0;                                        /* :911-919 */
    for (int i = 0; i < 3; ++i)
        if (w[i] > 1 + c->eps || w[i] < -c->eps) return 0;
    if (tag) {
        orc_contact k;
        memset(&k, 0, sizeof(k));
        k.ea = tag->ea; k.eb = tag->eb; k.feature = tag->feature; k.kind = 0;
        for (int i = 0; i < 4; ++i) k.p[i] = p[i];
        k.root = root; k.dist = dist;
        for (int i = 0; i < 3; ++i) { k.nor[i] = nor[i]; k.w[i] = w[i]; }
        push_contact(c, &k);
    }
    point_to_tri_impulse(c, p, nor, w, dist);
    return 1;
}

/* EdgeToEdge, dcollid3d.cpp:643-776 (the parallel branch returns at :665; the code after it is dead). */
static int edge_to_edge(orc_ctx* c, const int* p, double X[4][3], double h, double root, const tag_t* tag)
{
    double x21[3], x43[3], x31[3], tmp[3], v1[3], v2[3], nor[3], nor_mag, dist, a, b;
    sub3(X[1], X[0], x21);                                         /* :659-661 */
    sub3(X[3], X[2], x43);
    sub3(X[2], X[0], x31);
    cross3(x21, x43, tmp);
    if (mag3(tmp) < ROUND_EPS) return 0;                           /* :663-665 parallel edges ignored */
    a = (dot3(x43, x43) * dot3(x21, x31) - dot3(x21, x43) * dot3(x43, x31)) /
        (dot3(x21, x21) * dot3(x43, x43) - dot3(x21, x43) * dot3(x21, x43)); /* :719-722 */
    b = (dot3(x21, x43) * dot3(x21, x31) - dot3(x21, x21) * dot3(x43, x31)) /
        (dot3(x21, x21) * dot3(x43, x43) - dot3(x21, x43) * dot3(x21, x43));
    a = dmax(dmin(a, 1.0), 0.0);
    b = dmax(dmin(b, 1.0), 0.0);
    for (int i = 0; i < 3; ++i) { v1[i] = a * x21[i]; v2[i] = b * x43[i]; }
    for (int i = 0; i < 3; ++i) { v1[i] = X[0][i] + v1[i]; v2[i] = X[2][i] + v2[i]; }
    sub3(v2, v1, nor);
    nor_mag = mag3(nor);
    if (nor_mag < 1000 * MACH_EPS) {                               /* :731-744 intersecting edges */
        for (int j = 0; j < 3; ++j) {
            nor[j] = (1.0 - b) * c->xo[3 * p[2] + j] + b * c->xo[3 * p[3] + j];
            nor[j] -= (1.0 - a) * c->xo[3 * p[0] + j] + a * c->xo[3 * p[1] + j];
        }
    }
    dist = 0.0;                                                    /* distBetweenCoords, dcollid.cpp:950-957 */
    for (int i = 0; i < 3; ++i) dist += (v1[i] - v2[i]) * (v1[i] - v2[i]);
    dist = sqrt(dist);
    if (dist > h) return 0;                                        /* :747 */
    nor_mag = mag3(nor);
    if (nor_mag < MACH_EPS) {
        c->error = 1; /* reference: clean_up(ERROR) */
        return 0;
    }
    for (int i = 0; i < 3; ++i) nor[i] /= nor_mag;
    if (tag) {
        orc_contact k;
        memset(&k, 0, sizeof(k));
        k.ea = tag->ea; k.eb = tag->eb; k.feature = tag->feature; k.kind = 1;
        for (int i = 0; i < 4; ++i) k.p[i] = p[i];
        k.root = root; k.dist = dist;
        for (int i = 0; i < 3; ++i) k.nor[i] = nor[i];
        k.w[0] = a; k.w[1] = b; k.w[2] = 0.0;
        push_contact(c, &k);
    }
    edge_to_edge_impulse(c, p, nor, a, b, dist);
    return 1;
}

/* ------------------------------------------------------------------ CCD */
/* isCoplanar, dcollid3d.cpp:371-482.  roots[0..2] written; returns 1 iff some root > MACH_EPS. */
static int is_coplanar(const orc_ctx* c, const int* p, double dt, double* roots)
{
    double v[4][3], x[4][3];
    const double* v0 = c->av + 3 * p[0];
    const double* x0 = c->xo + 3 * p[0];
    for (int i = 1; i < 4; ++i)
        for (int j = 0; j < 3; ++j) {
            v[i][j] = c->av[3 * p[i] + j] - v0[j];
            x[i][j] = c->xo[3 * p[i] + j] - x0[j];
        }
    double a, b, cc, d, vv[3], vx[3], xx[3];
    vv[0] = v[1][1] * v[2][2] - v[1][2] * v[2][1];                 /* :402-412 */
    vv[1] = v[1][0] * v[2][2] - v[1][2] * v[2][0];
    vv[2] = v[1][0] * v[2][1] - v[1][1] * v[2][0];
    vx[0] = v[1][1] * x[2][2] - v[1][2] * x[2][1] - v[2][1] * x[1][2] + v[2][2] * x[1][1];
    vx[1] = v[1][0] * x[2][2] - v[1][2] * x[2][0] - v[2][0] * x[1][2] + v[2][2] * x[1][0];
    vx[2] = v[1][0] * x[2][1] - v[1][1] * x[2][0] - v[2][0] * x[1][1] + v[2][1] * x[1][0];
    xx[0] = x[1][1] * x[2][2] - x[1][2] * x[2][1];
    xx[1] = x[1][0] * x[2][2] - x[1][2] * x[2][0];
    xx[2] = x[1][0] * x[2][1] - x[1][1] * x[2][0];
    a = v[3][0] * vv[0] - v[3][1] * vv[1] + v[3][2] * vv[2];       /* :414-422 */
    b = x[3][0] * vv[0] - x[3][1] * vv[1] + x[3][2] * vv[2] + v[3][0] * vx[0] - v[3][1] * vx[1] + v[3][2] * vx[2];
    cc = x[3][0] * vx[0] - x[3][1] * vx[1] + x[3][2] * vx[2] + v[3][0] * xx[0] - v[3][1] * xx[1] + v[3][2] * xx[2];
    d = x[3][0] * xx[0] - x[3][1] * xx[1] + x[3][2] * xx[2];
    if (fabs(a) > MACH_EPS) {                                      /* :425-450 */
        b /= a; cc /= a; d /= a;
        a = b; b = cc; cc = d;
        double Q = (a * a - 3 * b) / 9;
        double R = (2 * a * a * a - 9 * a * b + 27 * cc) / 54;
        double Q3 = Q * Q * Q, R2 = R * R;
        if (R2 < Q3) {
            double Qsqrt = sqrt(Q);
            double theta = m_acos(R / sqrt(Q3));
            roots[0] = -2 * Qsqrt * m_cos(theta / 3) - a / 3;
            roots[1] = -2 * Qsqrt * m_cos((theta + 2 * M_PI) / 3) - a / 3;
            roots[2] = -2 * Qsqrt * m_cos((theta - 2 * M_PI) / 3) - a / 3;
        } else {
            double sgn = (R > 0) ? 1.0 : -1.0;
            double A = -sgn * m_pow13(fabs(R) + sqrt(R2 - Q3));
            double Bv = (fabs(A) < ROUND_EPS) ? 0.0 : Q / A;
            roots[0] = (A + Bv) - a / 3.0;
            if (fabs(A - Bv) < ROUND_EPS) roots[1] = roots[2] = -0.5 * (A + Bv) - a / 3.0;
        }
    } else {                                                       /* :451-463 */
        a = b; b = cc; cc = d;
        double delta = b * b - 4.0 * a * cc;
        if (fabs(a) > ROUND_EPS && delta > 0) {
            double ds = sqrt(delta);
            roots[0] = (-b + ds) / (2.0 * a);
            roots[1] = (-b - ds) / (2.0 * a);
        } else if (fabs(a) < ROUND_EPS && fabs(b) > ROUND_EPS) {
            roots[0] = -cc / b;
        }
    }
    for (int i = 0; i < 3; ++i) {                                  /* :465-469 */
        roots[i] = roots[i] - MACH_EPS;
        if (roots[i] < 0 || roots[i] > dt) roots[i] = -1;
    }
    double t;                                                      /* :471-476 */
    if (roots[0] > roots[1]) { t = roots[0]; roots[0] = roots[1]; roots[1] = t; }
    if (roots[0] > roots[2]) { t = roots[0]; roots[0] = roots[2]; roots[2] = t; }
    if (roots[1] > roots[2]) { t = roots[1]; roots[1] = roots[2]; roots[2] = t; }
    return roots[0] > MACH_EPS || roots[1] > MACH_EPS || roots[2] > MACH_EPS;
}

/* MovingPointToTri / MovingEdgeToEdge, dcollid3d.cpp:327-369 */
static int moving_test(orc_ctx* c, int edge, const int* p, double h, const tag_t* tag, double* hit_root)
{
    double dt = c->dt;
    double roots[4] = {-1, -1, -1, dt};
    if (hit_root) *hit_root = -1.0;
    if (!is_coplanar(c, p, dt, roots)) return 0;
    for (int i = 0; i < 4; ++i) {
        if (roots[i] < 0) continue;
        double X[4][3];
        for (int j = 0; j < 4; ++j)
            for (int k = 0; k < 3; ++k) X[j][k] = c->xo[3 * p[j] + k] + roots[i] * c->av[3 * p[j] + k];
        int hit = edge ? edge_to_edge(c, p, X, h, roots[i], tag) : point_to_tri(c, p, X, h, roots[i], tag);
        if (hit) {
            if (hit_root) *hit_root = roots[i];
            return 1;
        }
    }
    return 0;
}

static int static_test(orc_ctx* c, int edge, const int* p, double h, const tag_t* tag)
{
    double X[4][3];
    for (int j = 0; j < 4; ++j)
        for (int k = 0; k < 3; ++k) X[j][k] = c->xo[3 * p[j] + k]; /* "make sure the coords are old coords" */
    return edge ? edge_to_edge(c, p, X, h, 0.0, tag) : point_to_tri(c, p, X, h, 0.0, tag);
}

static int feature_test(orc_ctx* c, int moving, int edge, const int* p, double h, tag_t* tag)
{
    int r = moving ? moving_test(c, edge, p, h, tag, NULL) : static_test(c, edge, p, h, tag);
    tag->feature++;
    return r;
}

/* ------------------------------------------------------------------ element pairs */
static int elem_rigid(const orc_ctx* c, const int* pts, int n) /* isRigidBody(CD_HSE*), dcollid.cpp:1097 */
{
    for (int i = 0; i < n; ++i)
        if (is_rigid(c, pts[i])) return 1;
    return 0;
}

static void uf_merge(orc_ctx* c, int X, int Y);
static void uf_build(orc_ctx* c);
/* createImpZone(pts, 4, first = NO), dcollid.cpp:473-484: movable-rigid points never join a zone */
static void create_imp_zone(orc_ctx* c, const int* p)
{
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < i; ++j) {
            if (is_movable(c, p[i]) || is_movable(c, p[j])) continue;
            uf_merge(c, p[i], p[j]);
        }
}
/* `if (status && is_detImpZone) createImpZone(pts,4)` after every Moving* feature test
 * (dcollid3d.cpp:222-240, :268, :301, :320): status is sticky within the element pair */
#define ZONE_HOOK() do { if (status && moving && c->det_imp_zone) create_imp_zone(c, p); } while (0)

/* TriToTri (dcollid3d.cpp:570-627) / MovingTriToTri (:274-325) */
static int tri_tri(orc_ctx* c, int moving, const int* A, const int* Bt, double h, tag_t* tag)
{
    int status = 0, p[4];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            if (A[i] == Bt[j]) return 0;
    for (int k = 0; k < 2; ++k)
        for (int i = 0; i < 3; ++i) {
            /* proximity: k=0 tests vertices of tri1 against tri2 (:595-599);
             * CCD:       k=0 tests vertices of b against tri a   (:288-292) */
            const int* tri = moving ? (k == 0 ? A : Bt) : (k == 0 ? Bt : A);
            const int* other = moving ? (k == 0 ? Bt : A) : (k == 0 ? A : Bt);
            p[0] = tri[0]; p[1] = tri[1]; p[2] = tri[2]; p[3] = other[i];
            if (feature_test(c, moving, 0, p, h, tag)) status = 1;
            ZONE_HOOK();
        }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            p[0] = A[i]; p[1] = A[(i + 1) % 3]; p[2] = Bt[j]; p[3] = Bt[(j + 1) % 3];
            if (feature_test(c, moving, 1, p, h, tag)) status = 1;
            ZONE_HOOK();
        }
    return status;
}

/* TriToBond (dcollid3d.cpp:485-533) / MovingTriToBond (:203-244) */
static int tri_bond(orc_ctx* c, int moving, const int* tri, const int* bd, double h, tag_t* tag)
{
    int status = 0, p[4];
    for (int i = 0; i < 3; ++i)
        if (tri[i] == bd[0] || tri[i] == bd[1]) return 0;
    p[0] = tri[0]; p[1] = tri[1]; p[2] = tri[2];
    p[3] = bd[0];
    if (feature_test(c, moving, 0, p, h, tag)) status = 1;
    ZONE_HOOK();
    p[3] = bd[1];
    if (feature_test(c, moving, 0, p, h, tag)) status = 1;
    ZONE_HOOK();
    p[2] = bd[0]; p[3] = bd[1];
    for (int i = 0; i < 3; ++i) {
        p[0] = tri[i]; p[1] = tri[(i + 1) % 3];
        if (feature_test(c, moving, 1, p, h, tag)) status = 1;
        ZONE_HOOK();
    }
    return status;
}

/* BondToBond (dcollid3d.cpp:535-568) / MovingBondToBond (:246-272) */
static int bond_bond(orc_ctx* c, int moving, const int* b1, const int* b2, double h, tag_t* tag)
{
    int p[4] = {b1[0], b1[1], b2[0], b2[1]};
    for (int i = 0; i < 4; ++i)
        for (int j = i + 1; j < 4; ++j)
            if (p[i] == p[j]) return 0;
    int status = feature_test(c, moving, 1, p, h, tag);
    ZONE_HOOK();
    return status;
}

/* isProximity / isCollision, dcollid.cpp:753-836: dispatch + same-surface rigid filter */
static int element_pair(orc_ctx* c, int moving, int ea, int eb)
{
    double h = moving ? c->eps : c->thickness;
    tag_t tag = {ea, eb, 0};
    int a_tri = ea < c->T, b_tri = eb < c->T;
    if (a_tri && b_tri) {
        const int* A = c->tri + 3 * ea;
        const int* Bt = c->tri + 3 * eb;
        if (c->tri_surf[ea] == c->tri_surf[eb] && elem_rigid(c, A, 3)) return 0;
        return tri_tri(c, moving, A, Bt, h, &tag);
    }
    if (!a_tri && !b_tri) return bond_bond(c, moving, c->bond + 2 * (ea - c->T), c->bond + 2 * (eb - c->T), h, &tag);
    if (a_tri) return tri_bond(c, moving, c->tri + 3 * ea, c->bond + 2 * (eb - c->T), h, &tag);
    return tri_bond(c, moving, c->tri + 3 * eb, c->bond + 2 * (ea - c->T), h, &tag);
}

/* ------------------------------------------------------------------ broad phase */
/* Leaf boxes: CD_TRI/CD_BOND::{min,max}_{static,moving}_coord (dcollid.cpp:852-932) -+ 1e-6
 * (AABB.cpp:6-36).  Candidates: closed-interval overlap on all axes (AABB.cpp:56-60). */
static void leaf_box(const orc_ctx* c, int e, int moving, double* lo, double* hi)
{
    const int* pts = e < c->T ? c->tri + 3 * e : c->bond + 2 * (e - c->T);
    int n = e < c->T ? 3 : 2;
    for (int d = 0; d < 3; ++d) {
        double mn = 1.0e18, mx = -1.0e18;
        for (int i = 0; i < n; ++i) {
            double x0 = c->xo[3 * pts[i] + d];
            mn = dmin(mn, x0); mx = dmax(mx, x0);
            if (moving) {
                double x1 = c->xo[3 * pts[i] + d] + c->av[3 * pts[i] + d] * c->dt;
                mn = dmin(mn, x1); mx = dmax(mx, x1);
            }
        }
        lo[d] = mn - BOX_PAD;
        hi[d] = mx + BOX_PAD;
    }
}

static int cmp_pair(const void* a, const void* b)
{
    const int* p = (const int*)a;
    const int* q = (const int*)b;
    if (p[0] != q[0]) return p[0] < q[0] ? -1 : 1;
    if (p[1] != q[1]) return p[1] < q[1] ? -1 : 1;
    return 0;
}

static void push_cand(orc_ctx* c, int a, int b)
{
    if (c->n_cand == c->cap_cand) {
        c->cap_cand = c->cap_cand ? 2 * c->cap_cand : 4096;
        c->cand = (int*)realloc(c->cand, (size_t)c->cap_cand * 2 * sizeof(int));
    }
    c->cand[2 * c->n_cand] = a < b ? a : b;
    c->cand[2 * c->n_cand + 1] = a < b ? b : a;
    c->n_cand++;
}

/* Uniform-grid all-pairs box overlap; a pair is reported from the one cell that holds the
 * componentwise max of the two lower corners, so each unordered pair appears once. */
static void broad_phase(orc_ctx* c, int moving)
{
    int N = c->N;
    c->n_cand = 0;
    if (N < 2) return;
    double* lo = (double*)malloc((size_t)N * 3 * sizeof(double));
    double* hi = (double*)malloc((size_t)N * 3 * sizeof(double));
    double glo[3] = {1e300, 1e300, 1e300}, ghi[3] = {-1e300, -1e300, -1e300}, ext[3] = {0, 0, 0};
    for (int e = 0; e < N; ++e) {
        leaf_box(c, e, moving, lo + 3 * e, hi + 3 * e);
        for (int d = 0; d < 3; ++d) {
            if (lo[3 * e + d] < glo[d]) glo[d] = lo[3 * e + d];
            if (hi[3 * e + d] > ghi[d]) ghi[d] = hi[3 * e + d];
            ext[d] += hi[3 * e + d] - lo[3 * e + d];
        }
    }
    int G[3];
    double cell[3];
    for (int d = 0; d < 3; ++d) {
        double s = 2.0 * ext[d] / N;
        if (!(s > 1e-9)) s = 1e-9;
        double span = ghi[d] - glo[d];
        double g = floor(span / s) + 1.0;
        if (g > 1024.0) g = 1024.0;
        if (g < 1.0) g = 1.0;
        G[d] = (int)g;
        cell[d] = span / G[d];
        if (!(cell[d] > 0)) cell[d] = 1.0;
    }
    long ncell = (long)G[0] * G[1] * G[2];
    int* cnt = (int*)calloc((size_t)ncell + 1, sizeof(int));
#define CELL_OF(v, d) ((int)dmax(0.0, dmin((double)(G[d] - 1), floor(((v)-glo[d]) / cell[d]))))
    for (int pass = 0; pass < 2; ++pass) {
        int* items = NULL;
        if (pass == 1) {
            long tot = 0;
            for (long q = 0; q < ncell; ++q) { int t = cnt[q]; cnt[q] = (int)tot; tot += t; }
            cnt[ncell] = (int)tot;
            items = (int*)malloc((size_t)(tot ? tot : 1) * sizeof(int));
            int* fill = (int*)malloc((size_t)ncell * sizeof(int));
            memcpy(fill, cnt, (size_t)ncell * sizeof(int));
            for (int e = 0; e < N; ++e) {
                int a[3], b[3];
                for (int d = 0; d < 3; ++d) { a[d] = CELL_OF(lo[3 * e + d], d); b[d] = CELL_OF(hi[3 * e + d], d); }
                for (int i = a[0]; i <= b[0]; ++i)
                    for (int j = a[1]; j <= b[1]; ++j)
                        for (int k = a[2]; k <= b[2]; ++k) items[fill[((long)i * G[1] + j) * G[2] + k]++] = e;
            }
            free(fill);
            for (long q = 0; q < ncell; ++q) {
                int qi = (int)(q / ((long)G[1] * G[2])), qj = (int)((q / G[2]) % G[1]), qk = (int)(q % G[2]);
                for (int s = cnt[q]; s < cnt[q + 1]; ++s)
                    for (int t = s + 1; t < cnt[q + 1]; ++t) {
                        int e = items[s], f = items[t];
                        const double *l1 = lo + 3 * e, *h1 = hi + 3 * e, *l2 = lo + 3 * f, *h2 = hi + 3 * f;
                        if (!(l1[0] <= h2[0] && h1[0] >= l2[0] && l1[1] <= h2[1] && h1[1] >= l2[1] &&
                              l1[2] <= h2[2] && h1[2] >= l2[2]))
                            continue;
                        if (CELL_OF(dmax(l1[0], l2[0]), 0) != qi || CELL_OF(dmax(l1[1], l2[1]), 1) != qj ||
                            CELL_OF(dmax(l1[2], l2[2]), 2) != qk)
                            continue;
                        push_cand(c, e, f);
                    }
            }
            free(items);
        } else {
            for (int e = 0; e < N; ++e) {
                int a[3], b[3];
                for (int d = 0; d < 3; ++d) { a[d] = CELL_OF(lo[3 * e + d], d); b[d] = CELL_OF(hi[3 * e + d], d); }
                for (int i = a[0]; i <= b[0]; ++i)
                    for (int j = a[1]; j <= b[1]; ++j)
                        for (int k = a[2]; k <= b[2]; ++k) cnt[((long)i * G[1] + j) * G[2] + k]++;
            }
        }
    }
#undef CELL_OF
    free(cnt); free(lo); free(hi);
    qsort(c->cand, (size_t)c->n_cand, 2 * sizeof(int), cmp_pair);
}

/* aabbProximity + query / aabbCollision + query (dcollid.cpp:366-428, AABB.cpp:254-343) */
long orc_detect(orc_ctx* c, int mode)
{
    int moving = mode == ORC_COLLISION;
    if (c->det_imp_zone && !c->uf_ready) uf_build(c);
    broad_phase(c, moving);
    c->n_con = 0;
    c->n_true = 0;
    for (long q = 0; q < c->n_cand; ++q) {
        int ea = c->cand[2 * q], eb = c->cand[2 * q + 1];
        if (element_pair(c, moving, ea, eb)) {
            if (c->n_true == c->cap_true) {
                c->cap_true = c->cap_true ? 2 * c->cap_true : 1024;
                c->truep = (int*)realloc(c->truep, (size_t)c->cap_true * 2 * sizeof(int));
            }
            c->truep[2 * c->n_true] = ea;
            c->truep[2 * c->n_true + 1] = eb;
            c->n_true++;
        }
    }
    return c->n_true;
}

/* Narrow phase over an explicit list of ORDERED pairs (a, b), in the given order: used to replay
 * the exact callback sequence of the reference's tree query so that accumulators can be compared
 * with the compiled reference bit for bit (tests/golden).  The candidate list is left untouched. */
long orc_detect_ordered(orc_ctx* c, int mode, const int* pairs, long n)
{
    int moving = mode == ORC_COLLISION;
    if (c->det_imp_zone && !c->uf_ready) uf_build(c);
    c->n_con = 0;
    c->n_true = 0;
    for (long q = 0; q < n; ++q) {
        int ea = pairs[2 * q], eb = pairs[2 * q + 1];
        if (element_pair(c, moving, ea, eb)) {
            if (c->n_true == c->cap_true) {
                c->cap_true = c->cap_true ? 2 * c->cap_true : 1024;
                c->truep = (int*)realloc(c->truep, (size_t)c->cap_true * 2 * sizeof(int));
            }
            c->truep[2 * c->n_true] = ea;
            c->truep[2 * c->n_true + 1] = eb;
            c->n_true++;
        }
    }
    return c->n_true;
}

/* ------------------------------------------------------------------ per-point passes */
/* computeAverageVelocity, dcollid.cpp:160-220 */
void orc_avg_velocity(orc_ctx* c)
{
    for (int i = 0; i < 3 * c->V; ++i) {
        if (c->dt > ROUND_EPS) c->av[i] = (c->x[i] - c->xo[i]) / c->dt;
        else c->av[i] = 0.0;
        if (isnan(c->av[i]) || isinf(c->av[i])) c->error = 1;
    }
    memcpy(c->x, c->xo, (size_t)3 * c->V * sizeof(double));
}

/* ---- union-find restated (dcollid.cpp:995-1059), used only for movable rigid bodies */
static int uf_find(orc_ctx* c, int p)
{
    if (c->uf_root[p] != p) c->uf_root[p] = uf_find(c, c->uf_root[p]);
    return c->uf_root[p];
}
static void uf_merge(orc_ctx* c, int X, int Y)
{
    int PX = uf_find(c, X), PY = uf_find(c, Y);
    if (PX == PY) return;
    if (c->uf_weight[PX] > c->uf_weight[PY]) {
        c->uf_weight[PX] += c->uf_weight[PY];
        c->uf_root[PY] = PX;
        c->uf_next[c->uf_tail[PX]] = PY;
        c->uf_tail[PX] = c->uf_tail[PY];
    } else {
        c->uf_weight[PY] += c->uf_weight[PX];
        c->uf_root[PX] = PY;
        c->uf_next[c->uf_tail[PY]] = PX;
        c->uf_tail[PY] = c->uf_tail[PX];
    }
}
/* makeSet + createImpZoneForRG, dcollid.cpp:1015-1031, dcollid3d.cpp:54-68 */
static void uf_build(orc_ctx* c)
{
    for (int v = 0; v < c->V; ++v) {
        c->uf_root[v] = v; c->uf_next[v] = -1; c->uf_tail[v] = v; c->uf_weight[v] = 1;
    }
    int t = 0;
    while (t < c->T) {
        int s = c->tri_surf[t], t0 = t;
        while (t < c->T && c->tri_surf[t] == s) ++t;
        if (!is_movable(c, c->tri[3 * t0])) continue; /* first_tri's point 0 decides */
        for (int q = t0; q < t; ++q)
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < i; ++j) uf_merge(c, c->tri[3 * q + i], c->tri[3 * q + j]);
    }
    c->uf_ready = 1;
}

static double det3(double a[3][3]) /* myDet3d, dcollid.cpp:977-981 */
{
    return a[0][0] * (a[1][1] * a[2][2] - a[2][1] * a[1][2]) - a[0][1] * (a[1][0] * a[2][2] - a[2][0] * a[1][2]) +
           a[0][2] * (a[1][0] * a[2][1] - a[2][0] * a[1][1]);
}

/* updateImpactListVelocity, dcollid3d.cpp:70-200: make one union-find list move rigidly */
static void rigidify_list(orc_ctx* c, int head)
{
    double m = c->m, dt = c->dt;
    double x_cm[3] = {0, 0, 0}, v_cm[3] = {0, 0, 0}, L[3] = {0, 0, 0}, I[3][3] = {{0}}, tmp[3][3];
    int num = 0;
    for (int p = head; p >= 0; p = c->uf_next[p]) {
        num++;
        for (int i = 0; i < 3; ++i) { x_cm[i] += c->xo[3 * p + i]; v_cm[i] += c->av[3 * p + i]; }
    }
    for (int i = 0; i < 3; ++i) { x_cm[i] /= num; v_cm[i] /= num; }
    for (int p = head; p >= 0; p = c->uf_next[p]) {
        double dx[3], dv[3], Li[3];
        sub3(c->xo + 3 * p, x_cm, dx);
        sub3(c->av + 3 * p, v_cm, dv);
        cross3(dx, dv, Li);
        for (int i = 0; i < 3; ++i) Li[i] = m * Li[i];
        for (int i = 0; i < 3; ++i) L[i] = Li[i] + L[i];
    }
    for (int p = head; p >= 0; p = c->uf_next[p]) {
        double dx[3];
        sub3(c->xo + 3 * p, x_cm, dx);
        double mag_dx = mag3(dx);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                tmp[i][j] = -dx[i] * dx[j];
                if (i == j) tmp[i][j] += mag_dx * mag_dx;
                I[i][j] += tmp[i][j] * m;
            }
    }
    double w[3], mag_w;
    for (int i = 0; i < 3; ++i) {
        memcpy(tmp, I, sizeof(tmp));
        for (int j = 0; j < 3; ++j) tmp[j][i] = L[j];
        if (det3(I) < ROUND_EPS) w[i] = 0.0;
        else w[i] = det3(tmp) / det3(I);
    }
    mag_w = mag3(w);
    for (int p = head; p >= 0; p = c->uf_next[p]) {
        if (is_static(c, p)) continue;
        double dx[3], xF[3], xR[3], wxR[3], tmpV[3];
        sub3(c->xo + 3 * p, x_cm, dx);
        if (mag_w < ROUND_EPS) {
            for (int i = 0; i < 3; ++i) { xF[i] = dx[i]; wxR[i] = 0.0; }
            sub3(dx, xF, xR);
        } else {
            double s = dot3(dx, w) / dot3(w, w);
            for (int i = 0; i < 3; ++i) xF[i] = s * w[i];
            sub3(dx, xF, xR);
            double q = m_sin(dt * mag_w) / mag_w;
            for (int i = 0; i < 3; ++i) tmpV[i] = q * w[i];
            cross3(tmpV, xR, wxR);
        }
        for (int i = 0; i < 3; ++i) {
            double x_new = x_cm[i] + dt * v_cm[i] + xF[i] + m_cos(dt * mag_w) * xR[i] + wxR[i];
            c->av[3 * p + i] = (x_new - c->xo[3 * p + i]) / dt;
            if (isnan(c->av[3 * p + i])) c->error = 1;
        }
    }
}

/* updateAverageVelocity, dcollid.cpp:677-751 (+ updateImpactZoneVelocityForRG :267-288) */
void orc_apply(orc_ctx* c, int rigidify)
{
    for (int p = 0; p < c->V; ++p) {
        if (is_static(c, p)) continue;
        if (c->cnt[p] > 0) {
            c->has[p] = 1;
            for (int k = 0; k < 3; ++k) {
                c->av[3 * p + k] += (c->imp[3 * p + k] + c->fric[3 * p + k]) / c->cnt[p];
                if (isinf(c->av[3 * p + k]) || isnan(c->av[3 * p + k])) c->error = 1;
                c->imp[3 * p + k] = c->fric[3 * p + k] = 0.0;
            }
            c->cnt[p] = 0;
        }
        int b = c->vhs[p];
        if (c->cnt_rg[b] > 0) {
            c->has[p] = 1;
            for (int k = 0; k < 3; ++k) c->av[3 * p + k] += c->imp_rg[3 * b + k] / c->cnt_rg[b];
        }
    }
    /* the reference zeroes collsn_num_RG point by point (:732); collsnImpulse_RG is never zeroed */
    for (int b = 0; b < c->nhs; ++b) c->cnt_rg[b] = 0;
    if (rigidify && c->dt > 0.0) {
        if (!c->uf_ready) uf_build(c);
        for (int p = 0; p < c->V; ++p) {
            if (!is_movable(c, p)) continue;
            int r = uf_find(c, p);
            if (r == p && c->uf_weight[r] > 1) rigidify_list(c, r);
        }
    }
}

/* updateImpactZoneVelocity, dcollid.cpp:290-309: every union-find set with more than one point
 * (impact zones and movable bodies alike) is made to move rigidly; returns the number of sets */
int orc_zone_velocity(orc_ctx* c)
{
    int zones = 0;
    if (!c->uf_ready) uf_build(c);
    memset(c->sorted, 0, (size_t)c->V);
    for (int e = 0; e < c->N; ++e) {
        const int* pts = e < c->T ? c->tri + 3 * e : c->bond + 2 * (e - c->T);
        int n = e < c->T ? 3 : 2;
        for (int i = 0; i < n; ++i) {
            int p = pts[i], r = uf_find(c, p);
            if (c->sorted[p] || c->uf_weight[r] == 1) continue;
            for (int q = r; q >= 0; q = c->uf_next[q]) c->sorted[q] = 1;
            if (c->dt > 0.0) rigidify_list(c, r);
            zones++;
        }
    }
    return zones;
}

void orc_set_imp_zone(orc_ctx* c, int on) { c->det_imp_zone = on; } /* turnOn/OffImpZone, dcollid.cpp:222-223 */
void orc_enable_impact_zones(orc_ctx* c, int on) { c->impact_zones = on; }

/* computeImpactZone, dcollid.cpp:227-265.  out[0] iterations, out[1] zones in the last iteration,
 * out[2] true pairs summed over the iterations.  The reference loops without bound; `max_iter`
 * (<= 0: 100000) only guards the test harness, hitting it sets the error flag. */
void orc_impact_zone(orc_ctx* c, int max_iter, long* out)
{
    int is_collision = 1, zones = 0;
    long it = 0, pairs = 0;
    if (max_iter <= 0) max_iter = 100000;
    c->det_imp_zone = 1;
    while (is_collision) {
        long n = orc_detect(c, ORC_COLLISION);
        is_collision = n > 0;
        pairs += n;
        orc_apply(c, 1);
        zones = orc_zone_velocity(c);
        if (++it >= max_iter && is_collision) { c->error = 1; break; }
    }
    c->det_imp_zone = 0;
    if (out) { out[0] = it; out[1] = zones; out[2] = pairs; }
}

/* detectDomainBoundaryCollision, dcollid.cpp:116-158 -- once per unique point */
void orc_boundary(orc_ctx* c)
{
    double dt = c->dt, mu = c->lambda;
    for (int p = 0; p < c->V; ++p) {
        if (is_movable(c, p)) continue;
        double dv = 0;
        for (int j = 0; j < 3; ++j) {
            double cand = c->xo[3 * p + j] + dt * c->av[3 * p + j];
            if (cand <= c->lo[j]) {
                c->has[p] = 1;
                dv = fabs(c->av[3 * p + j]);
                c->av[3 * p + j] = 0.0;
            } else if (cand >= c->hi[j]) {
                c->has[p] = 1;
                dv = fabs(c->av[3 * p + j]);
                c->av[3 * p + j] = 0.0;
            }
        }
        double preVt = mag3(c->av + 3 * p);
        if (preVt > MACH_EPS)
            for (int j = 0; j < 3; ++j) c->av[3 * p + j] *= dmax(1.0 - mu * dv / preVt, 0.0);
    }
}

void orc_final_position(orc_ctx* c) /* dcollid.cpp:562-584 */
{
    for (int i = 0; i < 3 * c->V; ++i) c->x[i] = c->xo[i] + c->av[i] * c->dt;
}

/* reduceSuperelastOnce, dcollid.cpp:485-560: one Gauss-Seidel sweep over the edges of every
 * non-rigid element in hseList order (a shared edge is visited once per adjacent triangle);
 * an edge whose strain or strain rate exceeds 10 % gets both end points the mean avgVel. */
static double dist3(const double* p, const double* q) /* distance_between_positions */
{
    double s = 0.0;
    for (int i = 0; i < 3; ++i) s += (p[i] - q[i]) * (p[i] - q[i]);
    return sqrt(s);
}
int orc_strain_limit_once(orc_ctx* c, long* num_edges)
{
    const double tol = 0.10, dt = c->dt;
    int has = 0;
    *num_edges = 0;
    for (int e = 0; e < c->N; ++e) {
        const int* pts = e < c->T ? c->tri + 3 * e : c->bond + 2 * (e - c->T);
        int np = e < c->T ? 3 : 2;
        if (elem_rigid(c, pts, np)) continue;
        for (int j = 0; j < (np == 2 ? 1 : np); ++j) {
            int p0 = pts[j % np], p1 = pts[(j + 1) % np];
            double* a0 = c->av + 3 * p0;
            double* a1 = c->av + 3 * p1;
            const double *x0 = c->xo + 3 * p0, *x1 = c->xo + 3 * p1;
            double c0[3], c1[3];
            for (int k = 0; k < 3; ++k) { c0[k] = x0[k] + dt * a0[k]; c1[k] = x1[k] + dt * a1[k]; }
            double len_new = dist3(c0, c1), len_old = dist3(x0, x1);
            double len0 = e < c->T ? c->tri_len0[3 * e + j] : c->bond_len0[e - c->T];
            int fix;
            if (len_old > ROUND_EPS && len_new > ROUND_EPS) {
                double strain_rate = (len_new - len_old) / len_old;
                double strain = (len_new - len0) / len0;
                fix = fabs(strain) > tol || fabs(strain_rate) > tol;
            } else
                fix = 1;
            if (fix) {
                for (int k = 0; k < 3; ++k) {
                    double v = 0.5 * (a0[k] + a1[k]);
                    a0[k] = v; a1[k] = v;
                }
                ++*num_edges;
                has = 1;
            }
        }
    }
    return has;
}

/* reduceSuperelast, dcollid.cpp:586-596: at most 10 sweeps; returns the number of sweeps run */
int orc_strain_limit(orc_ctx* c, long* num_edges)
{
    int has = 1, niter = 0;
    long n = 0;
    if (!c->have_len0) { c->error = 1; return -1; }
    while (has && niter++ < 10) has = orc_strain_limit_once(c, &n);
    if (niter > 10) niter = 10; /* the reference's post-increment leaves 11 after ten sweeps */
    if (num_edges) *num_edges = n;
    return niter;
}

void orc_set_rest_lengths(orc_ctx* c, const double* tri_len0, const double* bond_len0)
{
    memcpy(c->tri_len0, tri_len0, (size_t)3 * c->T * sizeof(double));
    memcpy(c->bond_len0, bond_len0, (size_t)c->B * sizeof(double));
    c->have_len0 = 1;
}
void orc_enable_strain_limiting(orc_ctx* c, int on) { c->strain_limiting = on; }

void orc_final_velocity(orc_ctx* c, double* vel) /* dcollid.cpp:598-624 */
{
    for (int p = 0; p < c->V; ++p)
        if (c->has[p])
            for (int j = 0; j < 3; ++j) vel[3 * p + j] = c->av[3 * p + j];
}

/* updateFinalForRG, dcollid.cpp:626-675 (called at the end of updateFinalVelocity): walk hseList, points in element
 * order; the first point of a movable body that has has_collsn sets that body's centre-of-mass velocity to its avgVel and
 * the centre of mass to avgVel * dt + mrg_com[body]; mrg_com[body] is refreshed from the (possibly just updated) centre of
 * mass the first time the body is met and again right after such an update.  com / com_velo: 3*nhs, caller-owned
 * (HYPER_SURF::center_of_mass / center_of_mass_velo), in/out.  A body met for the very first time whose first point already
 * collides reads mrg_com before the reference ever wrote it (an empty std::vector there): seeded with the incoming centre
 * of mass, as oracle/ref_wrapper.cpp seeds the reference. */
void orc_update_final_for_rg(orc_ctx* c, double* com, double* com_velo)
{
    unsigned char* in_mrg = (unsigned char*)calloc((size_t)c->nhs + 1, 1);
    signed char* visited = (signed char*)malloc((size_t)c->nhs + 1);   /* -1 absent, 0 false, 1 true */
    memset(visited, -1, (size_t)c->nhs + 1);
    for (int e = 0; e < c->N; ++e) {
        const int np = e < c->T ? 3 : 2;
        const int* pts = e < c->T ? c->tri + 3 * e : c->bond + 2 * (e - c->T);
        for (int i = 0; i < np; ++i) {
            const int p = pts[i];
            if (!(c->flags[p] & 2)) continue;
            const int rg = c->vhs[p];
            if (!c->mrg_valid[rg]) {
                for (int j = 0; j < 3; ++j) c->mrg_com[3 * rg + j] = com[3 * rg + j];
                c->mrg_valid[rg] = 1;
            }
            if (c->has[p] && !in_mrg[rg]) {
                in_mrg[rg] = 1;
                for (int j = 0; j < 3; ++j) {
                    com_velo[3 * rg + j] = c->av[3 * p + j];
                    com[3 * rg + j] = c->av[3 * p + j] * c->dt + c->mrg_com[3 * rg + j];
                }
                visited[rg] = 0;
            }
            if (visited[rg] <= 0) {
                for (int j = 0; j < 3; ++j) c->mrg_com[3 * rg + j] = com[3 * rg + j];
                visited[rg] = 1;
            }
        }
    }
    free(in_mrg);
    free(visited);
}

void orc_set_has_collsn(orc_ctx* c, const unsigned char* has) { memcpy(c->has, has, (size_t)c->V); }

/* resolveCollision, dcollid.cpp:317-362, with detectProximity :390-406 and detectCollision :430-468 */
void orc_resolve(orc_ctx* c, double* vel, long* stats)
{
    for (int i = 0; i < 20; ++i) stats[i] = 0;
    orc_avg_velocity(c);
    stats[0] = orc_detect(c, ORC_PROXIMITY);
    stats[8] = c->n_cand;
    orc_apply(c, 1);
    int is_collision = 1, niter = 1, cd = 0;
    while (is_collision) {
        long n = orc_detect(c, ORC_COLLISION);
        is_collision = n > 0;
        stats[2 + cd] = n;
        stats[9 + cd] = c->n_cand;
        cd++;
        orc_apply(c, 1);
        if (++niter > MAX_ITER) break;
    }
    stats[1] = cd;
    stats[7] = is_collision; /* the reference now enters computeImpactZone (dcollid.cpp:466) */
    if (is_collision && c->impact_zones) {
        long z[3];
        orc_impact_zone(c, 0, z);
        stats[14] = z[0]; stats[15] = z[1];
    }
    orc_boundary(c);
    orc_final_position(c);
    if (c->strain_limiting) { /* reduceSuperelast sits between the final positions and velocities (:355) */
        long ne = 0;
        stats[16] = orc_strain_limit(c, &ne);
        stats[17] = ne;
    }
    orc_final_velocity(c, vel);
}

/* ------------------------------------------------------------------ plumbing */
orc_ctx* orc_create(int V, int T, const int* tri_idx, const int* tri_surf, int B, const int* bond_idx,
                    const unsigned char* vflags, const int* vhs, int nhs, const double* hs_mass)
{
    orc_ctx* c = (orc_ctx*)calloc(1, sizeof(orc_ctx));
    c->V = V; c->T = T; c->B = B; c->N = T + B; c->nhs = nhs;
    c->tri = (int*)malloc((size_t)(3 * T + 1) * sizeof(int));
    c->tri_surf = (int*)malloc((size_t)(T + 1) * sizeof(int));
    c->bond = (int*)malloc((size_t)(2 * B + 1) * sizeof(int));
    c->flags = (unsigned char*)malloc((size_t)V + 1);
    c->vhs = (int*)malloc((size_t)(V + 1) * sizeof(int));
    c->hs_mass = (double*)malloc((size_t)(nhs + 1) * sizeof(double));
    memcpy(c->tri, tri_idx, (size_t)3 * T * sizeof(int));
    memcpy(c->tri_surf, tri_surf, (size_t)T * sizeof(int));
    memcpy(c->bond, bond_idx, (size_t)2 * B * sizeof(int));
    memcpy(c->flags, vflags, (size_t)V);
    memcpy(c->vhs, vhs, (size_t)V * sizeof(int));
    memcpy(c->hs_mass, hs_mass, (size_t)nhs * sizeof(double));
    c->eps = 1e-6; c->thickness = 1e-4; c->k = 1000; c->m = 0.01; c->lambda = 0.02; c->cr = 0.0; c->dt = 1e-3;
    for (int i = 0; i < 3; ++i) { c->lo[i] = -1e30; c->hi[i] = 1e30; }
    c->mrg_com = (double*)calloc((size_t)3 * nhs + 1, sizeof(double));
    c->mrg_valid = (unsigned char*)calloc((size_t)nhs + 1, 1);
    size_t n3 = (size_t)3 * V + 1;
    c->xo = (double*)calloc(n3, sizeof(double));
    c->x = (double*)calloc(n3, sizeof(double));
    c->av = (double*)calloc(n3, sizeof(double));
    c->imp = (double*)calloc(n3, sizeof(double));
    c->fric = (double*)calloc(n3, sizeof(double));
    c->cnt = (int*)calloc((size_t)V + 1, sizeof(int));
    c->has = (unsigned char*)calloc((size_t)V + 1, 1);
    c->imp_rg = (double*)calloc((size_t)3 * nhs + 1, sizeof(double));
    c->cnt_rg = (int*)calloc((size_t)nhs + 1, sizeof(int));
    c->uf_root = (int*)malloc((size_t)(V + 1) * sizeof(int));
    c->uf_next = (int*)malloc((size_t)(V + 1) * sizeof(int));
    c->uf_tail = (int*)malloc((size_t)(V + 1) * sizeof(int));
    c->uf_weight = (int*)malloc((size_t)(V + 1) * sizeof(int));
    c->sorted = (unsigned char*)calloc((size_t)V + 1, 1);
    c->tri_len0 = (double*)calloc((size_t)3 * T + 1, sizeof(double));
    c->bond_len0 = (double*)calloc((size_t)B + 1, sizeof(double));
    return c;
}

void orc_destroy(orc_ctx* c)
{
    if (!c) return;
    free(c->tri); free(c->tri_surf); free(c->bond); free(c->flags); free(c->vhs); free(c->hs_mass);
    free(c->xo); free(c->x); free(c->av); free(c->imp); free(c->fric); free(c->cnt); free(c->has);
    free(c->imp_rg); free(c->cnt_rg); free(c->uf_root); free(c->uf_next); free(c->uf_tail); free(c->uf_weight); free(c->sorted); free(c->tri_len0); free(c->bond_len0);
    free(c->cand); free(c->truep); free(c->con); free(c->mrg_com); free(c->mrg_valid);
    free(c);
}

void orc_set_params(orc_ctx* c, double eps, double thickness, double k, double m, double lambda, double cr)
{
    c->eps = eps; c->thickness = thickness; c->k = k; c->m = m; c->lambda = lambda; c->cr = cr;
}
void orc_set_domain(orc_ctx* c, const double* lo, const double* hi)
{
    for (int i = 0; i < 3; ++i) { c->lo[i] = lo[i]; c->hi[i] = hi[i]; }
}
void orc_set_dt(orc_ctx* c, double dt) { c->dt = dt; }

void orc_set_state(orc_ctx* c, const double* x_old, const double* x_new)
{
    size_t n = (size_t)3 * c->V;
    memcpy(c->xo, x_old, n * sizeof(double));
    memcpy(c->x, x_new, n * sizeof(double));
    memset(c->imp, 0, n * sizeof(double));
    memset(c->fric, 0, n * sizeof(double));
    memset(c->cnt, 0, (size_t)c->V * sizeof(int));
    memset(c->has, 0, (size_t)c->V); /* recordOriginPosition clears has_collsn, dcollid.cpp:100 */
    c->uf_ready = 0;                 /* makeSet in assembleFromInterface, dcollid3d.cpp:44 */
}
void orc_set_avgvel(orc_ctx* c, const double* av) { memcpy(c->av, av, (size_t)3 * c->V * sizeof(double)); }

void orc_get_f64(orc_ctx* c, int field, double* out)
{
    const double* src = field == 0 ? c->xo : field == 1 ? c->x : field == 2 ? c->av : field == 3 ? c->imp : c->fric;
    memcpy(out, src, (size_t)3 * c->V * sizeof(double));
}
void orc_get_i32(orc_ctx* c, int field, int* out)
{
    for (int v = 0; v < c->V; ++v) out[v] = field == 0 ? c->cnt[v] : (int)c->has[v];
}
void orc_get_body(orc_ctx* c, double* imp_rg, int* cnt_rg)
{
    memcpy(imp_rg, c->imp_rg, (size_t)3 * c->nhs * sizeof(double));
    memcpy(cnt_rg, c->cnt_rg, (size_t)c->nhs * sizeof(int));
}
void orc_set_body(orc_ctx* c, const double* imp_rg, const int* cnt_rg)
{
    memcpy(c->imp_rg, imp_rg, (size_t)3 * c->nhs * sizeof(double));
    memcpy(c->cnt_rg, cnt_rg, (size_t)c->nhs * sizeof(int));
}
long orc_num_candidates(orc_ctx* c) { return c->n_cand; }
void orc_get_candidates(orc_ctx* c, int* out) { memcpy(out, c->cand, (size_t)c->n_cand * 2 * sizeof(int)); }
long orc_num_contacts(orc_ctx* c) { return c->n_con; }
void orc_get_contacts(orc_ctx* c, orc_contact* out) { memcpy(out, c->con, (size_t)c->n_con * sizeof(orc_contact)); }
long orc_num_true_pairs(orc_ctx* c) { return c->n_true; }
void orc_get_true_pairs(orc_ctx* c, int* out) { memcpy(out, c->truep, (size_t)c->n_true * 2 * sizeof(int)); }

int orc_feature(int kind, const double* x_old, const double* coords, const double* avgvel,
                const unsigned char* flags, const double* mass, double h, double dt, const double* params,
                double* roots_out, double* acc, double* hit_root)
{
    int vhs[4] = {0, 1, 2, 3};
    orc_ctx* c = orc_create(4, 0, NULL, NULL, 0, NULL, flags, vhs, 4, mass);
    orc_set_params(c, params[0], params[1], params[2], params[3], params[4], params[5]);
    c->dt = dt;
    memcpy(c->xo, x_old, 12 * sizeof(double));
    memcpy(c->av, avgvel, 12 * sizeof(double));
    int p[4] = {0, 1, 2, 3};
    double roots[4] = {-1, -1, -1, dt};
    double X[4][3];
    memcpy(X, coords, 12 * sizeof(double));
    int ret = 0;
    *hit_root = -1.0;
    switch (kind) {
    case 0: ret = is_coplanar(c, p, dt, roots); break;
    case 1: ret = point_to_tri(c, p, X, h, 0.0, NULL); break;
    case 2: ret = edge_to_edge(c, p, X, h, 0.0, NULL); break;
    case 3:
    case 4:
        is_coplanar(c, p, dt, roots);
        ret = moving_test(c, kind == 4, p, h, NULL, hit_root);
        break;
    default: ret = -2;
    }
    for (int i = 0; i < 4; ++i) roots_out[i] = roots[i];
    for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < 3; ++j) {
            acc[10 * i + j] = c->imp[3 * i + j];
            acc[10 * i + 3 + j] = c->fric[3 * i + j];
            acc[10 * i + 6 + j] = c->imp_rg[3 * i + j];
        }
        acc[10 * i + 9] = (double)(c->cnt[i] + 1000 * c->cnt_rg[i]);
    }
    if (c->error) ret = -1;
    orc_destroy(c);
    return ret;
}
