/* C-ABI harness around the UNMODIFIED reference sources  (TEST INFRASTRUCTURE).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load the library built from this file (oracle/_ref/libcollision_ref.so).
 * The product (collision_b200/) never links, imports or executes it.
 *
 * What it does: builds the FronTier-shaped object graph (POINT/TRI/BOND/SURFACE/
 * CURVE/INTERFACE + STATE, see standin/) from flat arrays, then drives the
 * reference's own CollisionSolver3d on it:
 *   - whole step:      assembleFromInterface + resolveCollision   (test.cpp:105-107)
 *   - single phases:   the private members resolveCollision calls (dcollid.cpp:317-362),
 *                      reached with the usual "#define private public" test trick
 *   - candidate pairs: AABB.cpp is compiled with -DisProximity=clsnHookProximity
 *                      -DisCollision=clsnHookCollision, so every narrow-phase callback
 *                      the reference's tree query makes (AABB.cpp:296,326) lands in the
 *                      two hooks below, which record (a, b, result) and forward to the
 *                      real CollisionSolver::isProximity / isCollision
 *   - single features: dcollid3d.cpp is textually included so its file-static
 *                      isCoplanar / PointToTri / EdgeToEdge / Moving* are callable
 *                      for known-answer vectors (tests/golden/).
 * No reference source is copied into this repository; the files are compiled from
 * /root/reference where they lie (see oracle/Makefile).
 */
#include <vector>
#include <iostream>
#include <fstream>
#include <algorithm>
#include <functional>
#include <map>
#include <set>
#include <unordered_set>
#include <unordered_map>
#include <stack>
#include <string>
#include <chrono>
#include <cmath>
#include <cstring>
#include <cstdio>
#include <stdexcept>
#define private public
#define protected public
#include "dcollid3d.cpp" /* found through -I/root/reference */
#undef private
#undef protected
#include "AABB.h"

#include <map>
#include <string>
#include <chrono>

/* ------------------------------------------------------------------ */
/* FronTier runtime services used by the reference                      */
/* ------------------------------------------------------------------ */
static bool g_debug_collision = false;
static std::map<std::string, double> g_clock_total;
static std::map<std::string, double> g_clock_start;

static double now_seconds()
{
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}

bool debugging(const char* s) { return g_debug_collision && std::strcmp(s, "collision") == 0; }
void clean_up(int) { throw clsn_ref_abort(); }
void start_clock(const char* s) { g_clock_start[s] = now_seconds(); }
void stop_clock(const char* s) { g_clock_total[s] += now_seconds() - g_clock_start[s]; }
double cpu_seconds() { return now_seconds(); }
bool create_directory(const char*, int) { return false; }

/* ------------------------------------------------------------------ */
/* Mesh container                                                       */
/* ------------------------------------------------------------------ */
struct RefCtx {
    int V, T, B, nhs;
    std::vector<POINT> pts;
    std::vector<STATE> st;
    std::vector<TRI> tris;
    std::vector<BOND> bonds;
    std::vector<HYPER_SURF> hss;
    std::vector<SURFACE> surfs;
    std::vector<CURVE> curves;
    std::vector<SURFACE*> surf_ptrs;
    std::vector<CURVE*> curve_ptrs;
    Table table;
    INTERFACE intfc;
    CollisionSolver3d* solver;
    double eps, thickness, k, m, lambda, cr;
    /* hook recording */
    bool record;
    std::vector<int> pairs; /* (a, b, result) triples, element = hseList index */
    long n_callbacks;
};

static RefCtx* g_active = nullptr;

static void push_params(RefCtx* c)
{
    CollisionSolver::setRoundingTolerance(c->eps);
    CollisionSolver::setFabricThickness(c->thickness);
    CollisionSolver::setSpringConstant(c->k);
    CollisionSolver::setPointMass(c->m);
    CollisionSolver::setFrictionConstant(c->lambda);
    CollisionSolver::setRestitutionCoef(c->cr);
}

static int element_index(const RefCtx* c, const CD_HSE* h)
{
    if (const CD_TRI* t = dynamic_cast<const CD_TRI*>(h)) return (int)(t->m_tri - c->tris.data());
    if (const CD_BOND* b = dynamic_cast<const CD_BOND*>(h))
        return c->T + (int)(b->m_bond - c->bonds.data());
    return -1;
}

/* The two symbols AABB.o calls instead of isProximity/isCollision (see header comment).
 * They are ordinary member functions of CollisionSolver as far as the ABI goes. */
extern "C" bool clsn_hook_prox(CollisionSolver* self, const CD_HSE* a, const CD_HSE* b)
    __asm__("_ZN15CollisionSolver17clsnHookProximityEPK6CD_HSES2_");
extern "C" bool clsn_hook_coll(CollisionSolver* self, const CD_HSE* a, const CD_HSE* b)
    __asm__("_ZN15CollisionSolver17clsnHookCollisionEPK6CD_HSES2_");

extern "C" bool clsn_hook_prox(CollisionSolver* self, const CD_HSE* a, const CD_HSE* b)
{
    bool r = self->isProximity(a, b);
    RefCtx* c = g_active;
    if (c) {
        c->n_callbacks++;
        if (c->record) {
            c->pairs.push_back(element_index(c, a));
            c->pairs.push_back(element_index(c, b));
            c->pairs.push_back(r ? 1 : 0);
        }
    }
    return r;
}
extern "C" bool clsn_hook_coll(CollisionSolver* self, const CD_HSE* a, const CD_HSE* b)
{
    bool r = self->isCollision(a, b);
    RefCtx* c = g_active;
    if (c) {
        c->n_callbacks++;
        if (c->record) {
            c->pairs.push_back(element_index(c, a));
            c->pairs.push_back(element_index(c, b));
            c->pairs.push_back(r ? 1 : 0);
        }
    }
    return r;
}

struct ActiveGuard {
    ActiveGuard(RefCtx* c)
    {
        g_active = c;
        push_params(c);
    }
    ~ActiveGuard()
    {
        g_active = nullptr;
        fedisableexcept(FE_ALL_EXCEPT); /* resolveCollision enables FP traps (dcollid.cpp:320) */
        feclearexcept(FE_ALL_EXCEPT);
    }
};

extern "C" {

/* hs_kind: 0 fabric, 1 static rigid (NEUMANN_BOUNDARY), 2 movable rigid body.
 * Tris must be grouped by non-decreasing tri_surf (surface s uses hyper-surface s);
 * bonds by non-decreasing bond_curve (curve c uses hyper-surface n_surf + c).
 * vhs[v] = hyper-surface index of vertex v.  vflags bit0 = is_fixed, bit1 = is_movableRG. */
void* clsn_ref_create(int V, int T, const int* tri_idx, const int* tri_surf, int n_surf, int B,
                      const int* bond_idx, const int* bond_curve, int n_curve, const int* hs_kind,
                      const double* hs_mass, const unsigned char* vflags, const int* vhs)
{
    RefCtx* c = new RefCtx();
    c->V = V;
    c->T = T;
    c->B = B;
    c->nhs = n_surf + n_curve;
    c->pts.resize(V);
    c->st.resize(V);
    c->tris.resize(T);
    c->bonds.resize(B);
    c->hss.resize(c->nhs);
    c->surfs.resize(n_surf);
    c->curves.resize(n_curve);
    c->eps = 1e-6;
    c->thickness = 1e-4;
    c->k = 1000;
    c->m = 0.01;
    c->lambda = 0.02;
    c->cr = 0.0;
    c->record = false;
    c->n_callbacks = 0;
    for (int h = 0; h < c->nhs; ++h) {
        HYPER_SURF& hs = c->hss[h];
        std::memset(&hs, 0, sizeof(hs));
        hs.wave_type = hs_kind[h] == 1 ? NEUMANN_BOUNDARY
                                       : (hs_kind[h] == 2 ? MOVABLE_BODY_BOUNDARY : FIRST_PHYSICS_WAVE_TYPE);
        hs.body_index = h;
        hs.total_mass = hs_mass[h];
    }
    for (int v = 0; v < V; ++v) {
        POINT& p = c->pts[v];
        STATE& s = c->st[v];
        std::memset(&p, 0, sizeof(p));
        std::memset(&s, 0, sizeof(s));
        p.global_index = v;
        p.indx = v;
        p._left_state = &s;
        p._right_state = &s;
        p.hs = &c->hss[vhs[v]];
        s.is_fixed = (vflags[v] & 1) != 0;
        s.is_movableRG = (vflags[v] & 2) != 0;
    }
    for (int s = 0; s < n_surf; ++s) {
        c->surfs[s].hyper_surf = &c->hss[s];
        c->surfs[s]._first_tri = nullptr;
        c->surfs[s]._is_bdry = 0;
    }
    {
        std::vector<TRI*> last(n_surf, nullptr);
        for (int t = 0; t < T; ++t) {
            TRI& tr = c->tris[t];
            for (int i = 0; i < 3; ++i) tr.__pts[i] = &c->pts[tri_idx[3 * t + i]];
            int s = tri_surf[t];
            tr.surf = &c->surfs[s];
            tr.next = nullptr;
            tr.prev = last[s];
            if (last[s]) last[s]->next = &tr;
            else c->surfs[s]._first_tri = &tr;
            last[s] = &tr;
            for (int i = 0; i < 3; ++i) tr.side_length0[i] = 0.0;
        }
    }
    for (int q = 0; q < n_curve; ++q) {
        c->curves[q].first = c->curves[q].last = nullptr;
        c->curves[q]._hsbdry_type = STRING_HSBDRY;
    }
    for (int b = 0; b < B; ++b) {
        BOND& bd = c->bonds[b];
        bd.start = &c->pts[bond_idx[2 * b]];
        bd.end = &c->pts[bond_idx[2 * b + 1]];
        CURVE& cv = c->curves[bond_curve[b]];
        bd.next = nullptr;
        bd.prev = cv.last;
        if (cv.last) cv.last->next = &bd;
        else cv.first = &bd;
        cv.last = &bd;
        bd.length0 = 0.0;
    }
    for (int s = 0; s < n_surf; ++s) c->surf_ptrs.push_back(&c->surfs[s]);
    c->surf_ptrs.push_back(nullptr);
    for (int q = 0; q < n_curve; ++q) c->curve_ptrs.push_back(&c->curves[q]);
    c->curve_ptrs.push_back(nullptr);
    for (int i = 0; i < 3; ++i) {
        c->table.rect_grid.L[i] = -1e30;
        c->table.rect_grid.U[i] = 1e30;
    }
    c->intfc.surfaces = c->surf_ptrs.data();
    c->intfc.curves = c->curve_ptrs.data();
    c->intfc.table = &c->table;
    c->solver = new CollisionSolver3d();
    return c;
}

void clsn_ref_destroy(void* h)
{
    RefCtx* c = (RefCtx*)h;
    delete c->solver;
    delete c;
}

void clsn_ref_set_params(void* h, double eps, double thickness, double k, double m, double lambda,
                         double cr)
{
    RefCtx* c = (RefCtx*)h;
    c->eps = eps;
    c->thickness = thickness;
    c->k = k;
    c->m = m;
    c->lambda = lambda;
    c->cr = cr;
}

void clsn_ref_set_domain(void* h, const double* L, const double* U)
{
    RefCtx* c = (RefCtx*)h;
    for (int i = 0; i < 3; ++i) {
        c->table.rect_grid.L[i] = L[i];
        c->table.rect_grid.U[i] = U[i];
    }
}

/* Rest lengths for the strain limiter (dcollid.cpp:513-516): measured on the given positions. */
void clsn_ref_set_rest_lengths(void* h, const double* x)
{
    RefCtx* c = (RefCtx*)h;
    for (int t = 0; t < c->T; ++t) {
        TRI& tr = c->tris[t];
        for (int j = 0; j < 3; ++j) {
            long a = tr.__pts[j]->global_index, b = tr.__pts[(j + 1) % 3]->global_index;
            tr.side_length0[j] = distance_between_positions(x + 3 * a, x + 3 * b, 3);
        }
    }
    for (int b = 0; b < c->B; ++b) {
        BOND& bd = c->bonds[b];
        bd.length0 = distance_between_positions(x + 3 * bd.start->global_index,
                                                x + 3 * bd.end->global_index, 3);
    }
}

/* Start-of-step state, as FT_Propagate's point hook + the spring solver leave it
 * (test.cpp:192-258): x_old, candidate Coords, vel; per-step accumulators cleared. */
void clsn_ref_set_state(void* h, const double* x_old, const double* x_new, const double* vel)
{
    RefCtx* c = (RefCtx*)h;
    for (int v = 0; v < c->V; ++v) {
        POINT& p = c->pts[v];
        STATE& s = c->st[v];
        for (int j = 0; j < 3; ++j) {
            s.x_old[j] = x_old[3 * v + j];
            p._coords[j] = x_new[3 * v + j];
            if (vel) {
                s.vel[j] = vel[3 * v + j];
                p.vel[j] = vel[3 * v + j];
            }
            s.collsnImpulse[j] = 0.0;
            s.friction[j] = 0.0;
        }
        s.collsn_num = 0;
    }
}

/* field ids: 0 x_old, 1 Coords, 2 vel, 3 avgVel, 4 collsnImpulse, 5 friction, 6 collsnImpulse_RG */
static double* dfield(RefCtx* c, int v, int f)
{
    STATE& s = c->st[v];
    switch (f) {
    case 0: return s.x_old;
    case 1: return c->pts[v]._coords;
    case 2: return s.vel;
    case 3: return s.avgVel;
    case 4: return s.collsnImpulse;
    case 5: return s.friction;
    case 6: return s.collsnImpulse_RG;
    }
    return nullptr;
}
int clsn_ref_get_f64(void* h, int field, double* out)
{
    RefCtx* c = (RefCtx*)h;
    if (field < 0 || field > 6) return -1;
    for (int v = 0; v < c->V; ++v)
        for (int j = 0; j < 3; ++j) out[3 * v + j] = dfield(c, v, field)[j];
    return 0;
}
int clsn_ref_set_f64(void* h, int field, const double* in)
{
    RefCtx* c = (RefCtx*)h;
    if (field < 0 || field > 6) return -1;
    for (int v = 0; v < c->V; ++v)
        for (int j = 0; j < 3; ++j) dfield(c, v, field)[j] = in[3 * v + j];
    return 0;
}
/* int field ids: 0 collsn_num, 1 collsn_num_RG, 2 has_collsn */
int clsn_ref_get_i32(void* h, int field, int* out)
{
    RefCtx* c = (RefCtx*)h;
    for (int v = 0; v < c->V; ++v) {
        STATE& s = c->st[v];
        out[v] = field == 0 ? s.collsn_num : (field == 1 ? s.collsn_num_RG : (int)s.has_collsn);
    }
    return 0;
}
int clsn_ref_set_i32(void* h, int field, const int* in)
{
    RefCtx* c = (RefCtx*)h;
    for (int v = 0; v < c->V; ++v) {
        STATE& s = c->st[v];
        if (field == 0) s.collsn_num = in[v];
        else if (field == 1) s.collsn_num_RG = in[v];
        else s.has_collsn = in[v] != 0;
    }
    return 0;
}

int clsn_ref_assemble(void* h, double dt)
{
    RefCtx* c = (RefCtx*)h;
    ActiveGuard g(c);
    try {
        /* updateFinalForRG reads mrg_com[body] before anything wrote it when a movable body
         * collides in the very first step (dcollid.cpp:649-650, an empty std::vector).  Seed it
         * the way a preceding collision-free step would have (dcollid.cpp:667-672). */
        for (int i = 0; i < c->nhs; ++i)
            if (c->hss[i].wave_type == MOVABLE_BODY_BOUNDARY && c->solver->mrg_com.count(i) == 0)
                c->solver->mrg_com[i] = std::vector<double>(c->hss[i].center_of_mass,
                                                            c->hss[i].center_of_mass + 3);
        c->solver->assembleFromInterface(&c->intfc, dt);
    } catch (const clsn_ref_abort&) {
        return -1;
    }
    return 0;
}

/* centre of mass / its velocity of hyper-surface i (caller-owned in the reference). */
void clsn_ref_set_body(void* h, int i, const double* com, const double* com_velo)
{
    RefCtx* c = (RefCtx*)h;
    for (int j = 0; j < 3; ++j) {
        c->hss[i].center_of_mass[j] = com[j];
        c->hss[i].center_of_mass_velo[j] = com_velo[j];
    }
}
void clsn_ref_get_body(void* h, int i, double* com, double* com_velo)
{
    RefCtx* c = (RefCtx*)h;
    for (int j = 0; j < 3; ++j) {
        com[j] = c->hss[i].center_of_mass[j];
        com_velo[j] = c->hss[i].center_of_mass_velo[j];
    }
}

void clsn_ref_record_origin(void* h)
{
    RefCtx* c = (RefCtx*)h;
    ActiveGuard g(c);
    c->solver->recordOriginPosition();
}

/* Whole step.  strain_limiting != 0: the reference's resolveCollision() verbatim.
 * strain_limiting == 0: the same member calls in the same order (dcollid.cpp:317-362)
 * minus reduceSuperelast. */
int clsn_ref_resolve(void* h, int strain_limiting)
{
    RefCtx* c = (RefCtx*)h;
    ActiveGuard g(c);
    try {
        if (strain_limiting) {
            c->solver->resolveCollision();
        } else {
            CollisionSolver* s = c->solver;
            s->setTraitsDimension();
            s->computeAverageVelocity();
            s->detectProximity();
            s->detectCollision();
            s->detectDomainBoundaryCollision();
            s->updateFinalPosition();
            s->updateFinalVelocity();
        }
    } catch (const clsn_ref_abort&) {
        return -1;
    }
    return 0;
}

/* Single phases, for per-pass parity from identical inputs.
 * 0 computeAverageVelocity   1 aabbProximity (tree + query, accumulators only)
 * 2 updateAverageVelocity    3 aabbCollision (tree + query, accumulators only)
 * 4 detectDomainBoundaryCollision  5 updateFinalPosition  6 reduceSuperelast
 * 7 updateFinalVelocity      8 detectProximity  9 detectCollision
 * Returns the tree's callback-true count for 1 and 3, has_collision for 9, else 0; -1 on abort. */
int clsn_ref_phase(void* h, int phase)
{
    RefCtx* c = (RefCtx*)h;
    ActiveGuard g(c);
    CollisionSolver* s = c->solver;
    int ret = 0;
    try {
        switch (phase) {
        case 0: s->computeAverageVelocity(); break;
        case 1:
            s->aabbProximity();
            ret = s->abt_proximity->getCount();
            break;
        case 2: s->updateAverageVelocity(); break;
        case 3:
            s->aabbCollision();
            ret = s->abt_collision->getCount();
            break;
        case 4: s->detectDomainBoundaryCollision(); break;
        case 5: s->updateFinalPosition(); break;
        case 6: s->reduceSuperelast(); break;
        case 7: s->updateFinalVelocity(); break;
        case 8: s->detectProximity(); break;
        case 9:
            s->detectCollision();
            ret = s->hasCollision() ? 1 : 0;
            break;
        case 10: s->turnOnImpZone(); break;
        case 11: s->turnOffImpZone(); break;
        case 12: s->updateImpactZoneVelocity(ret); break;
        case 13: s->computeImpactZone(); break;
        default: return -2;
        }
    } catch (const clsn_ref_abort&) {
        return -1;
    }
    return ret;
}

void clsn_ref_record(void* h, int on)
{
    RefCtx* c = (RefCtx*)h;
    c->record = on != 0;
    c->pairs.clear();
    c->n_callbacks = 0;
}
long clsn_ref_num_callbacks(void* h) { return ((RefCtx*)h)->n_callbacks; }
long clsn_ref_num_pairs(void* h) { return (long)(((RefCtx*)h)->pairs.size() / 3); }
void clsn_ref_get_pairs(void* h, int* out)
{
    RefCtx* c = (RefCtx*)h;
    std::memcpy(out, c->pairs.data(), c->pairs.size() * sizeof(int));
}

double clsn_ref_clock(const char* name)
{
    auto it = g_clock_total.find(name);
    return it == g_clock_total.end() ? 0.0 : it->second;
}
void clsn_ref_clock_reset() { g_clock_total.clear(); }
void clsn_ref_debug(int on) { g_debug_collision = on != 0; }

/* ------------------------------------------------------------------ */
/* Feature-level known-answer entry: one call of a file-static           */
/* primitive of dcollid3d.cpp on four free-standing points.             */
/* ------------------------------------------------------------------ */
/* kind: 0 isCoplanar  1 PointToTri  2 EdgeToEdge  3 MovingPointToTri  4 MovingEdgeToEdge
 * in : x_old[12], coords[12] (used by the static tests 1,2), avgVel[12], flags[4],
 *      mass[4] (total_mass of each point's hyper-surface), h, dt, params via set_params-like args
 * out: ret; roots[4]; acc[4][10] = imp[3], fric[3], impRG[3], (cnt + 1000*cntRG) per point;
 *      hit_root = the root at which the moving test fired (-1 if none). */
int clsn_ref_feature(int kind, const double* x_old, const double* coords, const double* avgVel,
                     const unsigned char* flags, const double* mass, double h, double dt,
                     const double* params /* eps,thickness,k,m,lambda,cr */, double* roots_out,
                     double* acc_out, double* hit_root)
{
    POINT p[4];
    STATE s[4];
    HYPER_SURF hs[4];
    POINT* pts[4];
    for (int i = 0; i < 4; ++i) {
        std::memset(&p[i], 0, sizeof(POINT));
        std::memset(&s[i], 0, sizeof(STATE));
        std::memset(&hs[i], 0, sizeof(HYPER_SURF));
        hs[i].total_mass = mass[i];
        p[i].hs = &hs[i];
        p[i]._left_state = &s[i];
        p[i]._right_state = &s[i];
        p[i].global_index = i;
        s[i].is_fixed = (flags[i] & 1) != 0;
        s[i].is_movableRG = (flags[i] & 2) != 0;
        s[i].impZone.root = &p[i];
        s[i].impZone.tail = &p[i];
        s[i].impZone.next_pt = nullptr;
        s[i].impZone.num_pts = 1;
        for (int j = 0; j < 3; ++j) {
            s[i].x_old[j] = x_old[3 * i + j];
            s[i].avgVel[j] = avgVel[3 * i + j];
            p[i]._coords[j] = coords[3 * i + j];
        }
        pts[i] = &p[i];
    }
    CollisionSolver::setRoundingTolerance(params[0]);
    CollisionSolver::setFabricThickness(params[1]);
    CollisionSolver::setSpringConstant(params[2]);
    CollisionSolver::setPointMass(params[3]);
    CollisionSolver::setFrictionConstant(params[4]);
    CollisionSolver::setRestitutionCoef(params[5]);
    CollisionSolver::setTimeStepSize(dt);
    double roots[4] = {-1, -1, -1, dt};
    int ret = 0;
    *hit_root = -1.0;
    try {
        switch (kind) {
        case 0: ret = isCoplanar(pts, dt, roots) ? 1 : 0; break;
        case 1: ret = PointToTri(pts, h, 0.0) ? 1 : 0; break;
        case 2: ret = EdgeToEdge(pts, h, 0.0) ? 1 : 0; break;
        case 3:
        case 4: {
            /* the hit root is not stored anywhere by the reference (only printed under a debug
             * string, dcollid3d.cpp:1026,1208), so replay the root loop of Moving*ToTri/Edge
             * (dcollid3d.cpp:327-369) on a scratch copy first to learn it ... */
            POINT q[4];
            STATE qs[4];
            POINT* qp[4];
            for (int i = 0; i < 4; ++i) {
                q[i] = p[i];
                qs[i] = s[i];
                q[i]._left_state = &qs[i];
                q[i]._right_state = &qs[i];
                qs[i].impZone.root = qs[i].impZone.tail = &q[i];
                qp[i] = &q[i];
            }
            if (isCoplanar(qp, dt, roots)) {
                for (int i = 0; i < 4; ++i) {
                    if (roots[i] < 0) continue;
                    for (int j = 0; j < 4; ++j)
                        for (int k = 0; k < 3; ++k)
                            q[j]._coords[k] = qs[j].x_old[k] + roots[i] * qs[j].avgVel[k];
                    bool hit = kind == 3 ? PointToTri(qp, h, roots[i]) : EdgeToEdge(qp, h, roots[i]);
                    if (hit) {
                        *hit_root = roots[i];
                        break;
                    }
                }
            }
            /* ... then run the reference's own function for the answer that is compared. */
            ret = (kind == 3 ? MovingPointToTri(pts, h) : MovingEdgeToEdge(pts, h)) ? 1 : 0;
            if ((ret != 0) != (*hit_root >= 0)) return -3;
            break;
        }
        default: return -2;
        }
    } catch (const clsn_ref_abort&) {
        return -1;
    }
    for (int i = 0; i < 4; ++i) roots_out[i] = roots[i];
    for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < 3; ++j) {
            acc_out[10 * i + j] = s[i].collsnImpulse[j];
            acc_out[10 * i + 3 + j] = s[i].friction[j];
            acc_out[10 * i + 6 + j] = s[i].collsnImpulse_RG[j];
        }
        acc_out[10 * i + 9] = (double)(s[i].collsn_num + 1000 * s[i].collsn_num_RG);
    }
    fedisableexcept(FE_ALL_EXCEPT);
    return ret;
}

/* n calls of clsn_ref_feature in one go (the comparison against the reference's primitives makes ~10^5 of them per
 * pass): pts[4n] indexes the per-vertex arrays x_old/avgVel[3V], flags[V], mass[V]; Coords = x_old on entry.
 * ret[n] (negative = that call aborted), hit_root[n]; returns the number of calls that returned true. */
long clsn_ref_feature_batch(long n, const int* kind, const int* pts, const double* x_old, const double* avgVel,
                            const unsigned char* flags, const double* mass, double h, double dt, const double* params,
                            int* ret, double* hit_root)
{
    long hits = 0;
    for (long c = 0; c < n; ++c) {
        double xo[12], av[12], m[4], roots[4], acc[40];
        unsigned char fl[4];
        for (int i = 0; i < 4; ++i) {
            const int p = pts[4 * c + i];
            for (int j = 0; j < 3; ++j) {
                xo[3 * i + j] = x_old[3 * (long)p + j];
                av[3 * i + j] = avgVel[3 * (long)p + j];
            }
            fl[i] = flags[p];
            m[i] = mass[p];
        }
        ret[c] = clsn_ref_feature(kind[c], xo, xo, av, fl, m, h, dt, params, roots, acc, &hit_root[c]);
        if (ret[c] > 0) ++hits;
    }
    return hits;
}

} /* extern "C" */
