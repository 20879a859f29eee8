// Host mirror of the reference's CollisionSolver over the C ABI -- see collid_b200.h.
#include "collid_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>

namespace clsn_host {

double CollisionSolver::s_eps = 1e-6;        // dcollid.cpp:29-35
double CollisionSolver::s_thickness = 0.0001;
double CollisionSolver::s_dt = 0.001;
double CollisionSolver::s_k = 1000;
double CollisionSolver::s_m = 0.01;
double CollisionSolver::s_lambda = 0.02;
double CollisionSolver::s_cr = 0.0;

void CollisionSolver::setRoundingTolerance(double v) { s_eps = v; }
double CollisionSolver::getRoundingTolerance() { return s_eps; }
void CollisionSolver::setFabricThickness(double v) { s_thickness = v; }
double CollisionSolver::getFabricThickness() { return s_thickness; }
void CollisionSolver::setTimeStepSize(double v) { s_dt = v; }
double CollisionSolver::getTimeStepSize() { return s_dt; }
void CollisionSolver::setSpringConstant(double v) { s_k = v; }
double CollisionSolver::getSpringConstant() { return s_k; }
void CollisionSolver::setFrictionConstant(double v) { s_lambda = v; }
double CollisionSolver::getFrictionConstant() { return s_lambda; }
void CollisionSolver::setPointMass(double v) { s_m = v; }
double CollisionSolver::getPointMass() { return s_m; }
void CollisionSolver::setRestitutionCoef(double v) { s_cr = v; }
double CollisionSolver::getRestitutionCoef() { return s_cr; }

bool isStaticRigidBody(const POINT* p) { return p->state->is_fixed; }
bool isMovableRigidBody(const POINT* p) { return p->state->is_movableRG; }
bool isRigidBody(const POINT* p) { return isStaticRigidBody(p) || isMovableRigidBody(p); }

CollisionSolver::CollisionSolver(int dim, int device)
    : has_collision(false), m_dim(dim), m_ctx(nullptr), m_device(device), m_pair_ctx(nullptr), m_topology_dirty(true)
{
    for (int i = 0; i < 3; ++i) { Boundary[i][0] = -1e30; Boundary[i][1] = 1e30; }
    std::memset(&m_stats, 0, sizeof(m_stats));
    std::memset(&m_zone_stats, 0, sizeof(m_zone_stats));
    int rc = clsn_create(&m_ctx, device);
    if (rc != CLSN_OK || !m_ctx) throw std::runtime_error("collision_b200: no usable CUDA device (there is no CPU fallback)");
    clsn_set_impact_zones(m_ctx, 1, 0);
    clsn_set_strain_limiting(m_ctx, 1);
}

void CollisionSolver::multiGPUUniqueId(unsigned char id[128])
{
    if (clsn_dist_unique_id(id) != CLSN_OK) throw std::runtime_error("collision_b200: NCCL is not available (clsn_dist_unique_id)");
}

void CollisionSolver::enableMultiGPU(int rank, int nranks, const unsigned char id[128])
{
    if (m_points.empty()) {   // the library wants the topology first: remember, join after the first assembleFromInterface
        m_dist_rank = rank;
        m_dist_nranks = nranks;
        std::memcpy(m_dist_id, id, 128);
        return;
    }
    int rc = clsn_dist_init(m_ctx, rank, nranks, id);
    if (rc != CLSN_OK) fail(rc, "clsn_dist_init");
    m_dist_nranks = 0;
}

void CollisionSolver::printDebugVariable() const
{
    long coplanar = 0, features = 0, contacts = 0;
    for (int i = 0; i < m_stats.n_ccd_passes; ++i) {
        coplanar += (long)m_stats.ccd[i].coplanar;
        features += (long)m_stats.ccd[i].features;
        contacts += (long)m_stats.ccd[i].contacts;
    }
    std::printf("    %ld isCoplanar true, %ld CCD feature tests after the culls, %ld contacts, %ld proximity contacts\n", coplanar, features,
                contacts, (long)m_stats.proximity.contacts);
}

void CollisionSolver::setStrainLimiting(bool on)
{
    int rc = clsn_set_strain_limiting(m_ctx, on ? 1 : 0);
    if (rc != CLSN_OK) fail(rc, "clsn_set_strain_limiting");
}

void CollisionSolver::setImpactZones(bool on, int max_iter)
{
    int rc = clsn_set_impact_zones(m_ctx, on ? 1 : 0, max_iter);
    if (rc != CLSN_OK) fail(rc, "clsn_set_impact_zones");
}

void CollisionSolver::computeImpactZone()  // dcollid.cpp:227-265, on the state resident on the GPU
{
    int rc = clsn_compute_impact_zone(m_ctx, 0, &m_zone_stats);
    if (rc != CLSN_OK) fail(rc, "clsn_compute_impact_zone");
}

CollisionSolver::~CollisionSolver()
{
    clearHseList();
    clsn_destroy(m_pair_ctx);
    clsn_destroy(m_ctx);
}

void CollisionSolver::fail(int rc, const char* where) const
{
    // the reference's clean_up(ERROR) ends the process; a library must not, so this throws
    char buf[512];
    std::snprintf(buf, sizeof(buf), "collision_b200 %s failed (%d): %s", where, rc, clsn_last_error(m_ctx));
    throw std::runtime_error(buf);
}

void CollisionSolver::clearHseList()
{
    for (CD_HSE* h : hseList) delete h;
    hseList.clear();
}

void CollisionSolver::setDomainBoundary(double* L, double* U)
{
    for (int i = 0; i < m_dim; ++i) { Boundary[i][0] = L[i]; Boundary[i][1] = U[i]; }
}

void CollisionSolver::recordOriginPosition()  // dcollid.cpp:91-107
{
    for (POINT* p : m_points) {
        p->state->has_collsn = false;
        if (isMovableRigidBody(p)) continue;
        for (int j = 0; j < 3; ++j) p->state->x_old[j] = p->coords[j];
    }
}

// Flatten the element list: vertex ids by first appearance, triangles then bonds (hseList order).
void CollisionSolver::gatherTopology(const INTERFACE* intfc)
{
    std::vector<POINT*> points;
    std::unordered_map<POINT*, int> ids;
    std::unordered_map<const HYPER_SURF*, int> body_of;
    std::vector<double> body_mass;
    std::vector<int32_t> tri, tri_surf, bond, body;
    std::vector<double> tri_len0, bond_len0;  // TRI::side_length0 / BOND::length0, read by reduceSuperelastOnce (dcollid.cpp:514-517)
    std::vector<uint8_t> flags;
    auto pid = [&](POINT* p) {
        auto it = ids.find(p);
        if (it != ids.end()) return it->second;
        int id = (int)points.size();
        ids.emplace(p, id);
        points.push_back(p);
        flags.push_back((uint8_t)((p->state->is_fixed ? CLSN_VFLAG_FIXED : 0) | (p->state->is_movableRG ? CLSN_VFLAG_MOVABLE_RG : 0)));
        auto b = body_of.find(p->hs);
        if (b == body_of.end()) {
            b = body_of.emplace(p->hs, (int)body_mass.size()).first;
            body_mass.push_back(p->hs ? p->hs->total_mass : 0.0);
        }
        body.push_back(b->second);
        return id;
    };
    std::unordered_map<const SURFACE*, int> surf_id;
    for (CD_HSE* h : hseList) {
        if (CD_TRI* t = dynamic_cast<CD_TRI*>(h)) {
            for (int i = 0; i < 3; ++i) tri.push_back(pid(t->m_tri->pts[i]));
            for (int i = 0; i < 3; ++i) tri_len0.push_back(t->m_tri->side_length0[i]);
            auto s = surf_id.emplace(t->m_tri->surf, (int)surf_id.size()).first;
            tri_surf.push_back(s->second);
        }
    }
    for (CD_HSE* h : hseList) {
        if (CD_BOND* b = dynamic_cast<CD_BOND*>(h)) {
            bond.push_back(pid(b->m_bond->start));
            bond.push_back(pid(b->m_bond->end));
            bond_len0.push_back(b->m_bond->length0);
        }
    }
    (void)intfc;
    const bool same = points == m_points && tri == m_tri && bond == m_bond && flags == m_flags && body == m_body &&
                      tri_surf == m_tri_surf && body_mass == m_body_mass;
    const bool same_len0 = tri_len0 == m_tri_len0 && bond_len0 == m_bond_len0;
    if (same && !m_topology_dirty) {
        if (!same_len0) {
            m_tri_len0.swap(tri_len0); m_bond_len0.swap(bond_len0);
            int rc = clsn_set_rest_lengths(m_ctx, m_tri_len0.data(), m_bond_len0.data());
            if (rc != CLSN_OK) fail(rc, "clsn_set_rest_lengths");
        }
        return;
    }
    m_points.swap(points); m_point_id.swap(ids); m_tri.swap(tri); m_tri_surf.swap(tri_surf); m_bond.swap(bond);
    m_flags.swap(flags); m_body.swap(body); m_body_mass.swap(body_mass);
    m_body_hs.assign(m_body_mass.size(), nullptr);
    for (auto& kv : body_of) m_body_hs[kv.second] = const_cast<HYPER_SURF*>(kv.first);
    const int V = (int)m_points.size();
    int rc = clsn_set_topology(m_ctx, V, (int)m_tri_surf.size(), m_tri.data(), m_tri_surf.data(), (int)m_bond.size() / 2,
                               m_bond.data(), m_flags.data(), m_body.data(), (int)m_body_mass.size(), m_body_mass.data());
    if (rc != CLSN_OK) fail(rc, "clsn_set_topology");
    m_tri_len0.swap(tri_len0); m_bond_len0.swap(bond_len0);
    rc = clsn_set_rest_lengths(m_ctx, m_tri_len0.data(), m_bond_len0.data());
    if (rc != CLSN_OK) fail(rc, "clsn_set_rest_lengths");
    if (!m_xold.resize(3 * (size_t)V) || !m_xnew.resize(3 * (size_t)V) || !m_xout.resize(3 * (size_t)V) ||
        !m_avg.resize(3 * (size_t)V) || !m_has.resize((size_t)V))
        fail(CLSN_E_NOMEM, "clsn_host_alloc");
    m_topology_dirty = false;
    if (m_dist_nranks > 0) enableMultiGPU(m_dist_rank, m_dist_nranks, m_dist_id);
}

void CollisionSolver3d::assembleFromInterface(const INTERFACE* intfc, double dt)  // dcollid3d.cpp:12-52
{
    setTimeStepSize(dt);
    clearHseList();
    for (SURFACE* s : intfc->surfaces) {
        if (s->is_bdry) continue;
        for (TRI* t = s->first_tri; t; t = t->next) hseList.push_back(new CD_TRI(t));
    }
    for (CURVE* c : intfc->curves) {
        if (!c->is_string) continue;
        for (BOND* b = c->first; b; b = b->next) hseList.push_back(new CD_BOND(b, m_dim));
    }
    gatherTopology(intfc);
    double L[3] = {intfc->L[0], intfc->L[1], intfc->L[2]}, U[3] = {intfc->U[0], intfc->U[1], intfc->U[2]};
    setDomainBoundary(L, U);
}

void CollisionSolver::resolveCollision()  // dcollid.cpp:317-362
{
    const size_t V = m_points.size();
    for (size_t v = 0; v < V; ++v) {
        const POINT* p = m_points[v];
        for (int j = 0; j < 3; ++j) {
            m_xold[3 * v + j] = p->state->x_old[j];
            m_xnew[3 * v + j] = p->coords[j];
        }
    }
    clsn_params prm;
    prm.eps = s_eps; prm.thickness = s_thickness; prm.dt = s_dt; prm.k = s_k; prm.m = s_m; prm.lambda = s_lambda; prm.cr = s_cr;
    for (int i = 0; i < 3; ++i) { prm.lo[i] = Boundary[i][0]; prm.hi[i] = Boundary[i][1]; }
    int rc = clsn_set_params(m_ctx, &prm);
    if (rc != CLSN_OK) fail(rc, "clsn_set_params");
    // upload, the whole step and the download of Coords / avgVel / has_collsn: one call, one synchronisation, pinned arrays
    rc = clsn_step_host_state(m_ctx, m_xold.data(), m_xnew.data(), m_xout.data(), m_avg.data(), m_has.data(), &m_stats);
    if (rc != CLSN_OK) fail(rc, "clsn_step_host_state");  // NaN/Inf: the reference calls clean_up(ERROR) here
    has_collision = m_stats.has_collision != 0;
    for (size_t v = 0; v < V; ++v) {
        POINT* p = m_points[v];
        STATE* sl = p->state;
        sl->has_collsn = m_has[v] != 0;
        for (int j = 0; j < 3; ++j) {
            p->coords[j] = m_xout[3 * v + j];      // updateFinalPosition, dcollid.cpp:562-584
            sl->avgVel[j] = m_avg[3 * v + j];
            if (sl->has_collsn) {                  // updateFinalVelocity, dcollid.cpp:598-624
                sl->vel[j] = sl->avgVel[j];
                p->vel[j] = sl->avgVel[j];
            }
            sl->collsnImpulse[j] = sl->friction[j] = 0.0;
        }
        sl->collsn_num = 0;
    }
    // updateFinalForRG, dcollid.cpp:626-675: centre of mass and its velocity of every movable body that was hit
    // (mrg_com lives inside the context); HYPER_SURF data is caller-owned, so it is gathered and scattered here
    const size_t nb = m_body_hs.size();
    std::vector<double> com(3 * nb, 0.0), velo(3 * nb, 0.0);
    for (size_t b = 0; b < nb; ++b)
        if (m_body_hs[b])
            for (int j = 0; j < 3; ++j) {
                com[3 * b + j] = m_body_hs[b]->center_of_mass[j];
                velo[3 * b + j] = m_body_hs[b]->center_of_mass_velo[j];
            }
    rc = clsn_update_rigid_bodies(m_ctx, com.data(), velo.data());
    if (rc != CLSN_OK) fail(rc, "clsn_update_rigid_bodies");
    for (size_t b = 0; b < nb; ++b)
        if (m_body_hs[b])
            for (int j = 0; j < 3; ++j) {
                m_body_hs[b]->center_of_mass[j] = com[3 * b + j];
                m_body_hs[b]->center_of_mass_velo[j] = velo[3 * b + j];
            }
}

// Single-pair entry points of the reference (collid.h:199-200).  They run the same kernels on a
// two-element problem and add the pair's contributions to the points' accumulators.
static bool single_pair(clsn_ctx*& ctx, int device, const CD_HSE* a, const CD_HSE* b, int mode, double eps, double thickness, double dt,
                        double k, double m, double lambda, double cr)
{
    const bool ta = a->num_pts() == 3, tb = b->num_pts() == 3;
    if (a->num_pts() == 1 || b->num_pts() == 1) throw std::runtime_error("This case has not been implemented");  // dcollid.cpp:787-791
    std::vector<POINT*> pts;
    auto pid = [&](POINT* p) {
        for (size_t i = 0; i < pts.size(); ++i)
            if (pts[i] == p) return (int)i;
        pts.push_back(p);
        return (int)pts.size() - 1;
    };
    std::vector<int32_t> tri, surf, bond;
    const CD_HSE* order[2] = {ta || !tb ? a : b, ta || !tb ? b : a};  // triangles first, like hseList
    for (const CD_HSE* h : order) {
        if (h->num_pts() == 3) {
            for (int i = 0; i < 3; ++i) tri.push_back(pid(h->Point_of_hse(i)));
            const CD_TRI* t = dynamic_cast<const CD_TRI*>(h);
            surf.push_back(surf.empty() ? 0 : (t->m_tri->surf == dynamic_cast<const CD_TRI*>(order[0])->m_tri->surf ? 0 : 1));
        } else {
            for (int i = 0; i < 2; ++i) bond.push_back(pid(h->Point_of_hse(i)));
        }
    }
    const int V = (int)pts.size();
    std::vector<uint8_t> flags(V);
    std::vector<int32_t> body(V);
    std::vector<double> mass(V), xo(3 * V), av(3 * V);
    for (int v = 0; v < V; ++v) {
        flags[v] = (uint8_t)((pts[v]->state->is_fixed ? 1 : 0) | (pts[v]->state->is_movableRG ? 2 : 0));
        body[v] = v;
        mass[v] = pts[v]->hs ? pts[v]->hs->total_mass : 0.0;
        for (int j = 0; j < 3; ++j) { xo[3 * v + j] = pts[v]->state->x_old[j]; av[3 * v + j] = pts[v]->state->avgVel[j]; }
    }
    // one small context per solver, created on first use: clsn_set_topology resets everything a pair evaluation reads
    // (body accumulators, has_collsn, the tree), so consecutive calls are independent of each other
    if (!ctx && clsn_create(&ctx, device) != CLSN_OK) {
        ctx = nullptr;
        throw std::runtime_error("collision_b200: no usable CUDA device");
    }
    clsn_ctx* c = ctx;
    clsn_params prm;
    prm.eps = eps; prm.thickness = thickness; prm.dt = dt; prm.k = k; prm.m = m; prm.lambda = lambda; prm.cr = cr;
    for (int i = 0; i < 3; ++i) { prm.lo[i] = -1e30; prm.hi[i] = 1e30; }
    clsn_pass_stats st;
    std::memset(&st, 0, sizeof(st));
    std::vector<double> imp(3 * V), fric(3 * V), irg(3 * V);
    std::vector<int32_t> cnt(V), crg(V);
    int rc = clsn_set_params(c, &prm);
    if (!rc) rc = clsn_set_topology(c, V, (int)surf.size(), tri.data(), surf.data(), (int)bond.size() / 2, bond.data(), flags.data(),
                                    body.data(), V, mass.data());
    if (!rc) rc = clsn_upload_state(c, xo.data(), xo.data());
    if (!rc) rc = clsn_set_avgvel(c, av.data());
    if (!rc) rc = clsn_detect(c, mode, &st);
    if (!rc) rc = clsn_get_accumulators(c, imp.data(), fric.data(), cnt.data(), irg.data(), crg.data());
    if (rc) {
        const std::string why = clsn_last_error(c);
        clsn_destroy(c);   // do not keep a context that failed half-way
        ctx = nullptr;
        throw std::runtime_error("collision_b200: single-pair evaluation failed: " + why);
    }
    for (int v = 0; v < V; ++v) {
        STATE* sl = pts[v]->state;
        for (int j = 0; j < 3; ++j) {
            sl->collsnImpulse[j] += imp[3 * v + j];
            sl->friction[j] += fric[3 * v + j];
            sl->collsnImpulse_RG[j] += irg[3 * v + j];
        }
        sl->collsn_num += cnt[v];
        sl->collsn_num_RG += crg[v];
    }
    return st.true_pairs > 0;
}

bool CollisionSolver::isProximity(const CD_HSE* a, const CD_HSE* b)
{
    return single_pair(m_pair_ctx, m_device, a, b, CLSN_PROXIMITY, s_eps, s_thickness, s_dt, s_k, s_m, s_lambda, s_cr);
}
bool CollisionSolver::isCollision(const CD_HSE* a, const CD_HSE* b)
{
    return single_pair(m_pair_ctx, m_device, a, b, CLSN_COLLISION, s_eps, s_thickness, s_dt, s_k, s_m, s_lambda, s_cr);
}

// ---- adapters (dcollid.cpp:852-939) -------------------------------------------------------------
static double ext_static(POINT* const* p, int n, int dim, bool mx, bool use_coords)
{
    double ans = mx ? -1e18 : 1e18;
    for (int i = 0; i < n; ++i) {
        const double x = use_coords ? p[i]->coords[dim] : p[i]->state->x_old[dim];
        ans = mx ? std::max(ans, x) : std::min(ans, x);
    }
    return ans;
}
static double ext_moving(POINT* const* p, int n, int dim, double dt, bool mx)
{
    double ans = mx ? -1e18 : 1e18;
    for (int i = 0; i < n; ++i) {
        const STATE* sl = p[i]->state;
        const double x0 = sl->x_old[dim], x1 = sl->x_old[dim] + sl->avgVel[dim] * dt;
        ans = mx ? std::max(std::max(ans, x0), x1) : std::min(std::min(ans, x0), x1);
    }
    return ans;
}
double CD_TRI::max_static_coord(int d) { return ext_static(m_tri->pts, 3, d, true, false); }
double CD_TRI::min_static_coord(int d) { return ext_static(m_tri->pts, 3, d, false, false); }
double CD_TRI::max_moving_coord(int d, double dt) { return ext_moving(m_tri->pts, 3, d, dt, true); }
double CD_TRI::min_moving_coord(int d, double dt) { return ext_moving(m_tri->pts, 3, d, dt, false); }
POINT* CD_TRI::Point_of_hse(int i) const { return i >= 3 ? nullptr : m_tri->pts[i]; }
double CD_BOND::max_static_coord(int d) { POINT* p[2] = {m_bond->start, m_bond->end}; return ext_static(p, 2, d, true, true); }
double CD_BOND::min_static_coord(int d) { POINT* p[2] = {m_bond->start, m_bond->end}; return ext_static(p, 2, d, false, true); }
double CD_BOND::max_moving_coord(int d, double dt) { POINT* p[2] = {m_bond->start, m_bond->end}; return ext_moving(p, 2, d, dt, true); }
double CD_BOND::min_moving_coord(int d, double dt) { POINT* p[2] = {m_bond->start, m_bond->end}; return ext_moving(p, 2, d, dt, false); }
POINT* CD_BOND::Point_of_hse(int i) const { return i >= 2 ? nullptr : (i == 0 ? m_bond->start : m_bond->end); }
double CD_POINT::max_static_coord(int d) { return m_point->state->x_old[d]; }
double CD_POINT::min_static_coord(int d) { return m_point->state->x_old[d]; }
double CD_POINT::max_moving_coord(int d, double dt) { POINT* p[1] = {m_point}; return ext_moving(p, 1, d, dt, true); }
double CD_POINT::min_moving_coord(int d, double dt) { POINT* p[1] = {m_point}; return ext_moving(p, 1, d, dt, false); }
POINT* CD_POINT::Point_of_hse(int i) const { return i >= 1 ? nullptr : m_point; }

}  // namespace clsn_host
