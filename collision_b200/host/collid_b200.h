// collid_b200.h -- C++ host mirror of the reference's collid.h for the GPU collision step.
//
// Same class and adapter names, same public methods and call sequence as antdvid/Collision
// (collid.h:43-92 CD_HSE/CD_TRI/CD_BOND/CD_POINT, :128-244 CollisionSolver/CollisionSolver3d):
//
//     CollisionSolver3d* solver = new CollisionSolver3d();
//     ... each step:
//     solver->assembleFromInterface(intfc, dt);      // dcollid3d.cpp:12-52
//     solver->setFrictionConstant(0.0);              // static setters, dcollid.cpp:57-89
//     solver->resolveCollision();                    // dcollid.cpp:317-362
//
// The mesh model below (POINT / TRI / BOND / SURFACE / CURVE / INTERFACE / STATE) carries exactly the
// fields the reference's three hot-path sources touch in FronTier's structs; inside a FronTier tree
// these definitions are replaced by <FronTier.h> + ifluid_state.h.  All mesh data stays
// caller-owned and is mutated in place like the reference does (Coords, STATE::avgVel/vel/has_collsn,
// POINT::vel).  The work itself is done by libcollision_b200.so through include/collision_b200.h;
// there is no CPU implementation behind this class.
#ifndef COLLID_B200_H_
#define COLLID_B200_H_

#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "collision_b200.h"

namespace clsn_host {

// ---- minimal FronTier-shaped mesh model ------------------------------------------------------
struct POINT;
struct UF {  // impact-zone union-find links (collid.h:22-27); the GPU path keeps its own lists
    POINT* next_pt;
    POINT* root;
    POINT* tail;
    int num_pts;
};
struct STATE {  // collid.h:29-39 + the rigid-body fields (dcollid.cpp:726-733, cdinit.cpp:192-204)
    double vel[3];
    double collsnImpulse[3];
    double collsnImpulse_RG[3];
    double friction[3];
    double avgVel[3];
    double x_old[3];
    int collsn_num;
    int collsn_num_RG;
    bool has_collsn;
    bool is_fixed;
    bool is_movableRG;
    UF impZone;
};
struct HYPER_SURF {
    int wave_type;
    int body_index;
    double total_mass;
    double center_of_mass[3];
    double center_of_mass_velo[3];
};
struct POINT {
    double coords[3];
    long global_index;
    STATE* state;  // left_state(p)
    HYPER_SURF* hs;
    double vel[3];
};
struct SURFACE;
struct TRI {
    POINT* pts[3];
    TRI* next;
    SURFACE* surf;
    double side_length0[3];
};
struct SURFACE {
    HYPER_SURF* hs;
    TRI* first_tri;
    bool is_bdry;
};
struct BOND {
    POINT* start;
    POINT* end;
    BOND* next;
    double length0;
};
struct CURVE {
    BOND* first;
    bool is_string;  // hsbdry_type(c) == STRING_HSBDRY
    HYPER_SURF* hs;
};
struct INTERFACE {
    std::vector<SURFACE*> surfaces;
    std::vector<CURVE*> curves;
    double L[3], U[3];  // table->rect_grid.L / U
};

// ---- element adapters (collid.h:43-92) --------------------------------------------------------
class CD_HSE {
public:
    virtual double max_static_coord(int) = 0;
    virtual double min_static_coord(int) = 0;
    virtual double max_moving_coord(int, double) = 0;
    virtual double min_moving_coord(int, double) = 0;
    virtual POINT* Point_of_hse(int) const = 0;
    virtual int num_pts() const = 0;
    virtual ~CD_HSE() {}
};
class CD_TRI : public CD_HSE {
public:
    TRI* m_tri;
    explicit CD_TRI(TRI* tri) : m_tri(tri) {}
    double max_static_coord(int);
    double min_static_coord(int);
    double max_moving_coord(int, double);
    double min_moving_coord(int, double);
    POINT* Point_of_hse(int) const;
    int num_pts() const { return 3; }
};
class CD_BOND : public CD_HSE {
public:
    BOND* m_bond;
    int m_dim;
    CD_BOND(BOND* bond, int dim) : m_bond(bond), m_dim(dim) {}
    double max_static_coord(int);
    double min_static_coord(int);
    double max_moving_coord(int, double);
    double min_moving_coord(int, double);
    POINT* Point_of_hse(int) const;
    int num_pts() const { return 2; }
};
// Declared but never defined in the reference (collid.h:82-92); given trivial bodies here.  Pairs that
// involve a CD_POINT stay unsupported, as in the reference (dcollid.cpp:787-791).
class CD_POINT : public CD_HSE {
public:
    POINT* m_point;
    explicit CD_POINT(POINT* point) : m_point(point) {}
    double max_static_coord(int);
    double min_static_coord(int);
    double max_moving_coord(int, double);
    double min_moving_coord(int, double);
    POINT* Point_of_hse(int) const;
    int num_pts() const { return 1; }
};

// Page-locked staging array (clsn_host_alloc): the per-vertex arrays that cross PCIe every step.  Pinned memory makes
// the library's cudaMemcpyAsync a true DMA and lets upload, step and download share one synchronisation.
template <class T>
class PinnedArray {
    T* p_ = nullptr;
    size_t n_ = 0;

public:
    PinnedArray() = default;
    PinnedArray(const PinnedArray&) = delete;
    PinnedArray& operator=(const PinnedArray&) = delete;
    ~PinnedArray() { clsn_host_free(p_); }
    bool resize(size_t n)   // contents are not kept
    {
        if (n == n_) return true;
        clsn_host_free(p_);
        p_ = nullptr;
        n_ = 0;
        void* q = nullptr;
        if (n && clsn_host_alloc(&q, n * sizeof(T)) != CLSN_OK) return false;
        p_ = static_cast<T*>(q);
        n_ = n;
        return true;
    }
    T* data() { return p_; }
    size_t size() const { return n_; }
    T& operator[](size_t i) { return p_[i]; }
};

// ---- solver (collid.h:128-244) ----------------------------------------------------------------
class CollisionSolver {
private:
    static double s_eps, s_thickness, s_dt, s_m, s_k, s_lambda, s_cr;  // process-global, as in the reference
    bool has_collision;
    double Boundary[3][2];

protected:
    int m_dim;
    std::vector<CD_HSE*> hseList;
    clsn_ctx* m_ctx;
    // gathered view of the mesh (rebuilt when the element list changes)
    std::vector<POINT*> m_points;
    std::unordered_map<POINT*, int> m_point_id;
    std::vector<int32_t> m_tri, m_tri_surf, m_bond;
    std::vector<uint8_t> m_flags;
    std::vector<int32_t> m_body;
    std::vector<double> m_body_mass;
    std::vector<HYPER_SURF*> m_body_hs;
    int m_dist_rank = -1, m_dist_nranks = 0;   // enableMultiGPU asked for before the topology was known
    unsigned char m_dist_id[128];   // body index -> the caller's HYPER_SURF (center_of_mass / center_of_mass_velo)
    PinnedArray<double> m_xold, m_xnew, m_xout, m_avg;   // x_old, candidate Coords in; final Coords, avgVel out
    PinnedArray<uint8_t> m_has;
    int m_device;
    clsn_ctx* m_pair_ctx;   // small context behind isProximity / isCollision, created on first use and kept
    clsn_step_stats m_stats;
    clsn_zone_stats m_zone_stats;
    std::vector<double> m_tri_len0, m_bond_len0;
    bool m_topology_dirty;
    void clearHseList();
    void gatherTopology(const INTERFACE*);
    void fail(int rc, const char* where) const;

public:
    explicit CollisionSolver(int dim, int device = 0);
    virtual ~CollisionSolver();
    static void setRoundingTolerance(double);
    static double getRoundingTolerance();
    static void setFabricThickness(double);
    static double getFabricThickness();
    static void setTimeStepSize(double);
    static double getTimeStepSize();
    static void setSpringConstant(double);
    static double getSpringConstant();
    static void setFrictionConstant(double);
    static double getFrictionConstant();
    static void setPointMass(double);
    static double getPointMass();
    static void setRestitutionCoef(double);
    static double getRestitutionCoef();
    static bool getImpZoneStatus() { return false; }  // s_detImpZone only lives inside the library's fail-safe loop
    // computeImpactZone (dcollid.cpp:227-265): entered by resolveCollision when MAX_ITER passes leave
    // collisions, like the reference's detectCollision (:464-467).  On by default; max_iter <= 0 = unbounded.
    void setImpactZones(bool on, int max_iter = 0);
    void computeImpactZone();
    // reduceSuperelast (dcollid.cpp:586-596) runs inside resolveCollision like the reference's (:355); on by default.
    void setStrainLimiting(bool on);

    virtual void assembleFromInterface(const INTERFACE*, double dt) = 0;
    virtual void createImpZoneForRG(const INTERFACE*) = 0;
    // Single-pair entry points (collid.h:199-200): the pair's own contributions are added to the points' accumulators.
    // Limitation: a rigid-rigid contact's body impulse (SpreadImpactZoneImpulse, dcollid.cpp:1101-1115) reaches only the
    // points of the two elements, not the rest of their bodies -- the whole-step path (resolveCollision) handles bodies.
    bool isProximity(const CD_HSE*, const CD_HSE*);
    bool isCollision(const CD_HSE*, const CD_HSE*);
    void resolveCollision();
    void recordOriginPosition();
    void setDomainBoundary(double* L, double* U);
    double getDomainBoundary(int dir, int side) { return Boundary[dir][side]; }
    bool hasCollision() { return has_collision; }
    const clsn_step_stats& lastStats() const { return m_stats; }
    bool stillColliding() const { return m_stats.still_colliding != 0; }  // MAX_ITER passes were not enough
    const clsn_zone_stats& lastZoneStats() const { return m_zone_stats; }
    // the reference's debug counters (collid.h:208-213: is_coplanar, edg_to_edg, pt_to_tri, printed by
    // printDebugVariable, dcollid.cpp:838-847) restated from the statistics of the last step
    void printDebugVariable() const;
    // Multi-GPU (include/collision_b200.h: clsn_dist_*): one solver per GPU of a node, one process (MPI rank) or thread
    // each.  Any rank obtains the 128-byte id once and hands it to all of them (MPI_Bcast ...); every rank then calls
    // enableMultiGPU -- before or after assembleFromInterface -- and keeps calling resolveCollision as before: all ranks
    // pass the same mesh state and get the same, complete result back (bit-identical to one GPU).
    static void multiGPUUniqueId(unsigned char id[128]);
    void enableMultiGPU(int rank, int nranks, const unsigned char id[128]);
};

class CollisionSolver3d : public CollisionSolver {
public:
    explicit CollisionSolver3d(int device = 0) : CollisionSolver(3, device) {}
    void assembleFromInterface(const INTERFACE*, double dt);
    void createImpZoneForRG(const INTERFACE*) {}  // the rigid-body lists are rebuilt from topology by clsn_set_topology
};

bool isStaticRigidBody(const POINT*);
bool isMovableRigidBody(const POINT*);
bool isRigidBody(const POINT*);

}  // namespace clsn_host
#endif
