/* CPU restatement of the reference's per-step collision pipeline -- TEST INFRASTRUCTURE.
 *
 * Plain C over flat arrays; every function cites the reference file:line it restates.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may build, load or call
 * this.  The product (collision_b200/) never does.
 *
 * Parity pinning: the reference ships no golden vectors, so this restatement is
 * pinned against the reference's own sources compiled unmodified (oracle/_ref, built from
 * /root/reference by oracle/Makefile) -- feature-level known answers and whole scenes, with the
 * outputs committed under tests/golden/ (tests/golden/make_golden.py).
 *
 * Canonical order.  The reference evaluates overlapping element pairs in the traversal order
 * of its insertion-built AABB tree (AABB.cpp:254-343); results depend on that order only
 * through floating-point summation order of the per-point accumulators.
 * This restatement evaluates pairs sorted by (a, b) with a < b (hseList indices), pair (a, b)
 * meaning isProximity/isCollision(hse[a], hse[b]), features in the reference's loop order.
 * The CUDA path reduces in exactly this order, so CUDA == oracle bit for bit, and
 * oracle == reference up to summation order (<= 1e-12 relative).
 */
#ifndef CLSN_COLLISION_ORACLE_H
#define CLSN_COLLISION_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_ctx orc_ctx;

typedef struct {
    int ea, eb;     /* element pair, ea < eb */
    int feature;    /* index of the feature test inside the pair, reference loop order */
    int kind;       /* 0 point-triangle, 1 edge-edge */
    int p[4];       /* point ids as passed to PointToTri / EdgeToEdge */
    double root;    /* time of impact (CCD) or 0 (proximity) */
    double dist;    /* distance at the test position */
    double nor[3];  /* unit normal handed to the impulse routine */
    double w[3];    /* point-tri: barycentric w0..w2 ; edge-edge: a, b, 0 */
} orc_contact;

enum { ORC_PROXIMITY = 0, ORC_COLLISION = 1 };
enum { ORC_LIBM_NATIVE = 0, ORC_LIBM_CR = 1 };
/* process-wide choice of acos/cos/pow used by isCoplanar: the host libm (as the reference) or
 * correctly rounded (binary128, rounded once).  See collision_oracle.c. */
void orc_set_libm(int mode);
int orc_get_libm(void);

orc_ctx* orc_create(int V, int T, const int* tri_idx, const int* tri_surf, int B, const int* bond_idx,
                    const unsigned char* vflags, const int* vhs, int nhs, const double* hs_mass);
void orc_destroy(orc_ctx*);
void orc_set_params(orc_ctx*, double eps, double thickness, double k, double m, double lambda, double cr);
void orc_set_domain(orc_ctx*, const double* lo, const double* hi);
void orc_set_dt(orc_ctx*, double dt);
/* start-of-step state: x_old, candidate positions; clears imp/fric/cnt (test.cpp:192-224) */
void orc_set_state(orc_ctx*, const double* x_old, const double* x_new);
void orc_set_avgvel(orc_ctx*, const double* avgvel);

void orc_avg_velocity(orc_ctx*);                 /* dcollid.cpp:160-220 */
long orc_detect(orc_ctx*, int mode);             /* aabbProximity / aabbCollision: returns # true pairs */
/* narrow phase over explicit ORDERED pairs (a,b) in the given order (replays the reference's callbacks) */
long orc_detect_ordered(orc_ctx*, int mode, const int* pairs, long n);
void orc_apply(orc_ctx*, int rigidify);          /* updateAverageVelocity, dcollid.cpp:677-751 */
void orc_boundary(orc_ctx*);                     /* dcollid.cpp:116-158 */
void orc_final_position(orc_ctx*);               /* dcollid.cpp:562-584 */
void orc_final_velocity(orc_ctx*, double* vel);  /* dcollid.cpp:598-624, vel updated in place */
/* whole step (dcollid.cpp:317-362 minus reduceSuperelast and the impact-zone fail-safe).
 * stats[0] proximity pairs true, stats[1] #CCD passes, stats[2..6] true pairs per CCD pass,
 * stats[7] 1 if still colliding after MAX_ITER passes, stats[8] candidates proximity,
 * stats[9..13] candidates per CCD pass, stats[14] impact-zone iterations, stats[15] zones,
 * stats[16] strain-limiting sweeps, stats[17] edges averaged in the last sweep (20 longs) */
void orc_resolve(orc_ctx*, double* vel, long* stats);

/* Impact-zone fail-safe (computeImpactZone, dcollid.cpp:227-265; createImpZone :473-484;
 * updateImpactZoneVelocity :290-309).  orc_enable_impact_zones makes orc_resolve enter it when
 * MAX_ITER passes leave collisions; the pieces are exposed for phase-by-phase pinning. */
void orc_enable_impact_zones(orc_ctx*, int on);
void orc_set_imp_zone(orc_ctx*, int on);            /* turnOnImpZone / turnOffImpZone */
int orc_zone_velocity(orc_ctx*);                    /* updateImpactZoneVelocity -> number of zones */
void orc_impact_zone(orc_ctx*, int max_iter, long* out3);

/* Strain limiting (reduceSuperelast, dcollid.cpp:485-596).  Rest lengths are the caller's data in the
 * reference (TRI::side_length0[3], BOND::length0).  orc_enable_strain_limiting makes orc_resolve run it
 * between the final positions and the final velocities like resolveCollision (:355). */
void orc_set_rest_lengths(orc_ctx*, const double* tri_len0, const double* bond_len0);
void orc_enable_strain_limiting(orc_ctx*, int on);
int orc_strain_limit_once(orc_ctx*, long* num_edges);
int orc_strain_limit(orc_ctx*, long* num_edges);

/* readbacks */
void orc_get_f64(orc_ctx*, int field, double* out); /* 0 x_old 1 x 2 avgVel 3 imp 4 fric (3V each) */
/* updateFinalForRG (dcollid.cpp:626-675): com / com_velo [3*nhs] in/out, mrg_com kept inside the context */
void orc_update_final_for_rg(orc_ctx*, double* com, double* com_velo);
void orc_set_has_collsn(orc_ctx*, const unsigned char* has);   /* V flags (tests: state taken from the reference) */
void orc_get_i32(orc_ctx*, int field, int* out);    /* 0 cnt 1 has_collsn (V each) */
void orc_get_body(orc_ctx*, double* imp_rg /*3*nhs*/, int* cnt_rg /*nhs*/);
void orc_set_body(orc_ctx*, const double* imp_rg, const int* cnt_rg);
long orc_num_candidates(orc_ctx*);
void orc_get_candidates(orc_ctx*, int* out /* 2 per pair, sorted */);
long orc_num_contacts(orc_ctx*);
void orc_get_contacts(orc_ctx*, orc_contact* out);
long orc_num_true_pairs(orc_ctx*);
void orc_get_true_pairs(orc_ctx*, int* out /* 2 per pair, sorted */);

/* single feature tests on 4 free points (known-answer tests).
 * kind 0 isCoplanar, 1 PointToTri, 2 EdgeToEdge, 3 MovingPointToTri, 4 MovingEdgeToEdge.
 * acc[4][10]: imp[3], fric[3], impRG[3], cnt + 1000*cntRG per point. */
int orc_feature(int kind, const double* x_old, const double* coords, const double* avgvel,
                const unsigned char* flags, const double* mass, double h, double dt, const double* params,
                double* roots, double* acc, double* hit_root);

#ifdef __cplusplus
}
#endif
#endif
