/* collision_b200 -- C ABI of the B200 collision step.
 *
 * Drop-in boundary for the per-step hot path of antdvid/Collision: everything that
 * CollisionSolver::resolveCollision() (dcollid.cpp:317-362) does between "the caller filled
 * x_old / candidate Coords" and "avgVel / Coords / vel / has_collsn are final", i.e.
 *   computeAverageVelocity   dcollid.cpp:160-220
 *   detectProximity          dcollid.cpp:390-406  (aabbProximity :366-388, AABB.cpp, TriToTri ...)
 *   detectCollision          dcollid.cpp:430-468  (aabbCollision :409-428, MovingTriToTri ...)
 *   updateAverageVelocity    dcollid.cpp:677-751  (+ updateImpactZoneVelocityForRG :267-288)
 *   detectDomainBoundaryCollision :116-158, updateFinalPosition :562-584, updateFinalVelocity :598-624
 * runs on the GPU behind these entry points.  The host-side mirror of the reference's own
 * interface (CollisionSolver3d, CD_HSE/CD_TRI/CD_BOND/CD_POINT) that calls them lives in
 * collision_b200/host/collid_b200.h; INTEGRATION.md shows the binding a maintainer adds.
 *
 * Plain C types only.  All host arrays are caller-owned; "3V" arrays are xyz per vertex.
 * Every function returns 0 on success or a negative clsn_status; clsn_last_error() gives text.
 * A context is bound to one CUDA device and is not thread-safe.  There is no CPU fallback:
 * without a usable device clsn_create fails.
 */
#ifndef COLLISION_B200_H
#define COLLISION_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct clsn_ctx clsn_ctx;

typedef enum {
    CLSN_OK = 0,
    CLSN_E_CUDA = -1,      /* CUDA runtime error                                        */
    CLSN_E_ARG = -2,       /* bad argument / call order                                 */
    CLSN_E_NUMERIC = -3,   /* NaN/Inf where the reference calls clean_up(ERROR)          */
    CLSN_E_NOMEM = -4,     /* device allocation failed                                  */
    CLSN_E_UNSUPPORTED = -5/* element pair type the reference does not implement either  */
} clsn_status;

/* CollisionSolver's static parameters (dcollid.cpp:28-35, setters :57-89) + domain box
 * (setDomainBoundary :109-114).  dt is what assembleFromInterface passes to setTimeStepSize. */
typedef struct {
    double eps;        /* setRoundingTolerance   default 1e-6  */
    double thickness;  /* setFabricThickness     default 1e-4  */
    double dt;         /* setTimeStepSize        default 1e-3  */
    double k;          /* setSpringConstant      default 1000  */
    double m;          /* setPointMass           default 0.01  */
    double lambda;     /* setFrictionConstant    default 0.02  */
    double cr;         /* setRestitutionCoef     default 0     */
    double lo[3];      /* domain lower corner                  */
    double hi[3];      /* domain upper corner                  */
} clsn_params;

enum { CLSN_PROXIMITY = 0, CLSN_COLLISION = 1 };
enum { CLSN_VFLAG_FIXED = 1, CLSN_VFLAG_MOVABLE_RG = 2 };
#define CLSN_MAX_CCD_PASSES 5 /* MAX_ITER, dcollid.cpp:433 */

/* One detection pass (aabbProximity/aabbCollision + tree query). */
typedef struct {
    int64_t candidates;     /* element pairs with overlapping leaf boxes = narrow-phase callbacks  */
    int64_t pairs_tested;   /* of those, pairs without a shared vertex handed to the feature tests */
    int64_t true_pairs;     /* pairs for which isProximity/isCollision returned true (tree count)  */
    int64_t contacts;       /* feature tests that fired                                            */
    int64_t contributions;  /* per-point impulse records reduced                                   */
    int64_t features;       /* feature tests evaluated (survived box cull and, CCD, the classifier) */
    int64_t box_survivors;  /* feature tests that survived the swept-box cull                       */
    int64_t coplanar;       /* CCD: features whose cubic has a usable root (isCoplanar true)        */
    int64_t exact_solves;   /* CCD: correctly rounded cubic solves (pipeline 1: only the features the
                               plain-FP64 fast path could not settle; pipeline 0: == coplanar)        */
} clsn_pass_stats;

typedef struct {
    clsn_pass_stats proximity;
    int32_t n_ccd_passes;
    int32_t has_collision;      /* hasCollision(): first CCD pass found something (dcollid.cpp:456) */
    int32_t still_colliding;    /* MAX_ITER passes did not resolve: the reference enters computeImpactZone
                                   (dcollid.cpp:464-467); run here when clsn_set_impact_zones is on   */
    int32_t zone_iterations;    /* iterations of that fail-safe (0: not entered)                    */
    clsn_pass_stats ccd[CLSN_MAX_CCD_PASSES];
    float ms_total;             /* device time of the whole step (CUDA events)                     */
    float ms_phase[10];         /* avgvel, build, refit, traverse, cull, roots, contact, reduce, finalize, other */
    int32_t zones;              /* sets of more than one point after the last fail-safe iteration  */
    int32_t strain_sweeps;      /* reduceSuperelastOnce sweeps run (0: strain limiting off)          */
    int32_t strain_edges;       /* edges averaged in the last of them (the reference's num_edges)   */
    int32_t reserved;
} clsn_step_stats;

/* computeImpactZone (dcollid.cpp:227-265). */
typedef struct {
    int32_t iterations;   /* CCD pass + updateAverageVelocity + updateImpactZoneVelocity rounds      */
    int32_t zones;        /* numZones of the last round: union-find sets with more than one point     */
    int32_t zone_points;  /* points in those sets                                                     */
    int32_t converged;    /* 1: the last round found no collision                                     */
    int64_t true_pairs;   /* colliding pairs summed over the rounds                                   */
    int64_t merges;       /* successful mergePoint calls (dcollid.cpp:1039-1059)                      */
} clsn_zone_stats;

/* Contact record for parity checks (same layout as oracle/collision_oracle.h: orc_contact). */
typedef struct {
    int32_t ea, eb;   /* element pair, ea < eb (hseList order: triangles then bonds) */
    int32_t feature;  /* index of the feature test inside the pair, reference loop order */
    int32_t kind;     /* 0 point-triangle, 1 edge-edge */
    int32_t p[4];     /* points as passed to PointToTri / EdgeToEdge */
    double root;      /* time of impact (CCD) or 0 */
    double dist;
    double nor[3];
    double w[3];      /* point-tri: w0..w2 ; edge-edge: a, b, 0 */
} clsn_contact;

/* ---- lifetime */
int clsn_create(clsn_ctx** out, int device);
void clsn_destroy(clsn_ctx*);
const char* clsn_last_error(const clsn_ctx*);

/* ---- setup (assembleFromInterface, dcollid3d.cpp:12-52): once per topology */
int clsn_set_params(clsn_ctx*, const clsn_params*);
/* tri_idx[3T], tri_surf[T] (surface id per triangle), bond_idx[2B], vflags[V] (CLSN_VFLAG_*),
 * vbody[V] hyper-surface/body id per vertex in [0,nbody), body_mass[nbody] = total_mass(hs). */
int clsn_set_topology(clsn_ctx*, int V, int T, const int32_t* tri_idx, const int32_t* tri_surf, int B,
                      const int32_t* bond_idx, const uint8_t* vflags, const int32_t* vbody, int nbody,
                      const double* body_mass);

/* ---- per step, host buffers (the call sequence a drop-in CollisionSolver3d makes) */
/* x_old[3V], x_new[3V] = Coords as the spring solver left them; clears per-step accumulators */
int clsn_upload_state(clsn_ctx*, const double* x_old, const double* x_new);
int clsn_resolve(clsn_ctx*, clsn_step_stats* stats); /* resolveCollision() minus strain limiting */
/* The fail-safe the reference enters when MAX_ITER passes leave collisions (detectCollision,
 * dcollid.cpp:464-467 -> computeImpactZone :227-265, createImpZone :473-484,
 * updateImpactZoneVelocity :290-309).  clsn_set_impact_zones(on) makes clsn_resolve / clsn_step_host
 * run it like the reference does (default off: it is outside the timed hot loop); max_iter <= 0
 * means the reference's unbounded loop (guarded at 100000 -> CLSN_E_NUMERIC).
 * clsn_compute_impact_zone is the loop alone, for a caller driving the phases itself.  The CCD passes
 * and the rigid projection of the zones run on the GPU; the union-find merges are replayed on the
 * host in canonical order from the contact records.  Whole-mesh contexts only (not clsn_set_slice). */
/* Strain limiting (reduceSuperelast, dcollid.cpp:485-596; called from resolveCollision :355 between the
 * final positions and the final velocities, so it only changes avgVel -> vel).  Rest lengths are the
 * application's data in the reference: tri_len0[3T] = TRI::side_length0 (edge j joins points j and
 * (j+1)%3), bond_len0[B] = BOND::length0; set them again after clsn_set_topology.
 * clsn_set_strain_limiting(on) makes clsn_resolve / clsn_step_host run it (default off: outside the
 * timed hot loop); clsn_strain_limit runs it alone on the resident avgVel.  The reference's sequential
 * sweep order is kept exactly through a wavefront schedule of the edge visits (csrc/strain.cuh). */
int clsn_set_rest_lengths(clsn_ctx*, const double* tri_len0, const double* bond_len0);
int clsn_set_strain_limiting(clsn_ctx*, int on);
int clsn_strain_limit(clsn_ctx*, int32_t* sweeps, int32_t* edges_last);
int clsn_set_impact_zones(clsn_ctx*, int on, int max_iter);
int clsn_compute_impact_zone(clsn_ctx*, int max_iter, clsn_zone_stats* out);
/* any pointer may be NULL.  x[3V] final Coords, avgvel[3V], has_collsn[V] */
int clsn_download_state(clsn_ctx*, double* x, double* avgvel, uint8_t* has_collsn);
/* upload + resolve + download + "vel = avgVel where has_collsn" (updateFinalVelocity) in one call */
int clsn_step_host(clsn_ctx*, const double* x_old, const double* x_new, double* x_out, double* vel_inout,
                   uint8_t* has_collsn_out, clsn_step_stats* stats);
/* the same single call (one synchronisation) with the raw per-point results instead of the merged velocity: avgvel_out[3V]
 * = STATE::avgVel of EVERY point, as the reference leaves it (dcollid.cpp:677-751); any output pointer may be NULL.  For
 * host mirrors that keep avgVel per point (collision_b200/host/collid_b200.cpp::resolveCollision). */
int clsn_step_host_state(clsn_ctx*, const double* x_old, const double* x_new, double* x_out, double* avgvel_out,
                         uint8_t* has_collsn_out, clsn_step_stats* stats);
/* page-locked host memory for callers that do not link the CUDA runtime themselves: host arrays passed to the calls above
 * from such a buffer move by asynchronous DMA (pageable arrays are staged by the driver).  Usable with every context of
 * the process; free with clsn_host_free. */
int clsn_host_alloc(void** out, size_t bytes);
void clsn_host_free(void*);

/* updateFinalForRG (dcollid.cpp:626-675, the tail of updateFinalVelocity): for every movable rigid body with a point
 * that collided in this step, write the body's centre-of-mass velocity (avgVel of the first such point in hseList order)
 * and centre of mass (that velocity * dt + the centre recorded at the end of the previous call, CollisionSolver::mrg_com,
 * kept inside the context) into the caller's HYPER_SURF data.  Arrays are [3 * nbody], indexed like body_mass; entries of
 * bodies that were not hit are left alone.  Call after clsn_resolve / clsn_step_host. */
int clsn_update_rigid_bodies(clsn_ctx*, double* center_of_mass, double* center_of_mass_velo);

/* ---- per step, device buffers (inputs already resident in HBM; 3V doubles, xyz per vertex) */
int clsn_upload_state_device(clsn_ctx*, const double* d_x_old, const double* d_x_new);
int clsn_download_state_device(clsn_ctx*, double* d_x, double* d_avgvel);

/* ---- single phases (parity tests drive these from identical inputs) */
int clsn_avg_velocity(clsn_ctx*);
int clsn_detect(clsn_ctx*, int mode, clsn_pass_stats* stats); /* tree + query + narrow phase    */
int clsn_apply(clsn_ctx*, int rigidify);                      /* reduce + updateAverageVelocity */
int clsn_boundary(clsn_ctx*);
int clsn_final_position(clsn_ctx*);
int clsn_set_avgvel(clsn_ctx*, const double* avgvel);         /* host 3V */

/* ---- measurement: CUDA events on the library's stream, kernel-launch counter */
int clsn_timer_start(clsn_ctx*);
int clsn_timer_stop(clsn_ctx*, float* ms);
int64_t clsn_launch_count(clsn_ctx*, int reset);
int clsn_synchronize(clsn_ctx*);
/* clsn_step_stats.ms_phase: CUDA-event marks between the kernel groups of clsn_resolve / clsn_step_host (about a hundred
 * event records per step).  Off by default -- ms_phase stays zero, ms_total is always measured; the environment variable
 * CLSN_PHASE_TIMING=1 sets the default of new contexts. */
int clsn_set_phase_timing(clsn_ctx*, int on);

/* ---- multi-GPU inside the library (csrc/dist.cuh): one context per GPU of a node, one process or thread each (the
 * natural fit for the reference's MPI-based host application: one MPI rank per GPU).  Any rank calls clsn_dist_unique_id
 * once and hands the 128 bytes to all ranks (MPI_Bcast / torch.distributed / a file); after clsn_set_topology every rank
 * calls clsn_dist_init.  From then on clsn_resolve / clsn_step_host run the distributed step: every rank is given the
 * same x_old / x_new and returns the same, complete result -- bit-identical to one GPU.  Mesh and tree are replicated, rank
 * r traverses its slice of the Morton-ordered query leaves, impulse records are stored straight into the receive buffer of
 * the rank that owns the point (NVLink peer stores from inside the narrow-phase kernels, mapped with cudaIpc), each rank
 * reduces its own vertex range, NCCL all-reduces the pass counters (the device-side `while (is_collision)`) and
 * all-gathers avgVel / has_collsn.  No host read-back inside the step.  clsn_upload_state / clsn_step_host are collective
 * in this mode (each rank moves only its share of the input arrays across PCIe, the rest arrives over NVLink), like
 * clsn_resolve, clsn_detect and clsn_apply: every rank has to make the same calls.  Not available in this mode: the impact-zone
 * fail-safe (CLSN_E_UNSUPPORTED if a step still collides after MAX_ITER passes with clsn_set_impact_zones on). */
int clsn_dist_unique_id(void* id128);
int clsn_dist_init(clsn_ctx*, int rank, int nranks, const void* id128);
int clsn_dist_nranks(const clsn_ctx*);

/* ---- multi-GPU, external exchange (the caller moves the records, e.g. with its own NCCL communicator): this rank
 * traverses query leaves [rank*N/nranks, (rank+1)*N/nranks) of the
 * Morton order; contribution records are exchanged by the caller (NCCL all-gather of the buffers
 * below) and every rank reduces the union in canonical order -> identical state on all ranks. */
int clsn_set_slice(clsn_ctx*, int rank, int nranks);
/* after clsn_detect: device pointers + counts of the records this rank produced */
int clsn_export_records(clsn_ctx*, void** d_point_records, int64_t* n_point_records, void** d_body_records,
                        int64_t* n_body_records, int64_t* true_pairs);
/* replace the record set to be reduced by clsn_apply with externally gathered ones (device ptrs) */
int clsn_import_records(clsn_ctx*, const void* d_point_records, int64_t n_point_records,
                        const void* d_body_records, int64_t n_body_records);
/* owner-computes variant: vertex v belongs to rank v / ceil(V/nranks).  clsn_bucket_records groups this
 * rank's point records by owner (counts[r] records for rank r, contiguous in *d_sorted) for an
 * all-to-all; each rank imports what it received, runs clsn_apply_stage(1) (reduce -> avgVel of its own
 * vertices), the ranks all-gather their avgVel / has_collsn / touched slices (clsn_state_device_ptrs:
 * 32-byte Vec4 per vertex, 1 byte, 1 byte), then clsn_apply_stage(2) (body records, rigid bodies). */
int clsn_bucket_records(clsn_ctx*, int nranks, int64_t* counts, void** d_sorted);
/* run on the caller's CUDA stream (cudaStream_t passed as void*; NULL = the context's own stream; for the
 * legacy default stream pass cudaStreamLegacy, i.e. (void*)0x1) */
int clsn_set_stream(clsn_ctx*, void* cuda_stream);
int clsn_apply_stage(clsn_ctx*, int stage, int rigidify);
int clsn_state_device_ptrs(clsn_ctx*, void** d_avgvel_vec4, void** d_has_collsn, void** d_touched);
#define CLSN_POINT_RECORD_BYTES 64
#define CLSN_BODY_RECORD_BYTES 48

/* ---- debug / parity readbacks (host buffers) */
int clsn_set_debug(clsn_ctx*, int record_candidates, int record_contacts);
/* From the second CCD pass on, pairs none of whose points was changed by the previous pass repeat their
 * (hit-free) outcome and are skipped; queries without a changed point then only visit subtrees that hold
 * one, and clsn_pass_stats.candidates counts only the pairs found.  on != 0 restores the full traversal so
 * that `candidates` equals the reference's callback count in every pass (results are identical). */
int clsn_set_exact_stats(clsn_ctx*, int on);
/* CCD narrow phase of MovingPointToTri / MovingEdgeToEdge (dcollid3d.cpp:327-369) after the cull.
 * 1 (default): a plain-FP64 fast path (k_fast) first settles every feature it can prove to miss at all of its
 * roots -- the outcome is then the static test at t = dt -- and only the rest gets the correctly rounded cubic
 * solve (k_exact); k_emit writes the records of the hit list.  0: staged -- correctly rounded solve of every
 * feature (k_roots), then the static tests and the records (k_contact); kept for A/B measurements.
 * 2 (experimental; measured on a B200: no gain over 1 -- the reduction gets faster by what the emission gets slower):
 * like 1, but the per-point record counts are taken from the hit list first, so that k_emit writes every impulse record
 * straight into its point's segment and the reduction needs no grouping pass (single-GPU contexts only; ranks of a
 * multi-GPU run keep exchanging the plain record list).
 * Results are bit-identical.  The environment variable CLSN_PIPELINE=0|1|2 sets the default of new contexts. */
int clsn_set_pipeline(clsn_ctx*, int pipeline);
int64_t clsn_num_candidates(clsn_ctx*);
int clsn_get_candidates(clsn_ctx*, int32_t* pairs /* 2 per pair, unsorted */);
int64_t clsn_num_contacts(clsn_ctx*);
int clsn_get_contacts(clsn_ctx*, clsn_contact* out /* unsorted */);
/* accumulators as updateAverageVelocity would see them, reduced in canonical order:
 * imp[3V], fric[3V], cnt[V], imp_rg[3*nbody], cnt_rg[nbody] */
int clsn_get_accumulators(clsn_ctx*, double* imp, double* fric, int32_t* cnt, double* imp_rg, int32_t* cnt_rg);
int clsn_set_body_accumulators(clsn_ctx*, const double* imp_rg, const int32_t* cnt_rg);

#ifdef __cplusplus
}
#endif
#endif
