"""Host-side mirror of the reference's solver interface over the C ABI (include/collision_b200.h).

`CollisionSolver3d` keeps the reference's public names and call sequence (collid.h:128-244,
test.cpp:105-107):

    solver = CollisionSolver3d()
    solver.assembleFromInterface(scene, dt)      # dcollid3d.cpp:12-52
    solver.setFrictionConstant(0.0)              # static parameter setters, dcollid.cpp:57-89
    solver.resolveCollision()                    # dcollid.cpp:317-362

The "interface" is a `collision_b200.scenes.Scene`-shaped object (flat arrays) instead of a FronTier
INTERFACE*; positions/velocities are the caller's numpy arrays, mutated in place like the
reference mutates Coords(p) / vel.  As in the reference the tolerances are process-global statics.

There is NO CPU path: the CUDA library must load and a B200 must be present, otherwise every
entry point raises.  Nothing here imports oracle/.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

_lib = None

PROXIMITY, COLLISION = 0, 1
MAX_CCD_PASSES = 5


class clsn_params(C.Structure):
    _fields_ = [("eps", C.c_double), ("thickness", C.c_double), ("dt", C.c_double), ("k", C.c_double),
                ("m", C.c_double), ("lambda_", C.c_double), ("cr", C.c_double), ("lo", C.c_double * 3),
                ("hi", C.c_double * 3)]


class clsn_pass_stats(C.Structure):
    _fields_ = [("candidates", C.c_int64), ("pairs_tested", C.c_int64), ("true_pairs", C.c_int64),
                ("contacts", C.c_int64), ("contributions", C.c_int64), ("features", C.c_int64),
                ("box_survivors", C.c_int64), ("coplanar", C.c_int64), ("exact_solves", C.c_int64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class clsn_step_stats(C.Structure):
    _fields_ = [("proximity", clsn_pass_stats), ("n_ccd_passes", C.c_int32), ("has_collision", C.c_int32),
                ("still_colliding", C.c_int32), ("zone_iterations", C.c_int32), ("ccd", clsn_pass_stats * MAX_CCD_PASSES),
                ("ms_total", C.c_float), ("ms_phase", C.c_float * 10), ("zones", C.c_int32),
                ("strain_sweeps", C.c_int32), ("strain_edges", C.c_int32), ("reserved", C.c_int32)]

    def as_dict(self):
        return dict(proximity=self.proximity.as_dict(), n_ccd_passes=int(self.n_ccd_passes),
                    has_collision=bool(self.has_collision), still_colliding=bool(self.still_colliding),
                    zone_iterations=int(self.zone_iterations), zones=int(self.zones),
                    strain_sweeps=int(self.strain_sweeps), strain_edges=int(self.strain_edges),
                    ccd=[self.ccd[i].as_dict() for i in range(int(self.n_ccd_passes))], ms_total=float(self.ms_total),
                    ms_phase=dict(zip(("avgvel", "build", "refit", "traverse", "cull", "roots", "contact", "reduce", "finalize", "other"),
                                      [float(v) for v in self.ms_phase])))


class clsn_zone_stats(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("zones", C.c_int32), ("zone_points", C.c_int32), ("converged", C.c_int32),
                ("true_pairs", C.c_int64), ("merges", C.c_int64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


CONTACT_DTYPE = np.dtype([("ea", "<i4"), ("eb", "<i4"), ("feature", "<i4"), ("kind", "<i4"), ("p", "<i4", (4,)),
                          ("root", "<f8"), ("dist", "<f8"), ("nor", "<f8", (3,)), ("w", "<f8", (3,))])


class CollisionError(RuntimeError):
    pass


def load_library():
    """Load (building if the sources are newer) collision_b200/libcollision_b200.so.  Raises if that
    is impossible -- there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("CLSN_LIB")  # tuning experiments only: an alternative build of the same sources
    if not path:
        path = _build.LIB_PATH
        if _build.needs_build():
            path = _build.build()
    L = C.CDLL(path)
    P, D, I, V = C.POINTER, C.c_double, C.c_int, C.c_void_p
    L.clsn_create.argtypes = [P(V), I]
    L.clsn_destroy.argtypes = [V]
    L.clsn_last_error.restype = C.c_char_p
    L.clsn_last_error.argtypes = [V]
    L.clsn_set_params.argtypes = [V, P(clsn_params)]
    L.clsn_set_topology.argtypes = [V, I, I, P(C.c_int32), P(C.c_int32), I, P(C.c_int32), P(C.c_uint8), P(C.c_int32),
                                    I, P(D)]
    L.clsn_upload_state.argtypes = [V, P(D), P(D)]
    L.clsn_resolve.argtypes = [V, P(clsn_step_stats)]
    L.clsn_download_state.argtypes = [V, P(D), P(D), P(C.c_uint8)]
    L.clsn_step_host.argtypes = [V, P(D), P(D), P(D), P(D), P(C.c_uint8), P(clsn_step_stats)]
    L.clsn_step_host_state.argtypes = [V, P(D), P(D), P(D), P(D), P(C.c_uint8), P(clsn_step_stats)]
    L.clsn_host_alloc.argtypes = [P(V), C.c_size_t]
    L.clsn_host_free.restype = None
    L.clsn_host_free.argtypes = [V]
    L.clsn_upload_state_device.argtypes = [V, V, V]
    L.clsn_download_state_device.argtypes = [V, V, V]
    L.clsn_avg_velocity.argtypes = [V]
    L.clsn_detect.argtypes = [V, I, P(clsn_pass_stats)]
    L.clsn_apply.argtypes = [V, I]
    L.clsn_boundary.argtypes = [V]
    L.clsn_final_position.argtypes = [V]
    L.clsn_set_avgvel.argtypes = [V, P(D)]
    L.clsn_set_slice.argtypes = [V, I, I]
    L.clsn_export_records.argtypes = [V, P(V), P(C.c_int64), P(V), P(C.c_int64), P(C.c_int64)]
    L.clsn_import_records.argtypes = [V, V, C.c_int64, V, C.c_int64]
    L.clsn_bucket_records.argtypes = [V, I, P(C.c_int64), P(V)]
    L.clsn_set_stream.argtypes = [V, V]
    L.clsn_apply_stage.argtypes = [V, I, I]
    L.clsn_state_device_ptrs.argtypes = [V, P(V), P(V), P(V)]
    L.clsn_timer_start.argtypes = [V]
    L.clsn_timer_stop.argtypes = [V, P(C.c_float)]
    L.clsn_launch_count.restype = C.c_int64
    L.clsn_launch_count.argtypes = [V, I]
    L.clsn_synchronize.argtypes = [V]
    L.clsn_set_debug.argtypes = [V, I, I]
    L.clsn_set_exact_stats.argtypes = [V, I]
    L.clsn_set_pipeline.argtypes = [V, I]
    L.clsn_set_phase_timing.argtypes = [V, I]
    L.clsn_num_candidates.restype = C.c_int64
    L.clsn_num_candidates.argtypes = [V]
    L.clsn_get_candidates.argtypes = [V, P(C.c_int32)]
    L.clsn_num_contacts.restype = C.c_int64
    L.clsn_num_contacts.argtypes = [V]
    L.clsn_get_contacts.argtypes = [V, V]
    L.clsn_get_accumulators.argtypes = [V, P(D), P(D), P(C.c_int32), P(D), P(C.c_int32)]
    L.clsn_set_body_accumulators.argtypes = [V, P(D), P(C.c_int32)]
    L.clsn_set_impact_zones.argtypes = [V, I, I]
    L.clsn_set_rest_lengths.argtypes = [V, P(D), P(D)]
    L.clsn_set_strain_limiting.argtypes = [V, I]
    L.clsn_strain_limit.argtypes = [V, P(C.c_int32), P(C.c_int32)]
    L.clsn_compute_impact_zone.argtypes = [V, I, P(clsn_zone_stats)]
    L.clsn_update_rigid_bodies.argtypes = [V, P(D), P(D)]
    L.clsn_dist_unique_id.argtypes = [V]
    L.clsn_dist_init.argtypes = [V, I, I, V]
    L.clsn_dist_nranks.argtypes = [V]
    _lib = L
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _bp(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


class Context:
    """Thin RAII wrapper of clsn_ctx (one per GPU)."""

    def __init__(self, device: int = 0):
        self.L = load_library()
        h = C.c_void_p()
        rc = self.L.clsn_create(C.byref(h), int(device))
        if rc != 0 or not h:
            raise CollisionError(f"clsn_create failed ({rc}): no usable CUDA device {device}; there is no CPU fallback")
        self.h = h
        self.V = 0
        self.nbody = 0

    def close(self):
        if getattr(self, "h", None):
            self.L.clsn_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != 0:
            raise CollisionError(f"collision_b200 error {rc}: {self.L.clsn_last_error(self.h).decode()}")


class CollisionSolver3d:
    """Drop-in shaped like the reference's CollisionSolver3d (collid.h:230-244)."""

    # process-global statics, like CollisionSolver::s_* (dcollid.cpp:28-35)
    s_eps = 1e-6
    s_thickness = 1e-4
    s_dt = 1e-3
    s_k = 1000.0
    s_m = 0.01
    s_lambda = 0.02
    s_cr = 0.0

    def __init__(self, device: int = 0, impact_zones: bool = True, max_zone_iterations: int = 0,
                 strain_limiting: bool = True):
        """impact_zones: enter computeImpactZone when MAX_ITER CCD passes leave collisions, as the
        reference's detectCollision does (dcollid.cpp:464-467); max_zone_iterations <= 0 keeps the
        reference's unbounded loop (guarded at 100000).  strain_limiting: run reduceSuperelast
        between the final positions and velocities, as resolveCollision does (:355)."""
        self.ctx = Context(device)
        self.setImpactZones(impact_zones, max_zone_iterations)
        self.setStrainLimiting(strain_limiting)
        self.scene = None
        self.has_collision = False
        self.last_stats = None
        self._lo = np.full(3, -1e30)
        self._hi = np.full(3, 1e30)
        self._x_old = None

    # ---- static parameter API (dcollid.cpp:57-89)
    @classmethod
    def setRoundingTolerance(cls, v): cls.s_eps = float(v)
    @classmethod
    def getRoundingTolerance(cls): return cls.s_eps
    @classmethod
    def setFabricThickness(cls, v): cls.s_thickness = float(v)
    @classmethod
    def getFabricThickness(cls): return cls.s_thickness
    @classmethod
    def setTimeStepSize(cls, v): cls.s_dt = float(v)
    @classmethod
    def getTimeStepSize(cls): return cls.s_dt
    @classmethod
    def setSpringConstant(cls, v): cls.s_k = float(v)
    @classmethod
    def getSpringConstant(cls): return cls.s_k
    @classmethod
    def setFrictionConstant(cls, v): cls.s_lambda = float(v)
    @classmethod
    def getFrictionConstant(cls): return cls.s_lambda
    @classmethod
    def setPointMass(cls, v): cls.s_m = float(v)
    @classmethod
    def getPointMass(cls): return cls.s_m
    @classmethod
    def setRestitutionCoef(cls, v): cls.s_cr = float(v)
    @classmethod
    def getRestitutionCoef(cls): return cls.s_cr

    @classmethod
    def set_params_from(cls, params):
        """Convenience: copy a scenes.Params into the statics."""
        cls.s_eps, cls.s_thickness, cls.s_k = params.eps, params.thickness, params.k
        cls.s_m, cls.s_lambda, cls.s_cr = params.m, params.friction, params.cr

    # ---- multi-GPU inside the library (include/collision_b200.h: clsn_dist_*)
    @staticmethod
    def dist_unique_id() -> bytes:
        """128 bytes to hand to every rank (called by one rank)."""
        buf = C.create_string_buffer(128)
        rc = load_library().clsn_dist_unique_id(C.cast(buf, C.c_void_p))
        if rc != 0:
            raise CollisionError(f"clsn_dist_unique_id failed ({rc}): NCCL not available")
        return buf.raw

    def dist_init(self, rank: int, nranks: int, unique_id: bytes):
        """Join `nranks` contexts (one per GPU of this node); call after assembleFromInterface.  resolveCollision then
        runs the distributed step and every rank returns the complete, identical result."""
        assert len(unique_id) == 128
        buf = C.create_string_buffer(unique_id, 128)
        self.ctx.check(self.ctx.L.clsn_dist_init(self.ctx.h, int(rank), int(nranks), C.cast(buf, C.c_void_p)))

    def dist_nranks(self) -> int:
        return int(self.ctx.L.clsn_dist_nranks(self.ctx.h))

    def setDomainBoundary(self, L, U):
        self._lo = np.asarray(L, dtype=np.float64).copy()
        self._hi = np.asarray(U, dtype=np.float64).copy()

    def getDomainBoundary(self, d, side):
        return float((self._lo, self._hi)[side][d])

    def hasCollision(self):
        return self.has_collision

    def setImpactZones(self, on=True, max_iterations=0):
        self.ctx.check(self.ctx.L.clsn_set_impact_zones(self.ctx.h, 1 if on else 0, int(max_iterations)))
        self.impact_zones = bool(on)
        self.max_zone_iterations = int(max_iterations)

    def setStrainLimiting(self, on=True):
        self.ctx.check(self.ctx.L.clsn_set_strain_limiting(self.ctx.h, 1 if on else 0))
        self.strain_limiting = bool(on)

    def setRestLengths(self, tri_len0, bond_len0):
        """TRI::side_length0[3] / BOND::length0 (the application's data in the reference)."""
        a = np.ascontiguousarray(tri_len0, dtype=np.float64)
        b = np.ascontiguousarray(bond_len0, dtype=np.float64)
        assert a.size == 3 * self.scene.T and b.size == self.scene.B
        self.ctx.check(self.ctx.L.clsn_set_rest_lengths(self.ctx.h, _dp(a), _dp(b)))

    def reduceSuperelast(self):
        """Strain limiting alone on the resident avgVel (dcollid.cpp:586-596): (sweeps, edges in the last)."""
        self._push_params()
        n, e = C.c_int32(), C.c_int32()
        self.ctx.check(self.ctx.L.clsn_strain_limit(self.ctx.h, C.byref(n), C.byref(e)))
        return int(n.value), int(e.value)

    def computeImpactZone(self, max_iterations=0):
        """The fail-safe loop alone, on the resident state (dcollid.cpp:227-265)."""
        self._push_params()
        zs = clsn_zone_stats()
        self.ctx.check(self.ctx.L.clsn_compute_impact_zone(self.ctx.h, int(max_iterations), C.byref(zs)))
        return zs.as_dict()

    # ---- assembly
    def assembleFromInterface(self, scene, dt):
        """Gather topology + flags once per mesh (dcollid3d.cpp:12-52): setTimeStepSize(dt), element
        list (triangles of every surface, then bonds of every string curve), rigid-body lists,
        domain boundary."""
        self.setTimeStepSize(dt)
        if scene is not self.scene:
            c = self.ctx
            tri = np.ascontiguousarray(scene.tri_idx, dtype=np.int32)
            surf = np.ascontiguousarray(scene.tri_surf, dtype=np.int32)
            bond = np.ascontiguousarray(scene.bond_idx, dtype=np.int32)
            fl = np.ascontiguousarray(scene.vflags, dtype=np.uint8)
            vb = np.ascontiguousarray(scene.vhs, dtype=np.int32)
            mass = np.ascontiguousarray(scene.hs_mass, dtype=np.float64)
            c.check(c.L.clsn_set_topology(c.h, scene.V, scene.T, _ip(tri), _ip(surf), scene.B, _ip(bond), _bp(fl),
                                          _ip(vb), len(mass), _dp(mass)))
            c.V, c.nbody = scene.V, len(mass)
            self.scene = scene
            self.setRestLengths(*scene.rest_lengths())
        self.setDomainBoundary(scene.lo, scene.hi)

    def _push_params(self):
        p = clsn_params()
        cls = type(self)
        p.eps, p.thickness, p.dt, p.k, p.m, p.lambda_, p.cr = (cls.s_eps, cls.s_thickness, cls.s_dt, cls.s_k, cls.s_m,
                                                             cls.s_lambda, cls.s_cr)
        for i in range(3):
            p.lo[i] = self._lo[i]
            p.hi[i] = self._hi[i]
        self.ctx.check(self.ctx.L.clsn_set_params(self.ctx.h, C.byref(p)))

    def recordOriginPosition(self, x):
        """x_old <- Coords for non-movable points (dcollid.cpp:91-107)."""
        if self._x_old is None:
            self._x_old = np.array(x, dtype=np.float64, copy=True)
        else:
            keep = (self.scene.vflags & 2) != 0
            self._x_old[~keep] = x[~keep]

    # ---- the step
    def updateFinalForRG(self, center_of_mass, center_of_mass_velo):
        """dcollid.cpp:626-675: (nbody,3) arrays of the caller's HYPER_SURF data, updated in place for every movable body
        that collided in the last step (the reference calls this at the end of updateFinalVelocity)."""
        a, b = center_of_mass, center_of_mass_velo
        assert a.dtype == np.float64 and a.flags.c_contiguous and b.dtype == np.float64 and b.flags.c_contiguous
        assert a.shape == (self.ctx.nbody, 3) and b.shape == (self.ctx.nbody, 3)
        self.ctx.check(self.ctx.L.clsn_update_rigid_bodies(self.ctx.h, _dp(a), _dp(b)))

    def resolveCollision(self, x_old, x, vel, x_out=None, bodies=None):
        """x_old: start-of-step positions; x: candidate positions on entry, final positions on return
        (or written to x_out when given); vel: updated where has_collsn (updateFinalVelocity);
        bodies: optional (center_of_mass, center_of_mass_velo), (nbody,3) each, updated like updateFinalForRG does.
        Returns has_collsn (V,) uint8."""
        c = self.ctx
        self._push_params()
        assert x.dtype == np.float64 and x.flags.c_contiguous and vel.dtype == np.float64 and vel.flags.c_contiguous
        xo = np.ascontiguousarray(x_old, dtype=np.float64)
        has = np.zeros(c.V, dtype=np.uint8)
        st = clsn_step_stats()
        out = x if x_out is None else x_out
        assert out.dtype == np.float64 and out.flags.c_contiguous
        c.check(c.L.clsn_step_host(c.h, _dp(xo), _dp(x), _dp(out), _dp(vel), _bp(has), C.byref(st)))
        self.has_collision = bool(st.has_collision)
        self.last_stats = st.as_dict()
        if bodies is not None:
            self.updateFinalForRG(*bodies)
        return has

    # ---- single phases / readbacks (parity tests)
    def upload(self, x_old, x_new):
        c = self.ctx
        self._push_params()
        a = np.ascontiguousarray(x_old, dtype=np.float64)
        b = np.ascontiguousarray(x_new, dtype=np.float64)
        c.check(c.L.clsn_upload_state(c.h, _dp(a), _dp(b)))

    def avg_velocity(self):
        self.ctx.check(self.ctx.L.clsn_avg_velocity(self.ctx.h))

    def detect(self, mode):
        self._push_params()
        st = clsn_pass_stats()
        self.ctx.check(self.ctx.L.clsn_detect(self.ctx.h, mode, C.byref(st)))
        return st.as_dict()

    def apply(self, rigidify=True):
        self.ctx.check(self.ctx.L.clsn_apply(self.ctx.h, 1 if rigidify else 0))

    def boundary(self):
        self._push_params()
        self.ctx.check(self.ctx.L.clsn_boundary(self.ctx.h))

    def final_position(self):
        self.ctx.check(self.ctx.L.clsn_final_position(self.ctx.h))

    def resolve_device(self):
        """clsn_resolve on whatever state is resident (after upload)."""
        self._push_params()
        st = clsn_step_stats()
        self.ctx.check(self.ctx.L.clsn_resolve(self.ctx.h, C.byref(st)))
        self.has_collision = bool(st.has_collision)
        self.last_stats = st.as_dict()
        return self.last_stats

    def upload_device(self, d_x_old_ptr, d_x_new_ptr):
        """x_old / x_new already in HBM (3V doubles each, e.g. torch tensors' data_ptr())."""
        self._push_params()
        self.ctx.check(self.ctx.L.clsn_upload_state_device(self.ctx.h, d_x_old_ptr, d_x_new_ptr))

    def timer_start(self):
        self.ctx.check(self.ctx.L.clsn_timer_start(self.ctx.h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        self.ctx.check(self.ctx.L.clsn_timer_stop(self.ctx.h, C.byref(ms)))
        return float(ms.value)

    def launch_count(self, reset=False) -> int:
        return int(self.ctx.L.clsn_launch_count(self.ctx.h, 1 if reset else 0))

    def synchronize(self):
        self.ctx.check(self.ctx.L.clsn_synchronize(self.ctx.h))

    def set_avgvel(self, av):
        a = np.ascontiguousarray(av, dtype=np.float64)
        self.ctx.check(self.ctx.L.clsn_set_avgvel(self.ctx.h, _dp(a)))

    def download(self):
        c = self.ctx
        x = np.empty((c.V, 3))
        av = np.empty((c.V, 3))
        has = np.empty(c.V, dtype=np.uint8)
        c.check(c.L.clsn_download_state(c.h, _dp(x), _dp(av), _bp(has)))
        return x, av, has

    def set_debug(self, candidates=True, contacts=True):
        self.ctx.check(self.ctx.L.clsn_set_debug(self.ctx.h, int(candidates), int(contacts)))

    def set_exact_stats(self, on=True):
        """full traversal in every pass so that stats['candidates'] equals the reference's callback count"""
        self.ctx.check(self.ctx.L.clsn_set_exact_stats(self.ctx.h, int(on)))

    def set_phase_timing(self, on: bool):
        """per-phase device times in last_stats["ms_phase"] (CUDA-event marks between the kernel groups of a step); off by default"""
        self.ctx.check(self.ctx.L.clsn_set_phase_timing(self.ctx.h, int(bool(on))))

    def set_pipeline(self, pipeline: int):
        """1: plain-FP64 fast path, correctly rounded solve of the undecided features only (default);
        0: staged correctly rounded solve of every feature; 2: experimental, 1 + impulse records emitted straight into
        per-point segments.  Same results."""
        self.ctx.check(self.ctx.L.clsn_set_pipeline(self.ctx.h, int(pipeline)))

    def candidates(self):
        c = self.ctx
        n = c.L.clsn_num_candidates(c.h)
        out = np.empty((n, 2), dtype=np.int32)
        if n:
            c.check(c.L.clsn_get_candidates(c.h, _ip(out)))
        return out

    def contacts(self):
        c = self.ctx
        n = c.L.clsn_num_contacts(c.h)
        out = np.empty(n, dtype=CONTACT_DTYPE)
        if n:
            c.check(c.L.clsn_get_contacts(c.h, out.ctypes.data_as(C.c_void_p)))
        return out

    def accumulators(self):
        c = self.ctx
        imp = np.empty((c.V, 3))
        fric = np.empty((c.V, 3))
        cnt = np.empty(c.V, dtype=np.int32)
        irg = np.empty((c.nbody, 3))
        crg = np.empty(c.nbody, dtype=np.int32)
        c.check(c.L.clsn_get_accumulators(c.h, _dp(imp), _dp(fric), _ip(cnt), _dp(irg), _ip(crg)))
        return imp, fric, cnt, irg, crg

    def set_body_accumulators(self, imp_rg, cnt_rg):
        a = np.ascontiguousarray(imp_rg, dtype=np.float64)
        b = np.ascontiguousarray(cnt_rg, dtype=np.int32)
        self.ctx.check(self.ctx.L.clsn_set_body_accumulators(self.ctx.h, _dp(a), _ip(b)))

    def close(self):
        self.ctx.close()
