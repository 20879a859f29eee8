"""ctypes binding of oracle/_ref/libcollision_ref.so -- TEST INFRASTRUCTURE.

The library is the reference's own AABB.cpp / dcollid.cpp / dcollid3d.cpp compiled unmodified
(oracle/Makefile, oracle/ref_wrapper.cpp).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libcollision_ref.so")
# the same reference objects linked against a correctly rounded acos/cos/sin/pow (oracle/cr_libm.c, oracle/Makefile)
LIB_PATH_CR = os.path.join(_HERE, "_ref", "libcollision_ref_cr.so")
LIBM_NATIVE, LIBM_CR = "native", "cr"

F_X_OLD, F_COORDS, F_VEL, F_AVGVEL, F_IMP, F_FRIC, F_IMP_RG = range(7)
I_CNT, I_CNT_RG, I_HAS_COLLSN = range(3)
(PH_AVG_VELOCITY, PH_PROXIMITY_DETECT, PH_APPLY, PH_COLLISION_DETECT, PH_BOUNDARY, PH_FINAL_POSITION,
 PH_STRAIN_LIMIT, PH_FINAL_VELOCITY, PH_DETECT_PROXIMITY, PH_DETECT_COLLISION,
 PH_IMPZONE_ON, PH_IMPZONE_OFF, PH_ZONE_VELOCITY, PH_COMPUTE_IMPACT_ZONE) = range(14)
K_ISCOPLANAR, K_POINT_TO_TRI, K_EDGE_TO_EDGE, K_MOVING_POINT_TO_TRI, K_MOVING_EDGE_TO_EDGE = range(5)

_libs = {}
_flavour = LIBM_NATIVE


def available(flavour: str = LIBM_NATIVE) -> bool:
    return os.path.exists(LIB_PATH_CR if flavour == LIBM_CR else LIB_PATH)


def set_libm(flavour: str):
    """Which build RefSolver() / feature() use from now on: LIBM_NATIVE = the host libm, exactly as the reference
    ships; LIBM_CR = the same objects with a correctly rounded libm bound in (what the CUDA path is held to)."""
    global _flavour
    assert flavour in (LIBM_NATIVE, LIBM_CR)
    _flavour = flavour


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _bp(a):
    return a.ctypes.data_as(C.POINTER(C.c_ubyte))


def lib(flavour: str | None = None):
    flavour = flavour or _flavour
    if flavour not in _libs:
        L = C.CDLL(LIB_PATH_CR if flavour == LIBM_CR else LIB_PATH)
        L.clsn_ref_create.restype = C.c_void_p
        L.clsn_ref_create.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int,
                                      C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int,
                                      C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_ubyte),
                                      C.POINTER(C.c_int)]
        L.clsn_ref_destroy.argtypes = [C.c_void_p]
        L.clsn_ref_set_params.argtypes = [C.c_void_p] + [C.c_double] * 6
        L.clsn_ref_set_domain.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.clsn_ref_set_rest_lengths.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.clsn_ref_set_state.argtypes = [C.c_void_p] + [C.POINTER(C.c_double)] * 3
        L.clsn_ref_get_f64.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        L.clsn_ref_set_f64.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        L.clsn_ref_get_i32.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.clsn_ref_set_i32.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.clsn_ref_assemble.argtypes = [C.c_void_p, C.c_double]
        L.clsn_ref_record_origin.argtypes = [C.c_void_p]
        L.clsn_ref_resolve.argtypes = [C.c_void_p, C.c_int]
        L.clsn_ref_phase.argtypes = [C.c_void_p, C.c_int]
        L.clsn_ref_record.argtypes = [C.c_void_p, C.c_int]
        L.clsn_ref_num_callbacks.restype = C.c_long
        L.clsn_ref_num_callbacks.argtypes = [C.c_void_p]
        L.clsn_ref_num_pairs.restype = C.c_long
        L.clsn_ref_num_pairs.argtypes = [C.c_void_p]
        L.clsn_ref_get_pairs.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.clsn_ref_clock.restype = C.c_double
        L.clsn_ref_clock.argtypes = [C.c_char_p]
        L.clsn_ref_feature.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                       C.POINTER(C.c_double), C.POINTER(C.c_ubyte), C.POINTER(C.c_double),
                                       C.c_double, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                       C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.clsn_ref_clock_reset.argtypes = []
        L.clsn_ref_set_body.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.clsn_ref_get_body.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.clsn_ref_feature_batch.restype = C.c_long
        L.clsn_ref_feature_batch.argtypes = [C.c_long, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double),
                                             C.POINTER(C.c_double), C.POINTER(C.c_ubyte), C.POINTER(C.c_double),
                                             C.c_double, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_int),
                                             C.POINTER(C.c_double)]
        _libs[flavour] = L
    return _libs[flavour]


class RefSolver:
    """The reference's CollisionSolver3d driven on a collision_b200.scenes.Scene."""

    def __init__(self, scene):
        L = self.L = lib()
        self.scene = scene
        self.V, self.T, self.B = scene.V, scene.T, scene.B
        self._keep = [np.ascontiguousarray(a) for a in (
            scene.tri_idx.astype(np.int32), scene.tri_surf.astype(np.int32),
            scene.bond_idx.astype(np.int32), scene.bond_curve.astype(np.int32),
            scene.hs_kind.astype(np.int32), scene.hs_mass.astype(np.float64),
            scene.vflags.astype(np.uint8), scene.vhs.astype(np.int32))]
        k = self._keep
        self.h = L.clsn_ref_create(self.V, self.T, _ip(k[0]), _ip(k[1]), scene.n_surf, self.B, _ip(k[2]),
                                   _ip(k[3]), scene.n_curve, _ip(k[4]), _dp(k[5]), _bp(k[6]), _ip(k[7]))
        p = scene.params
        L.clsn_ref_set_params(self.h, p.eps, p.thickness, p.k, p.m, p.friction, p.cr)
        lo = np.ascontiguousarray(scene.lo, dtype=np.float64)
        hi = np.ascontiguousarray(scene.hi, dtype=np.float64)
        L.clsn_ref_set_domain(self.h, _dp(lo), _dp(hi))

    def close(self):
        if self.h:
            self.L.clsn_ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- state
    def set_state(self, x_old, x_new, vel=None):
        x_old = np.ascontiguousarray(x_old, dtype=np.float64)
        x_new = np.ascontiguousarray(x_new, dtype=np.float64)
        v = None if vel is None else np.ascontiguousarray(vel, dtype=np.float64)
        self.L.clsn_ref_set_state(self.h, _dp(x_old), _dp(x_new), None if v is None else _dp(v))

    def set_rest_lengths(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        self.L.clsn_ref_set_rest_lengths(self.h, _dp(x))

    def get(self, field) -> np.ndarray:
        out = np.empty((self.V, 3), dtype=np.float64)
        self.L.clsn_ref_get_f64(self.h, field, _dp(out))
        return out

    def put(self, field, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        self.L.clsn_ref_set_f64(self.h, field, _dp(a))

    def geti(self, field) -> np.ndarray:
        out = np.empty(self.V, dtype=np.int32)
        self.L.clsn_ref_get_i32(self.h, field, _ip(out))
        return out

    def puti(self, field, a):
        a = np.ascontiguousarray(a, dtype=np.int32)
        self.L.clsn_ref_set_i32(self.h, field, _ip(a))

    def set_bodies(self, com, com_velo):
        """HYPER_SURF::center_of_mass / center_of_mass_velo of every hyper-surface ((nhs,3) arrays; caller-owned in the reference)"""
        com = np.ascontiguousarray(com, dtype=np.float64)
        vel = np.ascontiguousarray(com_velo, dtype=np.float64)
        for i in range(len(com)):
            self.L.clsn_ref_set_body(self.h, i, _dp(com[i]), _dp(vel[i]))

    def get_bodies(self, nhs):
        com, vel = np.zeros((nhs, 3)), np.zeros((nhs, 3))
        for i in range(nhs):
            a, b = np.zeros(3), np.zeros(3)
            self.L.clsn_ref_get_body(self.h, i, _dp(a), _dp(b))
            com[i], vel[i] = a, b
        return com, vel

    # ---- driving
    def assemble(self, dt):
        if self.L.clsn_ref_assemble(self.h, float(dt)) != 0:
            raise RuntimeError("reference aborted in assembleFromInterface")

    def resolve(self, strain_limiting=False):
        if self.L.clsn_ref_resolve(self.h, 1 if strain_limiting else 0) != 0:
            raise RuntimeError("reference aborted (clean_up(ERROR)) in resolveCollision")

    def phase(self, ph) -> int:
        r = self.L.clsn_ref_phase(self.h, ph)
        if r < 0:
            raise RuntimeError(f"reference aborted in phase {ph}")
        return r

    def record(self, on=True):
        self.L.clsn_ref_record(self.h, 1 if on else 0)

    def pairs(self) -> np.ndarray:
        """(n,3) int32: a, b, result for every narrow-phase callback since record()."""
        n = self.L.clsn_ref_num_pairs(self.h)
        out = np.empty((n, 3), dtype=np.int32)
        if n:
            self.L.clsn_ref_get_pairs(self.h, _ip(out))
        return out

    def num_callbacks(self) -> int:
        return int(self.L.clsn_ref_num_callbacks(self.h))


def clock(name: str) -> float:
    return float(lib().clsn_ref_clock(name.encode()))


def clock_reset():
    lib().clsn_ref_clock_reset()


def feature(kind, x_old, coords, avg_vel, flags, mass, h, dt, params):
    """One call of a dcollid3d.cpp primitive on 4 free points.
    Returns dict(ret, roots[4], acc[4,10], hit_root)."""
    x_old = np.ascontiguousarray(x_old, dtype=np.float64).reshape(12)
    coords = np.ascontiguousarray(coords, dtype=np.float64).reshape(12)
    avg_vel = np.ascontiguousarray(avg_vel, dtype=np.float64).reshape(12)
    flags = np.ascontiguousarray(flags, dtype=np.uint8).reshape(4)
    mass = np.ascontiguousarray(mass, dtype=np.float64).reshape(4)
    params = np.ascontiguousarray(params, dtype=np.float64).reshape(6)
    roots = np.zeros(4)
    acc = np.zeros((4, 10))
    hit = C.c_double(-1.0)
    r = lib().clsn_ref_feature(kind, _dp(x_old), _dp(coords), _dp(avg_vel), _bp(flags), _dp(mass),
                               float(h), float(dt), _dp(params), _dp(roots), _dp(acc), C.byref(hit))
    if r < 0:
        raise RuntimeError(f"reference feature call failed ({r})")
    return dict(ret=int(r), roots=roots, acc=acc, hit_root=float(hit.value))


def feature_batch(kind, pts, x_old, avg_vel, vflags, vmass, h, dt, params):
    """n calls of a dcollid3d.cpp primitive: kind[n], pts[n,4] index the per-vertex arrays.  Returns (ret[n], hit_root[n])."""
    kind = np.ascontiguousarray(kind, dtype=np.int32)
    pts = np.ascontiguousarray(pts, dtype=np.int32).reshape(-1, 4)
    n = len(kind)
    assert pts.shape[0] == n
    x_old = np.ascontiguousarray(x_old, dtype=np.float64)
    avg_vel = np.ascontiguousarray(avg_vel, dtype=np.float64)
    vflags = np.ascontiguousarray(vflags, dtype=np.uint8)
    vmass = np.ascontiguousarray(vmass, dtype=np.float64)
    params = np.ascontiguousarray(params, dtype=np.float64).reshape(6)
    ret = np.zeros(n, np.int32)
    hit = np.full(n, -1.0)
    if n:
        lib().clsn_ref_feature_batch(n, _ip(kind), _ip(pts), _dp(x_old), _dp(avg_vel), _bp(vflags), _dp(vmass), float(h),
                                     float(dt), _dp(params), _ip(ret), _dp(hit))
    if (ret < 0).any():
        raise RuntimeError("reference feature call aborted")
    return ret, hit
