"""ctypes binding of oracle/libcollision_oracle.so (the C restatement) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcollision_oracle.so")

PROXIMITY, COLLISION = 0, 1
F_X_OLD, F_X, F_AVGVEL, F_IMP, F_FRIC = range(5)
I_CNT, I_HAS_COLLSN = range(2)

CONTACT_DTYPE = np.dtype([("ea", "<i4"), ("eb", "<i4"), ("feature", "<i4"), ("kind", "<i4"), ("p", "<i4", (4,)),
                          ("root", "<f8"), ("dist", "<f8"), ("nor", "<f8", (3,)), ("w", "<f8", (3,))])
assert CONTACT_DTYPE.itemsize == 96

_lib = None


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc, no FMA contraction).  Building the checker is not using it."""
    src = os.path.join(_HERE, "collision_oracle.c")
    hdr = os.path.join(_HERE, "collision_oracle.h")
    if (not force and os.path.exists(LIB_PATH)
            and os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return LIB_PATH
    subprocess.check_call(["make", "-s", "-C", _HERE, "port"])
    return LIB_PATH


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _bp(a):
    return a.ctypes.data_as(C.POINTER(C.c_ubyte))


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        P, D, I, V = C.POINTER, C.c_double, C.c_int, C.c_void_p
        L.orc_create.restype = V
        L.orc_create.argtypes = [I, I, P(I), P(I), I, P(I), P(C.c_ubyte), P(I), I, P(D)]
        L.orc_destroy.argtypes = [V]
        L.orc_set_params.argtypes = [V] + [D] * 6
        L.orc_set_domain.argtypes = [V, P(D), P(D)]
        L.orc_set_dt.argtypes = [V, D]
        L.orc_set_state.argtypes = [V, P(D), P(D)]
        L.orc_set_avgvel.argtypes = [V, P(D)]
        L.orc_avg_velocity.argtypes = [V]
        L.orc_detect.restype = C.c_long
        L.orc_detect.argtypes = [V, I]
        L.orc_detect_ordered.restype = C.c_long
        L.orc_detect_ordered.argtypes = [V, I, P(I), C.c_long]
        L.orc_apply.argtypes = [V, I]
        L.orc_boundary.argtypes = [V]
        L.orc_final_position.argtypes = [V]
        L.orc_final_velocity.argtypes = [V, P(D)]
        L.orc_resolve.argtypes = [V, P(D), P(C.c_long)]
        L.orc_update_final_for_rg.argtypes = [V, P(D), P(D)]
        L.orc_set_has_collsn.argtypes = [V, P(C.c_ubyte)]
        L.orc_enable_impact_zones.argtypes = [V, I]
        L.orc_set_imp_zone.argtypes = [V, I]
        L.orc_zone_velocity.argtypes = [V]
        L.orc_zone_velocity.restype = I
        L.orc_impact_zone.argtypes = [V, I, P(C.c_long)]
        L.orc_set_rest_lengths.argtypes = [V, P(D), P(D)]
        L.orc_enable_strain_limiting.argtypes = [V, I]
        L.orc_strain_limit_once.argtypes = [V, P(C.c_long)]
        L.orc_strain_limit.argtypes = [V, P(C.c_long)]
        L.orc_get_f64.argtypes = [V, I, P(D)]
        L.orc_get_i32.argtypes = [V, I, P(I)]
        L.orc_get_body.argtypes = [V, P(D), P(I)]
        L.orc_set_body.argtypes = [V, P(D), P(I)]
        for n in ("orc_num_candidates", "orc_num_contacts", "orc_num_true_pairs"):
            getattr(L, n).restype = C.c_long
            getattr(L, n).argtypes = [V]
        L.orc_get_candidates.argtypes = [V, P(I)]
        L.orc_get_true_pairs.argtypes = [V, P(I)]
        L.orc_get_contacts.argtypes = [V, V]
        L.orc_set_libm.argtypes = [I]
        L.orc_feature.argtypes = [I, P(D), P(D), P(D), P(C.c_ubyte), P(D), D, D, P(D), P(D), P(D), P(D)]
        _lib = L
    return _lib


LIBM_NATIVE, LIBM_CR = 0, 1


def set_libm(mode: int):
    """LIBM_NATIVE: host libm as the reference uses it; LIBM_CR: correctly rounded (the CUDA contract)."""
    lib().orc_set_libm(int(mode))


class OracleSolver:
    """The C restatement driven on a collision_b200.scenes.Scene."""

    def __init__(self, scene, impact_zones=True, strain_limiting=True):
        """impact_zones: resolve() enters computeImpactZone when the CCD passes are exhausted, like the
        reference's detectCollision (dcollid.cpp:464-467).  strain_limiting: resolve() runs
        reduceSuperelast (:355) with the scene's rest lengths."""
        L = lib()
        self.scene = scene
        self.V, self.T, self.B = scene.V, scene.T, scene.B
        self.nhs = len(scene.hs_kind)
        k = [np.ascontiguousarray(a) for a in (
            scene.tri_idx.astype(np.int32), scene.tri_surf.astype(np.int32), scene.bond_idx.astype(np.int32),
            scene.vflags.astype(np.uint8), scene.vhs.astype(np.int32), scene.hs_mass.astype(np.float64))]
        self.h = L.orc_create(self.V, self.T, _ip(k[0]), _ip(k[1]), self.B, _ip(k[2]), _bp(k[3]), _ip(k[4]),
                              self.nhs, _dp(k[5]))
        p = scene.params
        L.orc_set_params(self.h, p.eps, p.thickness, p.k, p.m, p.friction, p.cr)
        lo = np.ascontiguousarray(scene.lo, dtype=np.float64)
        hi = np.ascontiguousarray(scene.hi, dtype=np.float64)
        L.orc_set_domain(self.h, _dp(lo), _dp(hi))
        L.orc_set_dt(self.h, scene.dt)
        L.orc_enable_impact_zones(self.h, 1 if impact_zones else 0)
        self.set_rest_lengths(*scene.rest_lengths())
        L.orc_enable_strain_limiting(self.h, 1 if strain_limiting else 0)

    def close(self):
        if self.h:
            lib().orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_dt(self, dt):
        lib().orc_set_dt(self.h, float(dt))

    def set_state(self, x_old, x_new):
        a = np.ascontiguousarray(x_old, dtype=np.float64)
        b = np.ascontiguousarray(x_new, dtype=np.float64)
        lib().orc_set_state(self.h, _dp(a), _dp(b))

    def set_avgvel(self, av):
        a = np.ascontiguousarray(av, dtype=np.float64)
        lib().orc_set_avgvel(self.h, _dp(a))

    def avg_velocity(self):
        lib().orc_avg_velocity(self.h)

    def detect(self, mode) -> int:
        return int(lib().orc_detect(self.h, mode))

    def detect_ordered(self, mode, pairs) -> int:
        """Replay ordered (a, b) pairs in the given order (the reference's callback sequence)."""
        pr = np.ascontiguousarray(pairs[:, :2], dtype=np.int32)
        return int(lib().orc_detect_ordered(self.h, mode, _ip(pr), pr.shape[0]))

    def apply(self, rigidify=True):
        lib().orc_apply(self.h, 1 if rigidify else 0)

    def boundary(self):
        lib().orc_boundary(self.h)

    def final_position(self):
        lib().orc_final_position(self.h)

    def final_velocity(self, vel):
        assert vel.dtype == np.float64 and vel.flags.c_contiguous
        lib().orc_final_velocity(self.h, _dp(vel))

    def set_has_collsn(self, has):
        a = np.ascontiguousarray(has, dtype=np.uint8)
        lib().orc_set_has_collsn(self.h, _bp(a))

    def update_final_for_rg(self, com, com_velo):
        """updateFinalForRG (dcollid.cpp:626-675) on caller-owned (nhs,3) centre-of-mass arrays, in place."""
        assert com.dtype == np.float64 and com.flags.c_contiguous and com_velo.dtype == np.float64 and com_velo.flags.c_contiguous
        lib().orc_update_final_for_rg(self.h, _dp(com), _dp(com_velo))

    def resolve(self, vel):
        assert vel.dtype == np.float64 and vel.flags.c_contiguous
        stats = (C.c_long * 20)()
        lib().orc_resolve(self.h, _dp(vel), stats)
        return list(stats)

    def set_rest_lengths(self, tri_len0, bond_len0):
        a = np.ascontiguousarray(tri_len0, dtype=np.float64)
        b = np.ascontiguousarray(bond_len0, dtype=np.float64)
        lib().orc_set_rest_lengths(self.h, _dp(a), _dp(b))

    def enable_strain_limiting(self, on=True):
        lib().orc_enable_strain_limiting(self.h, 1 if on else 0)

    def strain_limit(self):
        """reduceSuperelast: (sweeps run, edges averaged in the last sweep)"""
        n = C.c_long()
        it = lib().orc_strain_limit(self.h, C.byref(n))
        return int(it), int(n.value)

    def enable_impact_zones(self, on=True):
        lib().orc_enable_impact_zones(self.h, 1 if on else 0)

    def set_imp_zone(self, on):
        lib().orc_set_imp_zone(self.h, 1 if on else 0)

    def zone_velocity(self) -> int:
        return int(lib().orc_zone_velocity(self.h))

    def impact_zone(self, max_iter=0):
        out = (C.c_long * 3)()
        lib().orc_impact_zone(self.h, int(max_iter), out)
        return list(out)

    def get(self, field) -> np.ndarray:
        out = np.empty((self.V, 3), dtype=np.float64)
        lib().orc_get_f64(self.h, field, _dp(out))
        return out

    def geti(self, field) -> np.ndarray:
        out = np.empty(self.V, dtype=np.int32)
        lib().orc_get_i32(self.h, field, _ip(out))
        return out

    def get_body(self):
        imp = np.empty((self.nhs, 3), dtype=np.float64)
        cnt = np.empty(self.nhs, dtype=np.int32)
        lib().orc_get_body(self.h, _dp(imp), _ip(cnt))
        return imp, cnt

    def set_body(self, imp, cnt):
        imp = np.ascontiguousarray(imp, dtype=np.float64)
        cnt = np.ascontiguousarray(cnt, dtype=np.int32)
        lib().orc_set_body(self.h, _dp(imp), _ip(cnt))

    def candidates(self) -> np.ndarray:
        n = lib().orc_num_candidates(self.h)
        out = np.empty((n, 2), dtype=np.int32)
        if n:
            lib().orc_get_candidates(self.h, _ip(out))
        return out

    def true_pairs(self) -> np.ndarray:
        n = lib().orc_num_true_pairs(self.h)
        out = np.empty((n, 2), dtype=np.int32)
        if n:
            lib().orc_get_true_pairs(self.h, _ip(out))
        return out

    def contacts(self) -> np.ndarray:
        n = lib().orc_num_contacts(self.h)
        out = np.empty(n, dtype=CONTACT_DTYPE)
        if n:
            lib().orc_get_contacts(self.h, out.ctypes.data_as(C.c_void_p))
        return out


def feature(kind, x_old, coords, avg_vel, flags, mass, h, dt, params):
    x_old = np.ascontiguousarray(x_old, dtype=np.float64).reshape(12)
    coords = np.ascontiguousarray(coords, dtype=np.float64).reshape(12)
    avg_vel = np.ascontiguousarray(avg_vel, dtype=np.float64).reshape(12)
    flags = np.ascontiguousarray(flags, dtype=np.uint8).reshape(4)
    mass = np.ascontiguousarray(mass, dtype=np.float64).reshape(4)
    params = np.ascontiguousarray(params, dtype=np.float64).reshape(6)
    roots = np.zeros(4)
    acc = np.zeros((4, 10))
    hit = C.c_double(-1.0)
    r = lib().orc_feature(kind, _dp(x_old), _dp(coords), _dp(avg_vel), _bp(flags), _dp(mass), float(h),
                          float(dt), _dp(params), _dp(roots), _dp(acc), C.byref(hit))
    return dict(ret=int(r), roots=roots, acc=acc, hit_root=float(hit.value))
