"""Seeded synthetic scenes for the collision step (BASELINE.json `configs`).

A scene is plain arrays -- the same arrays are handed to the CUDA solver, to the C oracle
and to the compiled reference, so every comparison starts from identical bits.

Element order follows the reference's hseList (dcollid3d.cpp:25-43): all triangles,
surface by surface, then all bonds, curve by curve.  Element id e < T is triangle e,
e >= T is bond e - T.

The reference's own decks (in-string_string, in-ball_plane, in-box_boundary) are restated
geometrically: FronTier's level-set mesher is not reproducible, so surfaces are meshed here
(icosphere / box / strip) with the decks' centres, sizes, velocities and wave types.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

FABRIC, STATIC_RIGID, MOVABLE_RIGID = 0, 1, 2
FLAG_FIXED, FLAG_MOVABLE_RG = 1, 2


@dataclass
class Params:
    """CollisionSolver static parameters and their defaults (dcollid.cpp:28-35)."""
    eps: float = 1e-6          # rounding tolerance  s_eps
    thickness: float = 1e-4    # fabric thickness    s_thickness
    k: float = 1000.0          # spring constant     s_k
    m: float = 0.01            # point mass          s_m
    friction: float = 0.02     # friction constant   s_lambda  (test.cpp:106 sets 0)
    cr: float = 0.0            # restitution         s_cr

    def as_array(self) -> np.ndarray:
        return np.array([self.eps, self.thickness, self.k, self.m, self.friction, self.cr], dtype=np.float64)


@dataclass
class Scene:
    name: str
    x: np.ndarray              # (V,3) f64  start-of-step positions (x_old)
    vel: np.ndarray            # (V,3) f64
    tri_idx: np.ndarray        # (T,3) i32
    tri_surf: np.ndarray       # (T,)  i32 non-decreasing
    bond_idx: np.ndarray       # (B,2) i32
    bond_curve: np.ndarray     # (B,)  i32 non-decreasing
    hs_kind: np.ndarray        # (n_surf+n_curve,) i32   FABRIC / STATIC_RIGID / MOVABLE_RIGID
    hs_mass: np.ndarray        # (n_surf+n_curve,) f64   total_mass(hs)
    vflags: np.ndarray         # (V,) u8   bit0 is_fixed, bit1 is_movableRG
    vhs: np.ndarray            # (V,) i32  hyper-surface of each vertex
    dt: float
    lo: np.ndarray = field(default_factory=lambda: np.full(3, -1e30))
    hi: np.ndarray = field(default_factory=lambda: np.full(3, 1e30))
    params: Params = field(default_factory=Params)

    @property
    def V(self) -> int:
        return self.x.shape[0]

    @property
    def T(self) -> int:
        return self.tri_idx.shape[0]

    @property
    def B(self) -> int:
        return self.bond_idx.shape[0]

    @property
    def n_surf(self) -> int:
        return int(self.tri_surf.max()) + 1 if self.T else 0

    @property
    def n_curve(self) -> int:
        return int(self.bond_curve.max()) + 1 if self.B else 0

    def rest_lengths(self, x=None):
        """TRI::side_length0[3] / BOND::length0 -- the application's data in the reference, set from the
        unstretched mesh; here from `x` (default: the scene's initial positions).  Edge j of a
        triangle joins its points j and (j+1)%3.  Same operation order as FronTier's
        distance_between_positions: sqrt(((dx^2 + dy^2) + dz^2))."""
        x = self.x if x is None else x

        def dist(a, b):
            d = x[a] - x[b]
            return np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2])

        t = self.tri_idx
        tri_len0 = np.stack([dist(t[:, j], t[:, (j + 1) % 3]) for j in range(3)], axis=1) if self.T else np.zeros((0, 3))
        bond_len0 = dist(self.bond_idx[:, 0], self.bond_idx[:, 1]) if self.B else np.zeros(0)
        return np.ascontiguousarray(tri_len0, dtype=np.float64), np.ascontiguousarray(bond_len0, dtype=np.float64)

    def x_new(self) -> np.ndarray:
        """Candidate end-of-step positions, as the driver's spring solver leaves them
        (test.cpp:226-258): x_old + dt * vel."""
        return self.x + self.dt * self.vel


# ----------------------------------------------------------------------------- meshes
def grid_sheet(nx: int, ny: int, x0: float, x1: float, y0: float, y1: float):
    """nx x ny vertices, 2 triangles per cell, alternating diagonal-free (all same split)."""
    xs = np.linspace(x0, x1, nx)
    ys = np.linspace(y0, y1, ny)
    X, Y = np.meshgrid(xs, ys, indexing="ij")
    pts = np.stack([X.ravel(), Y.ravel(), np.zeros(nx * ny)], axis=1)
    i, j = np.meshgrid(np.arange(nx - 1), np.arange(ny - 1), indexing="ij")
    v00 = (i * ny + j).ravel()
    v10 = v00 + ny
    v01 = v00 + 1
    v11 = v10 + 1
    tris = np.empty((2 * v00.size, 3), dtype=np.int32)
    tris[0::2] = np.stack([v00, v10, v11], axis=1)
    tris[1::2] = np.stack([v00, v11, v01], axis=1)
    return pts, tris


def icosphere(level: int, center, radius: float):
    t = (1.0 + 5.0 ** 0.5) / 2.0
    verts = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t),
             (0, -1, -t), (0, 1, -t), (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    faces = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4),
             (11, 10, 2), (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8),
             (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    v = np.array(verts, dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array(faces, dtype=np.int64)
    for _ in range(level):
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], axis=0)
        es = np.sort(e, axis=1)
        key = es[:, 0] * (v.shape[0] + 1) + es[:, 1]
        uniq, inv = np.unique(key, return_inverse=True)
        a = uniq // (v.shape[0] + 1)
        b = uniq % (v.shape[0] + 1)
        mid = v[a] + v[b]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        base = v.shape[0]
        v = np.concatenate([v, mid], axis=0)
        n = f.shape[0]
        m01 = base + inv[0:n]
        m12 = base + inv[n:2 * n]
        m20 = base + inv[2 * n:3 * n]
        f = np.concatenate([
            np.stack([f[:, 0], m01, m20], axis=1),
            np.stack([f[:, 1], m12, m01], axis=1),
            np.stack([f[:, 2], m20, m12], axis=1),
            np.stack([m01, m12, m20], axis=1)], axis=0)
    pts = np.asarray(center, dtype=np.float64)[None, :] + radius * v
    return pts, f.astype(np.int32)


def box_surface(center, edge, cells):
    """Closed triangulated cuboid surface; `edge` = full edge lengths, `cells` = (nx,ny,nz)."""
    center = np.asarray(center, dtype=np.float64)
    half = np.asarray(edge, dtype=np.float64) / 2.0
    nx, ny, nz = cells
    gx = np.linspace(-half[0], half[0], nx + 1)
    gy = np.linspace(-half[1], half[1], ny + 1)
    gz = np.linspace(-half[2], half[2], nz + 1)
    ids = -np.ones((nx + 1, ny + 1, nz + 1), dtype=np.int64)
    pts = []

    def vid(i, j, k):
        if ids[i, j, k] < 0:
            ids[i, j, k] = len(pts)
            pts.append((gx[i], gy[j], gz[k]))
        return ids[i, j, k]

    tris = []

    def quad(a, b, c, d):
        tris.append((a, b, c))
        tris.append((a, c, d))

    for i in range(nx):
        for j in range(ny):
            for k in (0, nz):
                quad(vid(i, j, k), vid(i + 1, j, k), vid(i + 1, j + 1, k), vid(i, j + 1, k))
    for i in range(nx):
        for k in range(nz):
            for j in (0, ny):
                quad(vid(i, j, k), vid(i + 1, j, k), vid(i + 1, j, k + 1), vid(i, j, k + 1))
    for j in range(ny):
        for k in range(nz):
            for i in (0, nx):
                quad(vid(i, j, k), vid(i, j + 1, k), vid(i, j + 1, k + 1), vid(i, j, k + 1))
    return np.asarray(pts, dtype=np.float64) + center[None, :], np.asarray(tris, dtype=np.int32)


# ----------------------------------------------------------------------------- assembly
class _Builder:
    def __init__(self):
        self.x, self.vel, self.flags, self.vhs = [], [], [], []
        self.tris, self.tri_surf = [], []
        self.bonds, self.bond_curve = [], []
        self.surf_kind, self.surf_mass = [], []
        self.curve_kind, self.curve_mass = [], []
        self.nv = 0

    def add_surface(self, pts, tris, vel, kind=FABRIC, mass=None, point_mass=0.01):
        s = len(self.surf_kind)
        n = pts.shape[0]
        self.x.append(pts)
        self.vel.append(np.broadcast_to(np.asarray(vel, dtype=np.float64), (n, 3)).copy())
        flag = FLAG_FIXED if kind == STATIC_RIGID else (FLAG_MOVABLE_RG if kind == MOVABLE_RIGID else 0)
        self.flags.append(np.full(n, flag, dtype=np.uint8))
        self.vhs.append(np.full(n, -1 - s, dtype=np.int64))  # fixed up in build()
        self.tris.append(tris.astype(np.int64) + self.nv)
        self.tri_surf.append(np.full(tris.shape[0], s, dtype=np.int32))
        self.surf_kind.append(kind)
        self.surf_mass.append(float(n * point_mass if mass is None else mass))
        self.nv += n
        return s

    def add_curve(self, pts, vel):
        c = len(self.curve_kind)
        n = pts.shape[0]
        self.x.append(pts)
        self.vel.append(np.broadcast_to(np.asarray(vel, dtype=np.float64), (n, 3)).copy())
        self.flags.append(np.zeros(n, dtype=np.uint8))
        self.vhs.append(np.full(n, 1_000_000 + c, dtype=np.int64))
        idx = np.arange(n - 1, dtype=np.int64) + self.nv
        self.bonds.append(np.stack([idx, idx + 1], axis=1))
        self.bond_curve.append(np.full(n - 1, c, dtype=np.int32))
        self.curve_kind.append(FABRIC)
        self.curve_mass.append(0.0)
        self.nv += n
        return c

    def build(self, name, dt, lo=None, hi=None, params=None) -> Scene:
        n_surf = len(self.surf_kind)
        x = np.ascontiguousarray(np.concatenate(self.x, axis=0), dtype=np.float64)
        vel = np.ascontiguousarray(np.concatenate(self.vel, axis=0), dtype=np.float64)
        vhs = np.concatenate(self.vhs)
        vhs = np.where(vhs < 0, -1 - vhs, vhs - 1_000_000 + n_surf).astype(np.int32)
        tri = (np.concatenate(self.tris, axis=0) if self.tris else np.zeros((0, 3))).astype(np.int32)
        tsurf = (np.concatenate(self.tri_surf) if self.tris else np.zeros(0)).astype(np.int32)
        bond = (np.concatenate(self.bonds, axis=0) if self.bonds else np.zeros((0, 2))).astype(np.int32)
        bcur = (np.concatenate(self.bond_curve) if self.bonds else np.zeros(0)).astype(np.int32)
        sc = Scene(name=name, x=x, vel=vel, tri_idx=np.ascontiguousarray(tri),
                   tri_surf=np.ascontiguousarray(tsurf), bond_idx=np.ascontiguousarray(bond),
                   bond_curve=np.ascontiguousarray(bcur),
                   hs_kind=np.array(self.surf_kind + self.curve_kind, dtype=np.int32),
                   hs_mass=np.array(self.surf_mass + self.curve_mass, dtype=np.float64),
                   vflags=np.concatenate(self.flags).astype(np.uint8), vhs=vhs, dt=float(dt))
        if lo is not None:
            sc.lo = np.asarray(lo, dtype=np.float64)
            sc.hi = np.asarray(hi, dtype=np.float64)
        if params is not None:
            sc.params = params
        return sc


# ----------------------------------------------------------------------------- configs
def string_string(dt: float = 0.01, gap: float = 0.01) -> Scene:
    """Config 1 -- in-string_string: two crossing strings, 192 bonds each (bond length =
    span / floor(span / (0.25*h)), h = 0.5/50, cdinit.cpp:157-159), moving at +-0.2 in z."""
    b = _Builder()
    nb = int(0.48 / (0.25 * 0.01))
    s = np.linspace(0.0, 1.0, nb + 1)[:, None]
    p0 = np.array([0.25, 0.01, 0.25 - gap]) + s * np.array([0.0, 0.48, 0.0])
    p1 = np.array([0.01, 0.25, 0.25]) + s * np.array([0.48, 0.0, 0.0])
    b.add_curve(p0, (0, 0, 0.2))
    b.add_curve(p1, (0, 0, -0.2))
    return b.build("string_string", dt, lo=(0, 0, 0), hi=(0.5, 0.5, 0.5), params=Params(friction=0.0))


def ball_plane(dt: float = 0.005, level: int = 4, gap: float | None = None, seed: int = 7) -> Scene:
    """Config 2a -- in-ball_plane: movable rigid sphere c=(.25,.25,.4) r=.05 v=(0,0,.2) against a
    fabric cuboid c=(.25,.25,.47) edge=(.15,.15,.01) v=(0,0,-.2), domain z in [.25,.75].
    `gap` overrides the initial clearance so a single step already collides."""
    rng = np.random.default_rng(seed)
    b = _Builder()
    cz = 0.47
    if gap is not None:
        cz = 0.4 + 0.05 + gap + 0.005
    sp, st = icosphere(level, (0.25, 0.25, 0.4), 0.05)
    b.add_surface(sp, st, (0, 0, 0.2), kind=MOVABLE_RIGID)
    bp, bt = box_surface((0.25, 0.25, cz), (0.15, 0.15, 0.01), (15, 15, 1))
    bp = bp + 1e-6 * rng.uniform(-1, 1, bp.shape)
    b.add_surface(bp, bt, (0, 0, -0.2), kind=FABRIC)
    return b.build("ball_plane", dt, lo=(0, 0, 0.25), hi=(0.5, 0.5, 0.75), params=Params(friction=0.0))


def box_boundary(dt: float = 0.005) -> Scene:
    """Config 2b -- in-box_boundary: one movable rigid cuboid c=(.25,.25,.21) edge=(.1,.1,.2)
    v=(.4,0,-.4) in [0,1]x[0,.5]^2.  Same-surface rigid tri pairs are filtered
    (dcollid.cpp:762,805) and the wall clamp skips movable-RG points (:123): a negative control."""
    b = _Builder()
    p, t = box_surface((0.25, 0.25, 0.21), (0.1, 0.1, 0.2), (5, 5, 10))
    b.add_surface(p, t, (0.4, 0, -0.4), kind=MOVABLE_RIGID)
    return b.build("box_boundary", dt, lo=(0, 0, 0), hi=(1.0, 0.5, 0.5), params=Params(friction=0.0))


def sheet_wall(n: int = 24, dt: float = 0.01, seed: int = 3) -> Scene:
    """Config 2c -- fabric sheet flying into the domain wall so the boundary clamp
    (dcollid.cpp:116-158) fires; friction on."""
    rng = np.random.default_rng(seed)
    b = _Builder()
    p, t = grid_sheet(n, n, 0.1, 0.4, 0.1, 0.4)
    p[:, 2] = 0.004 + 0.002 * np.sin(7 * p[:, 0]) + 1e-5 * rng.uniform(-1, 1, p.shape[0])
    s = b.add_surface(p, t, (0.0, 0.0, 0.0), kind=FABRIC)
    sc = b.build("sheet_wall", dt, lo=(0, 0, 0), hi=(0.5, 0.5, 0.5), params=Params(friction=0.3))
    sc.vel[:] = np.array([0.05, -0.02, -0.5]) + 1e-3 * rng.uniform(-1, 1, sc.vel.shape)
    return sc


def drape(n: int = 256, level: int = 5, dt: float = 1e-3, clearance: float = 2e-4, seed: int = 1234) -> Scene:
    """Config 3 -- n x n-vertex sheet falling at (0,0,-1) onto a static (is_fixed) icosphere
    r=.25 at (.5,.5,.3).  The sheet starts `clearance` above the pole so contact begins at once."""
    rng = np.random.default_rng(seed)
    b = _Builder()
    sp, st = icosphere(level, (0.5, 0.5, 0.3), 0.25)
    b.add_surface(sp, st, (0, 0, 0), kind=STATIC_RIGID)
    p, t = grid_sheet(n, n, 0.0, 1.0, 0.0, 1.0)
    p[:, 2] = 0.55 + clearance
    p += 1e-5 * rng.uniform(-1, 1, p.shape)
    b.add_surface(p, t, (0, 0, -1.0), kind=FABRIC)
    return b.build(f"drape_{n}", dt, lo=(-1, -1, -1), hi=(2, 2, 2))


def layered_cloth(n_layers: int = 8, n: int = 251, dt: float = 1e-3, speed: float = 0.3,
                  amp: float = 0.02, seed: int = 2024, friction: float = 0.02) -> Scene:
    """Config 4 (n_layers=8, n=251 -> 1 000 000 triangles, 504 008 vertices) and its smaller
    siblings: nested folded layers, gaps 2..10 x thickness, alternating +-speed normal
    velocities, seeded jitter.  Every layer is its own fabric surface."""
    rng = np.random.default_rng(seed)
    b = _Builder()
    thickness = 1e-4
    z = 0.3
    for l in range(n_layers):
        p, t = grid_sheet(n, n, 0.05, 0.95, 0.05, 0.95)
        fold = amp * np.sin(2 * np.pi * 1.5 * p[:, 0]) * np.sin(2 * np.pi * p[:, 1])
        p[:, 2] = z + fold
        p += 1e-5 * rng.uniform(-1, 1, p.shape)
        v = np.zeros_like(p)
        v[:, 2] = speed if l % 2 == 0 else -speed
        v += 1e-3 * rng.uniform(-1, 1, p.shape)
        b.add_surface(p, t, (0, 0, 0), kind=FABRIC)
        b.vel[-1] = v
        z += thickness * rng.uniform(2.0, 10.0)
    return b.build(f"layered_{n_layers}x{n}", dt, lo=(-1, -1, -1), hi=(2, 2, 2),
                   params=Params(friction=friction))


def cloth_spheres(n_layers: int = 8, n: int = 501, n_side: int = 4, level: int = 3, dt: float = 1e-3,
                  seed: int = 4242) -> Scene:
    """Config 5 -- sheet stack + n_side^3 movable rigid icospheres on a seeded lattice with
    |v| ~ 5..20 (displacement per step >> edge length)."""
    rng = np.random.default_rng(seed)
    base = layered_cloth(n_layers, n, dt, seed=seed)
    b = _Builder()
    for s in range(base.n_surf):
        sel = base.tri_surf == s
        vids = np.unique(base.tri_idx[sel])
        remap = -np.ones(base.V, dtype=np.int64)
        remap[vids] = np.arange(vids.size)
        b.add_surface(base.x[vids], remap[base.tri_idx[sel]].astype(np.int32), (0, 0, 0))
        b.vel[-1] = base.vel[vids].copy()
    g = (np.arange(n_side) + 0.5) / n_side
    # z levels: symmetric about the stack (z = 0.3 +- 0.02 of fold + 5 mm of layers), the inner pair 4 cm off its mid-plane
    # (3.5 mm clear of the highest fold: in reach within one step at |v| >= 5), further levels 6.5 cm apart -- more than twice
    # the reach of a sphere in one step (radius 1.2 cm + 2 cm of travel), so that the swept volumes of two spheres never
    # meet: the scene is about fast rigid bodies in a cloth stack, not about rigid bodies ramming each other.
    half = [0.04 + 0.065 * i for i in range((n_side + 1) // 2)]
    zs = sorted([0.3025 - h for h in half] + [0.3025 + h for h in half])[: n_side] if n_side % 2 == 0 else \
        sorted([0.3025] + [0.3025 - h - 0.025 for h in half[: n_side // 2]] + [0.3025 + h + 0.025 for h in half[: n_side // 2]])
    for cx in g:
        for cy in g:
            for cz in zs:
                c = np.array([0.05 + 0.9 * cx, 0.05 + 0.9 * cy, cz]) + np.array([0.01, 0.01, 0.0005]) * rng.uniform(-1, 1, 3)
                d = rng.normal(size=3)
                d /= np.linalg.norm(d)
                sp, st = icosphere(level, c, 0.012)
                b.add_surface(sp, st, d * rng.uniform(5.0, 20.0), kind=MOVABLE_RIGID)
    return b.build(f"cloth_spheres_{n_layers}x{n}+{n_side ** 3}", dt, lo=(-5, -5, -5), hi=(5, 5, 5))


def two_sheets(n: int = 12, dt: float = 1e-3, gap: float = 3e-4, speed: float = 0.3, seed: int = 11,
               friction: float = 0.02) -> Scene:
    """Small two-layer approach case used all over the parity tests."""
    return layered_cloth(2, n, dt, speed=speed, amp=0.01, seed=seed, friction=friction)


def mixed(seed: int = 5, dt: float = 2e-3) -> Scene:
    """All element-pair types at once: fabric sheet, static-rigid sphere, movable-rigid
    sphere and two strings (tri-tri, tri-bond, bond-bond; fabric/static/movable branches)."""
    rng = np.random.default_rng(seed)
    b = _Builder()
    p, t = grid_sheet(20, 20, 0.1, 0.4, 0.1, 0.4)
    p[:, 2] = 0.25 + 0.004 * np.sin(20 * p[:, 0]) * np.cos(17 * p[:, 1])
    p += 1e-5 * rng.uniform(-1, 1, p.shape)
    b.add_surface(p, t, (0, 0, -0.25))
    sp, st = icosphere(2, (0.2, 0.2, 0.2185), 0.03)
    b.add_surface(sp, st, (0, 0, 0), kind=STATIC_RIGID)
    mp, mt = icosphere(2, (0.3, 0.3, 0.282), 0.03)
    b.add_surface(mp, mt, (0, 0, -0.6), kind=MOVABLE_RIGID)
    m2, mt2 = icosphere(1, (0.3, 0.3, 0.343), 0.03)
    b.add_surface(m2, mt2, (0, 0, -1.5), kind=MOVABLE_RIGID)
    s = np.linspace(0, 1, 61)[:, None]
    c0 = np.array([0.12, 0.15, 0.2512]) + s * np.array([0.26, 0.02, 0.0]) + 1e-5 * rng.uniform(-1, 1, (61, 3))
    c1 = np.array([0.15, 0.12, 0.2519]) + s * np.array([0.01, 0.26, 0.0]) + 1e-5 * rng.uniform(-1, 1, (61, 3))
    b.add_curve(c0, (0, 0, -0.1))
    b.add_curve(c1, (0, 0, -0.6))
    return b.build("mixed", dt, lo=(0, 0, 0), hi=(0.5, 0.5, 0.5), params=Params(friction=0.05, cr=0.2))
