"""Scene description and builders (host side, numpy only).

A :class:`Scene` is the host mirror of the reference's ``sdl_state`` world
(code/nans.h:374-386) generalised past its 16+16+1 body cap: dynamic bodies are
cubes ``[0, n_cubes)`` followed by spheres, statics are floor-type slabs (the
reference's ``Floor`` cube, code/nans.cpp:1667-1678).  Builders mirror the game
layer's ``Init`` (code/nans.cpp:1551-1717) and the benchmark configurations of
BASELINE.json / SURVEY.md §8(d).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32

# Reference vertex order, code/nans.cpp:395-407
CORNERS = np.array([[.5, .5, .5], [.5, .5, -.5], [-.5, .5, .5], [-.5, .5, -.5],
                    [.5, -.5, .5], [.5, -.5, -.5], [-.5, -.5, .5], [-.5, -.5, -.5]], F32)

# collision_type, code/nans.h:71-87
CC, CS, CF, SS, SF = 0, 1, 2, 3, 4


class Scene:
    """Plain numpy world state. All arrays float32, C-contiguous."""

    VEC_FIELDS = ("pos", "vel", "force", "ang", "angvel", "torque", "scale")

    def __init__(self, n_cubes: int, n_spheres: int = 0, n_statics: int = 1):
        nb = n_cubes + n_spheres
        self.n_cubes, self.n_spheres, self.n_statics = int(n_cubes), int(n_spheres), int(n_statics)
        for f in self.VEC_FIELDS:
            setattr(self, f, np.zeros((nb, 3), F32))
        self.scale[:] = 1.0
        self.mass = np.ones(nb, F32)
        self.moi = np.ones(nb, F32)
        self.radius = np.zeros(nb, F32)
        self.verts = np.zeros((n_cubes, 8, 3), F32)
        self.st_pos = np.zeros((n_statics, 3), F32)
        self.st_ang = np.zeros((n_statics, 3), F32)
        self.st_scale = np.ones((n_statics, 3), F32)
        self.st_mass = np.ones(n_statics, F32)
        self.st_moi = np.ones(n_statics, F32)
        self.st_verts = np.zeros((n_statics, 8, 3), F32)
        # optional: independent-world id per body (batched worlds, config C4); None = one world
        self.world_id = None

    @property
    def nb(self) -> int:
        return self.n_cubes + self.n_spheres

    ARRAYS = VEC_FIELDS + ("mass", "moi", "radius", "verts", "st_pos", "st_ang", "st_scale",
                           "st_mass", "st_moi", "st_verts")

    def copy(self) -> "Scene":
        s = Scene.__new__(Scene)
        for k, v in self.__dict__.items():
            setattr(s, k, v.copy() if isinstance(v, np.ndarray) else v)
        return s

    # -- body setup helpers (scalar inertia formulas of Init, code/nans.cpp:1609-1610,1663-1664)
    def set_cube(self, i, pos, size=1.0, mass=1.0, ang=(0, 0, 0), scale=None):
        self.pos[i] = pos
        self.ang[i] = ang
        self.mass[i] = mass
        self.moi[i] = F32(F32(mass) / F32(12.0)) * F32(F32(2.0) * F32(size) * F32(size))
        self.scale[i] = (size, size, size) if scale is None else scale

    def set_sphere(self, j, pos, radius=0.25, mass=2.0):
        i = self.n_cubes + j
        self.pos[i] = pos
        self.radius[i] = radius
        self.mass[i] = mass
        self.moi[i] = F32(F32(F32(2.0) / F32(5.0)) * F32(mass)) * F32(F32(radius) * F32(radius))
        self.scale[i] = (radius, radius, radius)

    def set_static(self, k, pos, scale, mass=100000.0, size_for_moi=100.0, ang=(0, 0, 0)):
        self.st_pos[k] = pos
        self.st_ang[k] = ang
        self.st_scale[k] = scale
        self.st_mass[k] = mass
        self.st_moi[k] = F32(F32(mass) / F32(12.0)) * F32(F32(2.0) * F32(size_for_moi) * F32(size_for_moi))

    def identity_vertices(self):
        """Frame-0 state of the reference: every Model is identity, so every cube's and the
        floor's collision shape is the unit cube at the origin (code/nans.cpp:1599-1600,1667-1668)."""
        self.verts[:] = CORNERS[None]
        self.st_verts[:] = CORNERS[None]


def demo_scene() -> Scene:
    """The reference's ``Init`` scene (code/nans.cpp:1598-1678): stack of 3 unit cubes, the
    (0.5,1,0.5)-scaled debug box (Cubes[3], :1930-1941), one sphere, the floor.  Vertices are
    the frame-0 identity-model ones, exactly as after ``Init``."""
    s = Scene(4, 1, 1)
    s.set_cube(0, (2.0, 3.5, 2.0))
    s.set_cube(1, (2.0, 1.0, 2.0))
    s.set_cube(2, (2.0, 4.5, 2.0))
    s.set_cube(3, (1.0, 1.0, 1.0), scale=(0.5, 1.0, 0.5))
    s.set_sphere(0, (0.1, 1.1, 1.1), radius=0.25, mass=2.0)
    s.set_static(0, (1.2, -0.5, 1.0), (100.0, 1.0, 100.0))
    s.identity_vertices()
    return s


def _lattice(n, dims, spacing, jitter, rng, origin):
    nx, ny, nz = dims
    idx = np.arange(n)
    # y-major: fill a full x-z layer before moving up
    y, rem = np.divmod(idx, nx * nz)
    z, x = np.divmod(rem, nx)
    p = np.stack([x, y, z], 1).astype(np.float64) * spacing + np.asarray(origin, np.float64)
    p += rng.uniform(-jitter, jitter, (n, 3))
    return p.astype(F32)


def cube_drop(n=10000, seed=1, dims=(22, 21, 22), spacing=1.25, jitter=0.1, walls=True) -> Scene:
    """Config C2: ``n`` unit cubes on a jittered lattice dropped into a static box made of the
    floor plus four floor-type wall slabs.  Exact lattices hit the reference's GJK
    collinearity degeneracy, hence the jitter."""
    rng = np.random.default_rng(seed)
    nx, ny, nz = dims
    assert n <= nx * ny * nz
    s = Scene(n, 0, 5 if walls else 1)
    lx, lz = nx * spacing, nz * spacing
    p = _lattice(n, dims, spacing, jitter, rng, (0.75, 1.0, 0.75))
    for i in range(n):
        s.set_cube(i, p[i])
    cx, cz = lx / 2 + 0.13, lz / 2 + 0.07           # off-centre: keeps GJK start directions generic
    size = float(max(lx, lz) + 8.0)
    s.set_static(0, (cx, -0.5, cz), (size, 1.0, size), size_for_moi=size)
    if walls:
        h = ny * spacing + 8.0
        s.set_static(1, (-0.5, h / 2 - 1.0, cz), (1.0, h, size), size_for_moi=size)
        s.set_static(2, (lx + 0.5, h / 2 - 1.0, cz), (1.0, h, size), size_for_moi=size)
        s.set_static(3, (cx, h / 2 - 1.0, -0.5), (size, h, 1.0), size_for_moi=size)
        s.set_static(4, (cx, h / 2 - 1.0, lz + 0.5), (size, h, 1.0), size_for_moi=size)
    return s


def cube_pile(n_side=100, seed=7, spacing=1.02, jitter=0.005, n=None, layers=None) -> Scene:
    """Config C5: ``n_side``^3 unit cubes (or the first ``n`` / ``layers`` layers), spacing 1.02,
    jittered, settling under gravity on one floor slab."""
    rng = np.random.default_rng(seed)
    ny = n_side if layers is None else layers
    total = n_side * n_side * ny
    n = total if n is None else min(n, total)
    s = Scene(n, 0, 1)
    p = _lattice(n, (n_side, ny, n_side), spacing, jitter, rng, (0.5, 0.52, 0.5))
    s.pos[:] = p
    s.mass[:] = 1.0
    s.moi[:] = F32(F32(1.0) / F32(12.0)) * F32(2.0)
    s.scale[:] = 1.0
    ext = n_side * spacing
    size = float(2 ** np.ceil(np.log2(ext + 16.0)))
    s.set_static(0, (ext / 2 + 0.13, -0.5, ext / 2 + 0.07), (size, 1.0, size), size_for_moi=size)
    return s


def cube_pile_slabs(n_slabs=2, side_x=100, ny=100, nz=100, seed=7, spacing=1.02, jitter=0.005, slab=None,
                    centre=False) -> Scene:
    """Config C5 for several GPUs: ONE pile of ``n_slabs * side_x`` x ``ny`` x ``nz`` unit cubes, numbered
    SLAB-MAJOR: slab r (x in [r*side_x, (r+1)*side_x)) holds the contiguous global indices
    [r*m, (r+1)*m), m = side_x*ny*nz, ordered inside like :func:`cube_pile` (layer by layer).  Contiguous index
    ranges are therefore spatial x-slabs, which is what csrc/slab.cu's exact decomposition needs.  ``slab=r``
    returns only that slab's bodies.  Every slab carries the SAME jitter pattern (the stream of ``cube_pile(seed)``),
    so a rank can build its share without the rest and every slab behaves like the single-GPU pile of that seed --
    how long a pile survives the reference's unstable one-pass solver depends on the jitter, and a
    slab that blows up early would make the halo exchange fail (loudly).  ``centre=True`` puts the middle of the
    pile's footprint at x = z = 0, which halves the largest coordinate (fp32 resolution there is what seeds the
    blow-up first)."""
    m = side_x * ny * nz
    which = range(n_slabs) if slab is None else [slab]
    s = Scene(m * len(which), 0, 1)
    x0 = -0.5 * n_slabs * side_x * spacing if centre else 0.0
    z0 = -0.5 * nz * spacing if centre else 0.0
    for k, r in enumerate(which):
        rng = np.random.default_rng(seed)
        s.pos[k * m:(k + 1) * m] = _lattice(m, (side_x, ny, nz), spacing, jitter, rng,
                                            (x0 + 0.5 + r * side_x * spacing, 0.52, z0 + 0.5))
    s.mass[:] = 1.0
    s.moi[:] = F32(F32(1.0) / F32(12.0)) * F32(2.0)
    s.scale[:] = 1.0
    # the floor: one static tile per slab, as wide as the slab (a single 1024-wide slab under an 8-slab pile puts
    # the reference's GJK/EPA far outside the range it behaves in: cubes 400 units from the floor's centre start
    # it almost parallel to the floor, and that pile blew apart by step 50).  Every rank holds all the tiles.
    w_x, ext_z = side_x * spacing, nz * spacing
    size_z = float(2 ** np.ceil(np.log2(ext_z + 16.0)))
    s.n_statics = n_slabs
    for f, shape in (("st_pos", 3), ("st_ang", 3), ("st_scale", 3)):
        setattr(s, f, np.zeros((n_slabs, shape), F32))
    s.st_mass, s.st_moi = np.ones(n_slabs, F32), np.ones(n_slabs, F32)
    s.st_verts = np.zeros((n_slabs, 8, 3), F32)
    for r in range(n_slabs):
        s.set_static(r, (x0 + (r + 0.5) * w_x, -0.5, z0 + ext_z / 2 + 0.07), (w_x, 1.0, size_z), size_for_moi=size_z)
    return s


def batched_worlds(n_worlds=4096, cubes_per=48, spheres_per=16, seed=0) -> Scene:
    """Config C4: independent worlds (RL-style batch), each ``cubes_per`` cubes + ``spheres_per``
    spheres over the same floor; bodies of different worlds never interact (``world_id``)."""
    nc, ns = n_worlds * cubes_per, n_worlds * spheres_per
    s = Scene(nc, ns, 1)
    wid = np.zeros(nc + ns, np.int32)
    for w in range(n_worlds):
        rng = np.random.default_rng(seed * 1000003 + w)
        pc = _lattice(cubes_per, (4, 3, 4), 1.3, 0.12, rng, (-2.0, 0.8, -2.0))
        ps = _lattice(spheres_per, (4, 1, 4), 1.3, 0.2, rng, (-2.0, 5.5, -2.0))
        c0, s0 = w * cubes_per, nc + w * spheres_per
        s.pos[c0:c0 + cubes_per] = pc
        s.pos[s0:s0 + spheres_per] = ps
        s.radius[s0:s0 + spheres_per] = rng.uniform(0.15, 0.4, spheres_per).astype(F32)
        wid[c0:c0 + cubes_per] = w
        wid[s0:s0 + spheres_per] = w
    s.mass[:nc] = 1.0
    s.moi[:nc] = F32(F32(1.0) / F32(12.0)) * F32(2.0)
    r = s.radius[nc:]
    s.mass[nc:] = 2.0
    s.moi[nc:] = (F32(0.4) * F32(2.0)) * (r * r)
    s.scale[nc:] = r[:, None]
    s.set_static(0, (1.2, -0.5, 1.0), (100.0, 1.0, 100.0))
    s.world_id = wid
    return s


def _rand_rot(rng, n):
    a = rng.uniform(-np.pi, np.pi, (n, 3))
    c, s = np.cos(a), np.sin(a)
    rx = np.zeros((n, 3, 3)); rx[:, 0, 0] = 1; rx[:, 1, 1] = c[:, 0]; rx[:, 1, 2] = -s[:, 0]; rx[:, 2, 1] = s[:, 0]; rx[:, 2, 2] = c[:, 0]
    ry = np.zeros((n, 3, 3)); ry[:, 1, 1] = 1; ry[:, 0, 0] = c[:, 1]; ry[:, 0, 2] = s[:, 1]; ry[:, 2, 0] = -s[:, 1]; ry[:, 2, 2] = c[:, 1]
    rz = np.zeros((n, 3, 3)); rz[:, 2, 2] = 1; rz[:, 0, 0] = c[:, 2]; rz[:, 0, 1] = -s[:, 2]; rz[:, 1, 0] = s[:, 2]; rz[:, 1, 1] = c[:, 2]
    return rx @ ry @ rz


def narrowphase_pairs(n, seed=1234, mix=(8, 7, 1), rotated=True, types=None):
    """Config C3: random shape pairs with WORLD-SPACE vertices as the input (sidesteps sinf/cosf
    parity).  Shape A at the origin, shape B centre U(-1.2,1.2)^3, unit cubes with
    Euler angles U(-pi,pi), sphere radii U(0.1,0.5).  ``mix`` = CC:CS:SS strata (8 Mi + 7 Mi + 1 Mi
    at the full 16 Mi size); ``types`` overrides with an explicit per-pair type array (CF/SF use a
    unit box as the 'floor' shape).  Returns a dict of arrays."""
    rng = np.random.default_rng(seed)
    if types is None:
        tot = float(sum(mix))
        n_cc = int(round(n * mix[0] / tot)); n_cs = int(round(n * mix[1] / tot))
        types = np.concatenate([np.full(n_cc, CC), np.full(n_cs, CS), np.full(n - n_cc - n_cs, SS)])
    types = np.ascontiguousarray(types, np.int32)
    n = len(types)
    pos_a = np.zeros((n, 3), F32)
    pos_b = rng.uniform(-1.2, 1.2, (n, 3)).astype(F32)
    ra = _rand_rot(rng, n) if rotated else np.broadcast_to(np.eye(3), (n, 3, 3))
    rb = _rand_rot(rng, n) if rotated else np.broadcast_to(np.eye(3), (n, 3, 3))
    verts_a = (np.einsum("nij,kj->nki", ra, CORNERS.astype(np.float64)) + pos_a[:, None, :]).astype(F32)
    verts_b = (np.einsum("nij,kj->nki", rb, CORNERS.astype(np.float64)) + pos_b[:, None, :]).astype(F32)
    rad_a = rng.uniform(0.1, 0.5, n).astype(F32)
    rad_b = rng.uniform(0.1, 0.5, n).astype(F32)
    return dict(type=types, pos_a=pos_a, verts_a=np.ascontiguousarray(verts_a), rad_a=rad_a,
                pos_b=pos_b, verts_b=np.ascontiguousarray(verts_b), rad_b=rad_b)


def random_small_world(rng, n_cubes, n_spheres, spread=2.0) -> Scene:
    """Random <=16+16 body world in the reference's vocabulary (one floor), used by parity fuzzing.
    Vertices are left for the caller to rebuild from (pos, ang, scale)."""
    s = Scene(n_cubes, n_spheres, 1)
    nb = s.nb
    s.pos[:] = rng.uniform(-spread, spread, (nb, 3))
    s.pos[:, 1] = rng.uniform(0.2, 3.0, nb)
    s.vel[:] = rng.normal(0, 2, (nb, 3))
    s.angvel[:] = rng.normal(0, 3, (nb, 3))
    s.ang[:] = rng.uniform(-200, 200, (nb, 3))
    s.force[:] = rng.normal(0, 5, (nb, 3)) * (rng.random((nb, 1)) < 0.3)
    s.torque[:] = rng.normal(0, 1, (nb, 3)) * (rng.random((nb, 1)) < 0.3)
    for i in range(n_cubes):
        s.set_cube(i, s.pos[i], size=1.0, mass=float(rng.uniform(0.5, 3.0)), ang=s.ang[i])
    for j in range(n_spheres):
        s.set_sphere(j, s.pos[n_cubes + j], radius=float(rng.uniform(0.1, 0.5)), mass=float(rng.uniform(0.5, 3.0)))
    s.set_static(0, (1.2, -0.5, 1.0), (100.0, 1.0, 100.0))
    return s
