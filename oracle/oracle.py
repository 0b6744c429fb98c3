"""oracle.py — TEST INFRASTRUCTURE (oracle/). Not part of the product path.

ctypes/numpy front-end to (a) the CPU restatement ``libnans_oracle.so`` and (b) the
harness around the reference's own prebuilt ``nans.so`` (``oracle/_ref``).  Only
tests/, ``__graft_entry__.smoke()`` and bench.py's CPU-baseline legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libnans_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REF_HARNESS_SO = os.path.join(REF_DIR, "libnans_ref_harness.so")

CC, CS, CF, SS, SF = 0, 1, 2, 3, 4

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int32)


def _fp(a):
    return a.ctypes.data_as(f32p) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(i32p) if a is not None else None


def build(force: bool = False) -> None:
    """Compile the restatement (and the reference harness when possible)."""
    if force or not os.path.exists(ORACLE_SO) or (
            os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(HERE, "nans_oracle.c"))):
        subprocess.check_call(["make", "-C", HERE, "libnans_oracle.so"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)


# --------------------------------------------------------------------------- restatement
class WorldStruct(C.Structure):
    _fields_ = [("n_cubes", C.c_int32), ("n_spheres", C.c_int32), ("n_statics", C.c_int32),
                ("pos", f32p), ("vel", f32p), ("force", f32p),
                ("ang", f32p), ("angvel", f32p), ("torque", f32p),
                ("mass", f32p), ("moi", f32p), ("scale", f32p), ("radius", f32p), ("verts", f32p),
                ("st_pos", f32p), ("st_ang", f32p), ("st_scale", f32p),
                ("st_mass", f32p), ("st_moi", f32p), ("st_verts", f32p)]


CONTACT_DTYPE = np.dtype([("type", "<i4"), ("a", "<i4"), ("b", "<i4"),
                          ("point_a", "<f4", 3), ("point_b", "<f4", 3), ("n", "<f4", 3)])
assert CONTACT_DTYPE.itemsize == 48

NP_STATS_DTYPE = np.dtype([("gjk_result", "<i4"), ("gjk_iters", "<i4"), ("epa_iters", "<i4"),
                           ("max_faces", "<i4"), ("max_edges", "<i4"), ("emptied", "<i4")])

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build()
        L = C.CDLL(ORACLE_SO)
        L.oracle_detect_collisions.restype = C.c_int
        L.oracle_step.restype = C.c_int
        L.oracle_check_collision.restype = C.c_int
        _lib = L
    return _lib


class World:
    """numpy-backed world in the restatement's layout (cubes first, then spheres)."""

    FIELDS3 = ("pos", "vel", "force", "ang", "angvel", "torque", "scale")

    def __init__(self, n_cubes: int, n_spheres: int = 0, n_statics: int = 1):
        nb = n_cubes + n_spheres
        self.n_cubes, self.n_spheres, self.n_statics = n_cubes, n_spheres, n_statics
        for f in self.FIELDS3:
            setattr(self, f, np.zeros((nb, 3), np.float32))
        self.scale[:] = 1.0
        self.mass = np.ones(nb, np.float32)
        self.moi = np.ones(nb, np.float32)
        self.radius = np.zeros(nb, np.float32)
        self.verts = np.zeros((n_cubes, 8, 3), np.float32)
        self.st_pos = np.zeros((n_statics, 3), np.float32)
        self.st_ang = np.zeros((n_statics, 3), np.float32)
        self.st_scale = np.ones((n_statics, 3), np.float32)
        self.st_mass = np.ones(n_statics, np.float32)
        self.st_moi = np.ones(n_statics, np.float32)
        self.st_verts = np.zeros((n_statics, 8, 3), np.float32)

    @property
    def nb(self):
        return self.n_cubes + self.n_spheres

    def copy(self) -> "World":
        w = World.__new__(World)
        for k, v in self.__dict__.items():
            setattr(w, k, v.copy() if isinstance(v, np.ndarray) else v)
        return w

    def struct(self) -> WorldStruct:
        s = WorldStruct(self.n_cubes, self.n_spheres, self.n_statics)
        for f in ("pos", "vel", "force", "ang", "angvel", "torque", "mass", "moi", "scale", "radius",
                  "verts", "st_pos", "st_ang", "st_scale", "st_mass", "st_moi", "st_verts"):
            a = getattr(self, f)
            assert a.dtype == np.float32 and a.flags.c_contiguous, f
            setattr(s, f, _fp(a))
        return s

    # stage calls -----------------------------------------------------------
    def integrate_forces(self, dt):
        s = self.struct(); lib().oracle_integrate_forces(C.byref(s), C.c_float(dt))

    def integrate_velocities(self, dt):
        s = self.struct(); lib().oracle_integrate_velocities(C.byref(s), C.c_float(dt))

    def rebuild_vertices(self):
        s = self.struct(); lib().oracle_rebuild_vertices(C.byref(s))

    def detect(self, cap=None, prefilter=False) -> np.ndarray:
        """prefilter: False/0 = the reference's all-pairs loop; True/1 = all pairs, AABB-rejected; "grid"/2 = uniform
        grid over the same AABB predicate (same list, same order; seconds at 10^6 bodies)."""
        prefilter = 2 if prefilter == "grid" else int(prefilter)
        cap = cap or max(64, 16 * self.nb)
        out = np.zeros(cap, CONTACT_DTYPE)
        s = self.struct()
        n = lib().oracle_detect_collisions(C.byref(s), out.ctypes.data_as(C.c_void_p), cap, int(prefilter))
        assert n <= cap, "contact capacity exceeded"
        return out[:n].copy()

    def solve(self, dt, contacts: np.ndarray):
        contacts = np.ascontiguousarray(contacts)
        s = self.struct()
        lib().oracle_solve_constraints(C.byref(s), C.c_float(dt), contacts.ctypes.data_as(C.c_void_p),
                                       len(contacts))

    def step(self, dt, prefilter=False, cap=None) -> np.ndarray:
        prefilter = 2 if prefilter == "grid" else int(prefilter)
        cap = cap or max(64, 16 * self.nb)
        out = np.zeros(cap, CONTACT_DTYPE)
        s = self.struct()
        n = lib().oracle_step(C.byref(s), C.c_float(dt), out.ctypes.data_as(C.c_void_p), cap, int(prefilter))
        assert n <= cap
        return out[:n].copy()


def model_vertices(pos, ang, scale):
    pos = np.asarray(pos, np.float32); ang = np.asarray(ang, np.float32); scale = np.asarray(scale, np.float32)
    model = np.zeros(16, np.float32); verts = np.zeros(24, np.float32)
    lib().oracle_model_vertices(_fp(pos), _fp(ang), _fp(scale), _fp(model), _fp(verts))
    return model.reshape(4, 4), verts.reshape(8, 3)


def check_collision_batch(type_, posA, vertsA, radA, posB, vertsB, radB, want_stats=False):
    """Restatement CheckCollision over n pairs. Returns dict(hit,gjk,N,PA,PB[,stats])."""
    n = len(type_)
    type_ = np.ascontiguousarray(type_, np.int32)
    arrs = [np.ascontiguousarray(a, np.float32) for a in (posA, vertsA, radA, posB, vertsB, radB)]
    hit = np.zeros(n, np.int32); gjk = np.zeros(n, np.int32)
    N = np.zeros((n, 3), np.float32); PA = np.zeros((n, 3), np.float32); PB = np.zeros((n, 3), np.float32)
    stats = np.zeros(n, NP_STATS_DTYPE) if want_stats else None
    lib().oracle_check_collision_batch(n, _ip(type_), *[_fp(a) for a in arrs], _ip(hit), _ip(gjk),
                                       _fp(N), _fp(PA), _fp(PB),
                                       stats.ctypes.data_as(C.c_void_p) if want_stats else None)
    r = dict(hit=hit, gjk=gjk, N=N, PA=PA, PB=PB)
    if want_stats:
        r["stats"] = stats
    return r


# --------------------------------------------------------------------------- reference binary
VEC3 = ("<f4", 3)
CUBE_DTYPE = np.dtype([("Model", "<f4", 16), ("Vertices", "<f4", (8, 3)),
                       ("Position", *VEC3), ("V", *VEC3), ("Forces", *VEC3),
                       ("Angles", *VEC3), ("W", *VEC3), ("Torque", *VEC3),
                       ("Size", "<f4"), ("Mass", "<f4"), ("MOI", "<f4")])
SPHERE_DTYPE = np.dtype([("Model", "<f4", 16),
                         ("Position", *VEC3), ("V", *VEC3), ("Forces", *VEC3),
                         ("Angles", *VEC3), ("W", *VEC3), ("Torque", *VEC3),
                         ("Radius", "<f4"), ("Mass", "<f4"), ("MOI", "<f4")])
CAMERA_DTYPE = np.dtype([("FOV", "<f4"), ("Pitch", "<f4"), ("Yaw", "<f4"), ("Speed", "<f4"),
                         ("Position", *VEC3), ("Target", *VEC3), ("Direction", *VEC3),
                         ("Up", *VEC3), ("Front", *VEC3), ("Right", *VEC3)])
LIGHT_DTYPE = np.dtype([("raw", "<f4", 16)])
STATE_DTYPE = np.dtype([("Cubes", CUBE_DTYPE, 16), ("Spheres", SPHERE_DTYPE, 16), ("Floor", CUBE_DTYPE),
                        ("Camera", CAMERA_DTYPE), ("Lights", LIGHT_DTYPE, 4),
                        ("CubeCount", "<u4"), ("SphereCount", "<u4"), ("_pad", "<u4"),
                        ("Pairs", "<u8", 3)])
REF_PAIR_DTYPE = np.dtype([("Type", "<i4"), ("DLNormal", "<f4"), ("DLNormalSum", "<f4"),
                           ("DLTangent1", "<f4"), ("DLTangent1Sum", "<f4"),
                           ("DLTangent2", "<f4"), ("DLTangent2Sum", "<f4"),
                           ("PointA", *VEC3), ("PointB", *VEC3), ("N", *VEC3), ("T1", *VEC3), ("T2", *VEC3),
                           ("IndexA", "<i4"), ("IndexB", "<i4")])
INPUT_DTYPE = np.dtype([("Buttons", "<i4", (13, 2)), ("Sensitivity", "<f4"),
                        ("XRel", "<i4"), ("YRel", "<i4"), ("X", "<i4"), ("Y", "<i4")])
assert CUBE_DTYPE.itemsize == 244 and SPHERE_DTYPE.itemsize == 148
assert STATE_DTYPE.itemsize == 6896 and STATE_DTYPE.fields["Pairs"][1] == 6872
assert STATE_DTYPE.fields["Floor"][1] == 6272 and STATE_DTYPE.fields["CubeCount"][1] == 6860
assert REF_PAIR_DTYPE.itemsize == 96 and INPUT_DTYPE.itemsize == 124


class MemoryStruct(C.Structure):
    _fields_ = [("PermanentStorage", C.c_void_p), ("PermanentStorageSize", C.c_uint64),
                ("TransientStorage", C.c_void_p), ("TransientStorageSize", C.c_uint64),
                ("IsInitialized", C.c_int32)]


_ref = None
_ref_tried = False


def ref_so_path():
    for p in (os.environ.get("NANS_REF_SO"), os.path.join(REF_DIR, "nans.so")):
        if p and os.path.exists(p):
            return p
    return None


def ref():
    """Harness around the reference's prebuilt nans.so, or None when not staged."""
    global _ref, _ref_tried
    if _ref_tried:
        return _ref
    _ref_tried = True
    so = ref_so_path()
    if so is None or not os.path.exists(REF_HARNESS_SO):
        try:
            build()
        except Exception:
            pass
        so = ref_so_path()
    if so is None or not os.path.exists(REF_HARNESS_SO):
        return None
    H = C.CDLL(REF_HARNESS_SO, mode=C.RTLD_GLOBAL)
    rc = H.nansref_load(so.encode())
    if rc != 0:
        raise RuntimeError(f"nansref_load({so}) failed: {rc}")
    for name in ("nansref_check_collision", "nansref_detect_collisions", "nansref_physics_step"):
        getattr(H, name).restype = C.c_int
    _ref = H
    return _ref


def ref_new_state() -> np.ndarray:
    return np.zeros(1, STATE_DTYPE)


def _sp(state):
    return state.ctypes.data_as(C.c_void_p)


def ref_check_collision_batch(type_, posA, vertsA, radA, posB, vertsB, radB):
    H = ref()
    n = len(type_)
    type_ = np.ascontiguousarray(type_, np.int32)
    arrs = [np.ascontiguousarray(a, np.float32) for a in (posA, vertsA, radA, posB, vertsB, radB)]
    hit = np.zeros(n, np.int32)
    N = np.zeros((n, 3), np.float32); PA = np.zeros((n, 3), np.float32); PB = np.zeros((n, 3), np.float32)
    H.nansref_check_collision_batch(n, _ip(type_), *[_fp(a) for a in arrs], _ip(hit), _fp(N), _fp(PA), _fp(PB))
    return dict(hit=hit, N=N, PA=PA, PB=PB)


def ref_detect(state, dt=1 / 60., cap=1024) -> np.ndarray:
    out = np.zeros(cap, REF_PAIR_DTYPE)
    n = ref().nansref_detect_collisions(_sp(state), C.c_float(dt), out.ctypes.data_as(C.c_void_p), cap)
    assert n <= cap
    return out[:n].copy()


def ref_physics_step(state, dt, cap=1024) -> np.ndarray:
    out = np.zeros(cap, REF_PAIR_DTYPE)
    n = ref().nansref_physics_step(_sp(state), C.c_float(dt), out.ctypes.data_as(C.c_void_p), cap)
    assert n <= cap
    return out[:n].copy()


# --------------------------------------------------------------------------- conversions
def world_from_ref_state(state) -> World:
    """Restatement world holding the same bodies as a reference sdl_state."""
    s = state[0]
    nc, ns = int(s["CubeCount"]), int(s["SphereCount"])
    w = World(nc, ns, 1)
    cubes, sph = s["Cubes"][:nc], s["Spheres"][:ns]
    for dst, src in (("pos", "Position"), ("vel", "V"), ("force", "Forces"),
                     ("ang", "Angles"), ("angvel", "W"), ("torque", "Torque")):
        getattr(w, dst)[:nc] = cubes[src]
        getattr(w, dst)[nc:] = sph[src]
    w.mass[:nc], w.mass[nc:] = cubes["Mass"], sph["Mass"]
    w.moi[:nc], w.moi[nc:] = cubes["MOI"], sph["MOI"]
    w.scale[:nc] = cubes["Size"][:, None]
    w.radius[nc:] = sph["Radius"]
    w.verts[:] = cubes["Vertices"]
    f = s["Floor"]
    w.st_pos[0], w.st_ang[0] = f["Position"], f["Angles"]
    w.st_scale[0] = (f["Size"], 1.0, f["Size"])
    w.st_mass[0], w.st_moi[0] = f["Mass"], f["MOI"]
    w.st_verts[0] = f["Vertices"]
    return w


def contacts_from_ref_pairs(pairs: np.ndarray) -> np.ndarray:
    out = np.zeros(len(pairs), CONTACT_DTYPE)
    out["type"], out["a"], out["b"] = pairs["Type"], pairs["IndexA"], pairs["IndexB"]
    fl = (pairs["Type"] == CF) | (pairs["Type"] == SF)
    out["b"][fl] = 0  # the reference stores IndexB = IndexA for floor pairs (code/nans.cpp:1407,1440)
    out["point_a"], out["point_b"], out["n"] = pairs["PointA"], pairs["PointB"], pairs["N"]
    return out
