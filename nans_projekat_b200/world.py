"""Host-side mirror of the reference's physics interface, over the C ABI (include/nans_b200.h).

Method names follow the reference's stage functions (code/nans.cpp): ``integrate_forces``
(IntegrateForces :975), ``detect_collisions`` (DetectCollisions :1352), ``solve_constraints``
(SolveConstraints :1539), ``integrate_velocities`` (IntegrateVelocities :1332 + the draw section's
model rebuild), ``step`` (the four in the order of :1758-1762), ``add_force`` / ``add_torque``
(CubeAddForce/CubeAddTorque/SphereAddForce :81-110), ``check_collision`` (CheckCollision :907).
Everything runs in libnans_b200.so on the GPU; this file only marshals numpy buffers.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .scenes import Scene

CONTACT_DTYPE = np.dtype([("type", "<i4"), ("a", "<i4"), ("b", "<i4"),
                          ("point_a", "<f4", 3), ("point_b", "<f4", 3), ("n", "<f4", 3)])


def _fp(a):
    return a.ctypes.data_as(_lib.f32p) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(_lib.i32p) if a is not None else None


def _view(scene, fields) -> _lib.SceneView:
    v = _lib.SceneView()
    keep = []
    for f in fields:
        a = getattr(scene, f, None)
        if a is None:
            continue
        if f == "world_id":
            a = np.ascontiguousarray(a, np.int32)
            v.world_id = _ip(a)
        else:
            a = np.ascontiguousarray(a, np.float32)
            setattr(v, f, _fp(a))
        keep.append((f, a))
    v._keepalive = keep
    return v


class World:
    """A device-resident world (one arena). Mirrors ``sdl_state`` + the physics stage functions."""

    ALL_FIELDS = Scene.ARRAYS + ("world_id",)

    def __init__(self, scene: Scene, device: int = 0, max_pairs: int = 0, max_contacts: int = 0,
                 stream: int | None = None, arena_ptr: int | None = None, arena_bytes: int = 0):
        L = _lib.lib()
        self.n_cubes, self.n_spheres, self.n_statics = scene.n_cubes, scene.n_spheres, scene.n_statics
        self.desc = _lib.WorldDesc(scene.n_cubes, scene.n_spheres, scene.n_statics, max_pairs, max_contacts,
                                   device, arena_ptr, arena_bytes, stream)
        self._h = C.c_void_p()
        _lib.check(L.nans_world_create(C.byref(self.desc), C.byref(self._h)))
        self.upload(scene)

    @staticmethod
    def arena_bytes(scene: Scene, max_pairs: int = 0, max_contacts: int = 0) -> int:
        d = _lib.WorldDesc(scene.n_cubes, scene.n_spheres, scene.n_statics, max_pairs, max_contacts, 0, None, 0, None)
        return int(_lib.lib().nans_world_arena_bytes(C.byref(d)))

    @property
    def nb(self):
        return self.n_cubes + self.n_spheres

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            _lib.lib().nans_world_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- state transfer ---------------------------------------------------------------
    def upload(self, scene, fields=None):
        v = _view(scene, fields or self.ALL_FIELDS)
        _lib.check(_lib.lib().nans_world_upload(self._h, C.byref(v)))

    def download(self, fields=("pos", "vel", "force", "ang", "angvel", "torque", "verts")) -> Scene:
        s = Scene(self.n_cubes, self.n_spheres, self.n_statics)
        self.download_into(s, fields)   # the view aliases the Scene's own (contiguous f32) arrays
        return s

    def download_into(self, scene, fields):
        v = _view(scene, fields)
        _lib.check(_lib.lib().nans_world_download(self._h, C.byref(v)))
        for f, a in v._keepalive:      # a field that was not a contiguous f32 array went through a copy: hand it back
            if a is not getattr(scene, f):
                getattr(scene, f)[...] = a

    # pipelined I/O (poses one frame late): copies overlap the step on their own streams
    def upload_async(self, scene, fields=("force", "torque")):
        v = _view(scene, fields)
        self._pending_views = getattr(self, "_pending_views", [])[-4:] + [v]   # keep the host arrays alive
        _lib.check(_lib.lib().nans_world_upload_async(self._h, C.byref(v)))

    def download_async(self, scene, fields=("pos", "ang")) -> int:
        v = _view(scene, fields)
        self._pending_views = getattr(self, "_pending_views", [])[-4:] + [v]
        t = C.c_int32(0)
        _lib.check(_lib.lib().nans_world_download_async(self._h, C.byref(v), C.byref(t)))
        return t.value

    def wait(self, ticket: int = -1):
        _lib.check(_lib.lib().nans_world_wait(self._h, ticket))

    def add_force(self, body_row: int, force=(0, 0, 0), torque=(0, 0, 0)):
        f = np.asarray(force, np.float32); t = np.asarray(torque, np.float32)
        _lib.check(_lib.lib().nans_world_add_force(self._h, body_row, _fp(f), _fp(t)))

    def add_torque(self, body_row: int, torque):
        self.add_force(body_row, (0, 0, 0), torque)

    def set_body(self, body_row: int, pos=None, vel=None, angvel=None):
        arrs = [None if x is None else np.asarray(x, np.float32) for x in (pos, vel, angvel)]
        _lib.check(_lib.lib().nans_world_set_body(self._h, body_row, *[_fp(a) for a in arrs]))

    def snapshot(self):
        """Device-side copy of the dynamic state (asynchronous)."""
        _lib.check(_lib.lib().nans_world_snapshot(self._h))

    def restore(self):
        """Back to the last snapshot (device-to-device, asynchronous)."""
        _lib.check(_lib.lib().nans_world_restore(self._h))

    # ---- stages -------------------------------------------------------------------------
    def integrate_forces(self, dt):
        _lib.check(_lib.lib().nans_integrate_forces(self._h, dt))

    def detect_collisions(self):
        _lib.check(_lib.lib().nans_detect_collisions(self._h))

    def solve_constraints(self, dt):
        _lib.check(_lib.lib().nans_solve_constraints(self._h, dt))

    def integrate_velocities(self, dt):
        _lib.check(_lib.lib().nans_integrate_velocities(self._h, dt))

    def rebuild_vertices(self):
        _lib.check(_lib.lib().nans_rebuild_vertices(self._h))

    def step(self, dt):
        _lib.check(_lib.lib().nans_step(self._h, dt))

    def models(self, device_ptr: int | None = None) -> np.ndarray | None:
        """Model matrices of every body (cubes, spheres, statics), column-major like glm::mat4: the "Model" uniform
        of the reference's draw section (code/nans.cpp:1870-1881, 1913-1941, 1971-1990) as instanced draw data.
        With ``device_ptr`` they are written there (a renderer's instance buffer) and nothing is copied back."""
        n = self.nb + self.n_statics
        if device_ptr is not None:
            _lib.check(_lib.lib().nans_world_models(self._h, C.c_void_p(device_ptr), None))
            return None
        out = np.zeros((n, 16), np.float32)
        _lib.check(_lib.lib().nans_world_models(self._h, None, _fp(out)))
        return out

    SOLVER_EXACT, SOLVER_SHUFFLED = 0, 1

    def set_solver(self, mode):
        """Sweep order of SolveConstraints: "exact" (the reference's list order, default) or "shuffled" (the same
        single Gauss-Seidel pass in a fixed pseudo-random order: ~10x shallower dependency graph, NOT the
        reference's results)."""
        m = {"exact": 0, "shuffled": 1}.get(mode, mode)
        _lib.check(_lib.lib().nans_world_set_solver(self._h, int(m)))

    STAGES = ("integrate_forces", "broadphase", "narrowphase", "contacts", "solver", "integrate_velocities", "step")

    def step_profiled(self, dt) -> dict:
        """One step with CUDA events between the stages (measurement only). ms per stage."""
        ms = np.zeros(8, np.float32)
        _lib.check(_lib.lib().nans_step_profiled(self._h, dt, _fp(ms)))
        return dict(zip(self.STAGES, ms[:7].tolist()))

    def synchronize(self):
        _lib.check(_lib.lib().nans_synchronize(self._h))

    # ---- results ------------------------------------------------------------------------
    def stats(self, strict=True) -> dict:
        st = _lib.StepStats()
        _lib.check(_lib.lib().nans_get_stats(self._h, C.byref(st)), allow_capacity=not strict)
        return {k: getattr(st, k) for k, _ in st._fields_ if k != "reserved"}

    def contacts(self) -> np.ndarray:
        n = C.c_int32(0)
        _lib.check(_lib.lib().nans_get_contacts(self._h, None, 0, C.byref(n)))
        out = np.zeros(max(n.value, 1), CONTACT_DTYPE)
        _lib.check(_lib.lib().nans_get_contacts(self._h, out.ctypes.data_as(C.c_void_p), len(out), C.byref(n)))
        return out[:n.value]

    def set_contacts(self, contacts: np.ndarray):
        c = np.ascontiguousarray(contacts)
        assert c.dtype.itemsize == CONTACT_DTYPE.itemsize
        _lib.check(_lib.lib().nans_set_contacts(self._h, c.ctypes.data_as(C.c_void_p), len(c)))

    def pairs(self):
        n = C.c_int32(0)
        _lib.check(_lib.lib().nans_get_pairs(self._h, None, None, 0, C.byref(n)))
        a = np.zeros(max(n.value, 1), np.int32); b = np.zeros(max(n.value, 1), np.int32)
        _lib.check(_lib.lib().nans_get_pairs(self._h, _ip(a), _ip(b), len(a), C.byref(n)))
        return a[:n.value], b[:n.value]


def check_collision(type_, pos_a, verts_a, rad_a, pos_b, verts_b, rad_b, device: int = 0) -> dict:
    """CheckCollision (code/nans.cpp:907-966) over n pairs on the GPU; host numpy in and out."""
    n = len(type_)
    t = np.ascontiguousarray(type_, np.int32)
    arrs = [np.ascontiguousarray(a, np.float32) for a in (pos_a, verts_a, rad_a, pos_b, verts_b, rad_b)]
    hit = np.zeros(n, np.int32); gjk = np.zeros(n, np.int32)
    N = np.zeros((n, 3), np.float32); PA = np.zeros((n, 3), np.float32); PB = np.zeros((n, 3), np.float32)
    _lib.check(_lib.lib().nans_check_collision_batch(n, _ip(t), _fp(arrs[0]), _fp(arrs[1]), _fp(arrs[2]),
                                                     _fp(arrs[3]), _fp(arrs[4]), _fp(arrs[5]),
                                                     _ip(hit), _ip(gjk), _fp(N), _fp(PA), _fp(PB), device))
    return dict(hit=hit, gjk=gjk, N=N, PA=PA, PB=PB)


def kernel_launches() -> int:
    return int(_lib.lib().nans_kernel_launches())
