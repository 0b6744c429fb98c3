"""ctypes binding of libnans_b200.so (include/nans_b200.h).

There is no CPU fallback: if the CUDA library is missing this import raises, and every call
fails loudly when no CUDA device is usable.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libnans_b200.so")

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int32)


class WorldDesc(C.Structure):
    _fields_ = [("n_cubes", C.c_int32), ("n_spheres", C.c_int32), ("n_statics", C.c_int32),
                ("max_pairs", C.c_int32), ("max_contacts", C.c_int32), ("device", C.c_int32),
                ("arena", C.c_void_p), ("arena_bytes", C.c_uint64), ("stream", C.c_void_p)]


class SceneView(C.Structure):
    _fields_ = [(n, f32p) for n in ("pos", "vel", "force", "ang", "angvel", "torque", "mass", "moi",
                                     "scale", "radius", "verts", "st_pos", "st_ang", "st_scale",
                                     "st_mass", "st_moi", "st_verts")] + [("world_id", i32p)]


class StepStats(C.Structure):
    _fields_ = [("n_pairs", C.c_int32), ("n_contacts", C.c_int32), ("n_gjk_found", C.c_int32),
                ("solver_levels", C.c_int32), ("overflow", C.c_int32), ("max_epa_faces", C.c_int32),
                ("accum_fallbacks", C.c_int32), ("reserved", C.c_int32 * 1)]


# every symbol include/nans_b200.h declares
EXPORTS = ("nans_world_create", "nans_world_destroy", "nans_world_arena_bytes", "nans_last_error",
           "nans_device_count", "nans_world_upload", "nans_world_download", "nans_world_add_force",
           "nans_world_set_body", "nans_world_upload_async", "nans_world_download_async", "nans_world_wait", "nans_world_snapshot", "nans_world_restore", "nans_integrate_forces", "nans_detect_collisions",
           "nans_solve_constraints", "nans_integrate_velocities", "nans_rebuild_vertices", "nans_step", "nans_step_profiled",
           "nans_synchronize", "nans_get_stats", "nans_get_contacts", "nans_get_pairs", "nans_set_contacts",
           "nans_check_collision_batch", "nans_check_collision_device", "nans_check_collision_device_status", "nans_kernel_launches", "nans_debug_solver_trace", "nans_debug_scan", "nans_world_set_solver", "nans_world_models",
           "nans_slab_unique_id", "nans_slab_init", "nans_slab_ipc_handle", "nans_slab_connect", "nans_slab_step",
           "nans_slab_status", "nans_slab_row_gids")

_lib = None


class NansError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NansError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        f"(nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    for name in EXPORTS:
        getattr(L, name)  # raises AttributeError if a declared symbol is not exported
    L.nans_last_error.restype = C.c_char_p
    L.nans_world_arena_bytes.restype = C.c_uint64
    L.nans_world_arena_bytes.argtypes = [C.POINTER(WorldDesc)]
    L.nans_kernel_launches.restype = C.c_uint64
    L.nans_world_create.argtypes = [C.POINTER(WorldDesc), C.POINTER(C.c_void_p)]
    L.nans_world_destroy.argtypes = [C.c_void_p]
    L.nans_world_destroy.restype = None
    for name in ("nans_world_upload", "nans_world_download"):
        getattr(L, name).argtypes = [C.c_void_p, C.POINTER(SceneView)]
    L.nans_world_upload_async.argtypes = [C.c_void_p, C.POINTER(SceneView)]
    L.nans_world_download_async.argtypes = [C.c_void_p, C.POINTER(SceneView), i32p]
    L.nans_world_wait.argtypes = [C.c_void_p, C.c_int32]
    L.nans_world_add_force.argtypes = [C.c_void_p, C.c_int32, f32p, f32p]
    L.nans_world_set_body.argtypes = [C.c_void_p, C.c_int32, f32p, f32p, f32p]
    for name in ("nans_integrate_forces", "nans_solve_constraints", "nans_integrate_velocities", "nans_step"):
        getattr(L, name).argtypes = [C.c_void_p, C.c_float]
    for name in ("nans_detect_collisions", "nans_rebuild_vertices", "nans_synchronize", "nans_world_snapshot",
                 "nans_world_restore"):
        getattr(L, name).argtypes = [C.c_void_p]
    L.nans_step_profiled.argtypes = [C.c_void_p, C.c_float, f32p]
    L.nans_get_stats.argtypes = [C.c_void_p, C.POINTER(StepStats)]
    L.nans_get_contacts.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, i32p]
    L.nans_get_pairs.argtypes = [C.c_void_p, i32p, i32p, C.c_int32, i32p]
    L.nans_set_contacts.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    L.nans_check_collision_batch.argtypes = [C.c_int32, i32p, f32p, f32p, f32p, f32p, f32p, f32p,
                                             i32p, i32p, f32p, f32p, f32p, C.c_int32]
    L.nans_check_collision_device.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.nans_check_collision_device_status.argtypes = [i32p]
    L.nans_world_set_solver.argtypes = [C.c_void_p, C.c_int32]
    L.nans_world_models.argtypes = [C.c_void_p, C.c_void_p, f32p]
    L.nans_slab_unique_id.argtypes = [C.c_void_p]
    L.nans_slab_init.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
    L.nans_slab_ipc_handle.argtypes = [C.c_void_p, C.c_void_p]
    L.nans_slab_connect.argtypes = [C.c_void_p, C.c_void_p]
    L.nans_slab_step.argtypes = [C.c_void_p, C.c_float]
    L.nans_slab_status.argtypes = [C.c_void_p, i32p, i32p, C.POINTER(C.c_int64)]
    L.nans_slab_row_gids.argtypes = [C.c_void_p, i32p, C.c_int32, i32p]
    L.nans_debug_solver_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
    L.nans_debug_scan.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    _lib = L
    return L


def check(rc: int, allow_capacity: bool = False):
    if rc == 0:
        return
    msg = lib().nans_last_error().decode(errors="replace")
    if rc == -3 and allow_capacity:
        return
    raise NansError(f"libnans_b200 error {rc}: {msg}")
