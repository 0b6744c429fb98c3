"""One world over several GPUs: the Python host of csrc/slab.cu (include/nans_b200.h, ``nans_slab_*``).

One process per GPU.  Everything that happens per step -- the NCCL all-gather of the slab boxes, the
fixed-capacity halo send/recv, detection, the cross-GPU dataflow solve over NVLink peer memory -- is queued by
``nans_slab_step`` on the world's stream inside libnans_b200.so; this file only does the one-time rendezvous
(the NCCL unique id and the CUDA IPC blobs travel through ``torch.distributed``) and the numpy marshalling.
The decomposed world is bit-identical to the same world on one GPU (tests/test_slab_gpu.py, tools/slab_check.py);
the ordering argument behind that is modelled on the CPU in tests/slab_protocol_model.py.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .scenes import Scene
from .world import World, _ip


def partition(n: int, world_size: int):
    """Contiguous, near-equal index ranges [lo, hi) per rank."""
    per = (n + world_size - 1) // world_size
    return [(min(r * per, n), min((r + 1) * per, n)) for r in range(world_size)]


def local_scene(scene: Scene, lo: int, hi: int, ghost_cap: int, capacity: int | None = None) -> Scene:
    """Rows [lo, hi) of a cube-only scene, padded with ghost rows up to ``capacity`` (default: owned + ghost_cap)."""
    assert scene.n_spheres == 0, "slab mode supports cube-only worlds"
    n_owned = hi - lo
    cap = capacity if capacity is not None else n_owned + ghost_cap
    assert cap >= n_owned + ghost_cap
    s = Scene(cap, 0, scene.n_statics)
    for f in Scene.VEC_FIELDS + ("mass", "moi", "radius"):
        getattr(s, f)[:n_owned] = getattr(scene, f)[lo:hi]
    s.verts[:n_owned] = scene.verts[lo:hi]
    # ghost rows: overwritten by the halo unpack before they are ever live
    s.pos[n_owned:] = (0.0, -1.0e6, 0.0)
    s.mass[n_owned:] = 1.0
    s.moi[n_owned:] = 1.0
    for f in ("st_pos", "st_ang", "st_scale", "st_mass", "st_moi", "st_verts"):
        getattr(s, f)[...] = getattr(scene, f)
    return s


class SlabWorld:
    """This rank's share of one global world.  ``owned`` holds the rank's own bodies (global indices
    ``[gid_base, gid_base + owned.n_cubes)``); ``capacity`` rows are allocated on every rank (it must be the
    same everywhere: the ranks address each other's arenas by offset), ``halo_cap`` of them receive ghosts."""

    def __init__(self, owned: Scene, rank: int, world_size: int, dist, device: int, gid_base: int, halo_cap: int,
                 capacity: int, stream: int | None = None, max_pairs: int = 0, max_contacts: int = 0):
        self.rank, self.size, self.n_owned, self.gid_base, self.halo_cap = rank, world_size, owned.n_cubes, gid_base, halo_cap
        L = self.L = _lib.lib()
        self.world = World(local_scene(owned, 0, owned.n_cubes, halo_cap, capacity), device=device, stream=stream,
                           max_pairs=max_pairs, max_contacts=max_contacts)
        h = self.world._h
        uid = [bytes(128)]
        if rank == 0:
            buf = C.create_string_buffer(128)
            _lib.check(L.nans_slab_unique_id(buf))
            uid = [buf.raw]
        dist.broadcast_object_list(uid, src=0)
        _lib.check(L.nans_slab_init(h, rank, world_size, C.c_char_p(uid[0]), self.n_owned, gid_base, halo_cap))
        blob = C.create_string_buffer(128)
        _lib.check(L.nans_slab_ipc_handle(h, blob))
        blobs = [None] * world_size
        dist.all_gather_object(blobs, blob.raw)
        _lib.check(L.nans_slab_connect(h, C.c_char_p(b"".join(blobs))))
        dist.barrier()

    def rebuild_vertices(self):
        self.world.rebuild_vertices()

    def step(self, dt):
        _lib.check(self.L.nans_slab_step(self.world._h, dt))

    def status(self) -> dict:
        """Synchronises; raises if the exchange pattern was ever violated (the world would no longer be exact)."""
        err, live, hb = C.c_int32(0), C.c_int32(0), C.c_int64(0)
        _lib.check(self.L.nans_slab_status(self.world._h, C.byref(err), C.byref(live), C.byref(hb)))
        return {"live_rows": live.value, "ghosts": live.value - self.n_owned, "halo_message_bytes": hb.value}

    def download_owned(self, fields=("pos", "vel", "ang", "angvel", "verts")) -> Scene:
        s = self.world.download(fields=fields)
        out = Scene(self.n_owned, 0, s.n_statics)
        for f in fields:
            getattr(out, f)[...] = getattr(s, f)[:self.n_owned]
        return out

    def row_gids(self) -> np.ndarray:
        n = C.c_int32(0)
        _lib.check(self.L.nans_slab_row_gids(self.world._h, None, 0, C.byref(n)))
        out = np.zeros(max(n.value, 1), np.int32)
        _lib.check(self.L.nans_slab_row_gids(self.world._h, _ip(out), len(out), C.byref(n)))
        return out[:n.value]

    def contacts_global(self) -> np.ndarray:
        """This rank's contacts with GLOBAL body ids (CC: a, b cubes; CF: a cube, b static)."""
        c = self.world.contacts().copy()
        gid = self.row_gids()
        c["a"] = gid[c["a"]]
        cc = c["type"] == 0
        c["b"][cc] = gid[c["b"][cc]]
        return c

    def close(self):
        self.world.close()
