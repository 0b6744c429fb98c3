"""One world over several GPUs: slab decomposition by contiguous body-index ranges, NCCL halo exchange.

One process per GPU (``torch.distributed``, backend ``nccl``).  Rank r owns global body rows
``[lo_r, hi_r)``; per step (device side in ``csrc/slab.cu``, C ABI ``nans_slab_*``):

1. ``integrate_forces`` on the owned rows;
2. bounding box of the owned bodies -> ``all_gather`` (6 floats per rank);
3. every rank packs, for each LOWER rank q, its owned bodies whose AABB reaches into q's box
   (order-preserving, 160 B/body: pose, velocities, 8 vertices, global id); counts by ``all_to_all``,
   payload by grouped NCCL send/recv; received bodies become ghost rows behind the owned ones
   (local row order == global index order);
4. broadphase + GJK/EPA + contact list, locally: a pair is emitted by the owner of its lower-index
   body, so every pair of the global world is tested exactly once;
5. the exact-order solve as a pipeline over ranks: receive the post-solve velocities of the own
   boundary bodies from the lower ranks, solve, send the ghosts' velocities (32 B/body) to their owners;
6. ``integrate_velocities`` + vertex rebuild on the owned rows.

The result is bit-identical to stepping the whole world on one GPU (tools/slab_check.py), because in
the reference's sweep order every contact that a lower rank applies to a body precedes every contact
its owner applies.  Requirement: a body may be a ghost on at most ONE lower rank (true for slabs
thicker than a body, i.e. only neighbouring ranks touch); violations are counted in ``multi_ghost``.
The solve is latency-bound and serial across ranks; everything else scales with 1/ranks.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .scenes import Scene
from .world import World, _fp

HALO_FLOATS = 40   # 10 float4 per body
VEL_FLOATS = 8     # 2 float4 per body


def partition(n: int, world_size: int):
    """Contiguous, near-equal index ranges [lo, hi) per rank."""
    per = (n + world_size - 1) // world_size
    return [(min(r * per, n), min((r + 1) * per, n)) for r in range(world_size)]


def local_scene(scene: Scene, lo: int, hi: int, ghost_cap: int) -> Scene:
    """The rank's slice of a cube-only scene, with room for ghost rows behind the owned ones."""
    assert scene.n_spheres == 0, "slab mode supports cube-only worlds"
    n_owned = hi - lo
    s = Scene(n_owned + ghost_cap, 0, scene.n_statics)
    for f in Scene.VEC_FIELDS + ("mass", "moi", "radius"):
        getattr(s, f)[:n_owned] = getattr(scene, f)[lo:hi]
    s.verts[:n_owned] = scene.verts[lo:hi]
    # dummy ghost rows (overwritten by the halo unpack before they are ever used)
    s.pos[n_owned:] = (0.0, -1.0e6, 0.0)
    s.mass[n_owned:] = 1.0
    s.moi[n_owned:] = 1.0
    for f in ("st_pos", "st_ang", "st_scale", "st_mass", "st_moi", "st_verts"):
        getattr(s, f)[...] = getattr(scene, f)
    return s


class SlabWorld:
    """A rank's share of one global world.  Construct on every rank with the same global scene."""

    def __init__(self, scene: Scene, rank: int, world_size: int, device: int, ghost_frac: float = 0.6,
                 stream=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.size = rank, world_size
        self.ranges = partition(scene.n_cubes, world_size)
        self.lo, self.hi = self.ranges[rank]
        self.n_owned = self.hi - self.lo
        self.ghost_cap = int(max(1024, ghost_frac * max(self.n_owned, 1)))
        self.dev = torch.device("cuda", device)
        self.stream = stream or torch.cuda.Stream(device=self.dev)
        self.world = World(local_scene(scene, self.lo, self.hi, self.ghost_cap), device=device,
                           stream=self.stream.cuda_stream)
        self._L = _lib.lib()
        self._set_partition(self.n_owned, 0)
        f32 = torch.float32
        self.send_buf = [torch.empty(self.ghost_cap * HALO_FLOATS, dtype=f32, device=self.dev) if q < rank else None
                         for q in range(world_size)]
        self.ghost_buf = torch.empty(self.ghost_cap * HALO_FLOATS, dtype=f32, device=self.dev)
        self.vel_out = torch.empty(self.ghost_cap * VEL_FLOATS, dtype=f32, device=self.dev)
        self.vel_in = torch.empty(self.ghost_cap * VEL_FLOATS, dtype=f32, device=self.dev)
        self.n_ghosts = 0
        self.multi_ghost = 0
        self.halo_bytes = 0

    # -- thin wrappers ------------------------------------------------------------------------
    def _set_partition(self, n_owned, n_ghosts):
        _lib.check(self._L.nans_world_set_partition(self.world._h, n_owned, n_ghosts))

    def bounds(self) -> np.ndarray:
        b = np.zeros(6, np.float32)
        _lib.check(self._L.nans_world_bounds(self.world._h, _fp(b)))
        return b

    def rebuild_vertices(self):
        self._set_partition(self.n_owned, 0)
        self.world.rebuild_vertices()

    # -- one step -------------------------------------------------------------------------------
    def step(self, dt):
        torch, dist = self.torch, self.dist
        w, L, R, r = self.world, self._L, self.size, self.rank
        with torch.cuda.stream(self.stream):
            self._set_partition(self.n_owned, 0)
            w.integrate_forces(dt)
            # 2. bounds of every rank's owned bodies
            mine = torch.from_numpy(self.bounds()).to(self.dev)
            allb = torch.empty(R * 6, dtype=torch.float32, device=self.dev)
            dist.all_gather_into_tensor(allb, mine)
            allb = allb.cpu().numpy().reshape(R, 6)
            # 3. halos: my owned bodies that reach into a lower rank's box go to that rank
            send_cnt = np.zeros(R, np.int64)
            list_off = np.zeros(R + 1, np.int64)
            for q in range(r):
                cnt = C.c_int32(0)
                box = np.ascontiguousarray(allb[q], np.float32)
                _lib.check(L.nans_slab_pack_halo(w._h, _fp(box), self.lo, self.send_buf[q].data_ptr(),
                                                 self.ghost_cap, int(list_off[q]), C.byref(cnt)))
                send_cnt[q] = cnt.value
                list_off[q + 1] = list_off[q] + cnt.value
            sc = torch.from_numpy(send_cnt).to(self.dev)
            rc = torch.empty_like(sc)
            dist.all_to_all_single(rc, sc)
            recv_cnt = rc.cpu().numpy()
            if int(recv_cnt.sum()) > self.ghost_cap:
                raise _lib.NansError(f"rank {r}: {int(recv_cnt.sum())} ghosts exceed the ghost capacity {self.ghost_cap}")
            ops, goff = [], np.zeros(R + 1, np.int64)
            for p in range(R):
                goff[p + 1] = goff[p] + recv_cnt[p]
            for q in range(r):
                if send_cnt[q]:
                    ops.append(dist.P2POp(dist.isend, self.send_buf[q][:int(send_cnt[q]) * HALO_FLOATS], q))
            for p in range(r + 1, R):
                if recv_cnt[p]:
                    ops.append(dist.P2POp(dist.irecv, self.ghost_buf[int(goff[p]) * HALO_FLOATS:
                                                                     int(goff[p + 1]) * HALO_FLOATS], p))
            if ops:
                for req in dist.batch_isend_irecv(ops):
                    req.wait()
            self.n_ghosts = int(goff[R])
            self.halo_bytes = int(send_cnt.sum() + recv_cnt.sum()) * HALO_FLOATS * 4
            if self.n_ghosts:
                _lib.check(L.nans_slab_unpack_halo(w._h, self.ghost_buf.data_ptr(), self.n_ghosts, self.n_owned))
            self._set_partition(self.n_owned, self.n_ghosts)
            # 4. detection on owned + ghosts (pairs are emitted by the owner of the lower-index body)
            w.detect_collisions()
            # 5. exact-order solve, pipelined over ranks
            for q in range(r):                      # ascending: the reference's sweep order
                n = int(send_cnt[q])
                if n:
                    dist.recv(self.vel_in[:n * VEL_FLOATS], q)
                    _lib.check(L.nans_slab_unpack_owned_vel(w._h, int(list_off[q]), n, self.vel_in.data_ptr()))
            w.solve_constraints(dt)
            for p in range(r + 1, R):
                n = int(recv_cnt[p])
                if n:
                    _lib.check(L.nans_slab_pack_ghost_vel(w._h, self.n_owned + int(goff[p]), n, self.vel_out.data_ptr()))
                    dist.send(self.vel_out[:n * VEL_FLOATS], p)
            # 6. positions, angles, vertices
            w.integrate_velocities(dt)
            # a body sent to more than one lower rank would need its velocity forwarded between them
            self.multi_ghost += int((send_cnt > 0).sum() > 1)

    # -- results ----------------------------------------------------------------------------------
    def download_owned(self, fields=("pos", "vel", "ang", "angvel", "verts")) -> Scene:
        self._set_partition(self.n_owned, self.n_ghosts)
        s = self.world.download(fields=fields)
        out = Scene(self.n_owned, 0, s.n_statics)
        for f in fields:
            getattr(out, f)[...] = getattr(s, f)[:self.n_owned]
        return out

    def contacts_global(self) -> np.ndarray:
        """This rank's contacts with GLOBAL body ids (CC: a, b cubes; CF: a cube, b static)."""
        c = self.world.contacts().copy()
        gid = np.arange(self.lo, self.lo + self.n_owned + self.n_ghosts, dtype=np.int32)
        if self.n_ghosts:
            # ghost rows carry their global id in the halo record (first float of the 10th float4)
            g = self.ghost_buf[:self.n_ghosts * HALO_FLOATS].view(-1, HALO_FLOATS)[:, 36].contiguous()
            gid[self.n_owned:] = g.view(self.torch.int32).cpu().numpy()
        c["a"] = gid[c["a"]]
        cc = c["type"] == 0
        c["b"][cc] = gid[c["b"][cc]]
        return c

    def close(self):
        self.world.close()
