"""One world over several GPUs: slab decomposition by contiguous body-index ranges, NCCL halo exchange.

One process per GPU (``torch.distributed``).  Rank r owns global body rows ``[lo_r, hi_r)``; per step:

1. ``integrate_forces`` on the owned rows;
2. bounding box of the owned bodies -> ``all_gather`` (6 floats per rank);
3. every rank packs, for each LOWER rank q, its owned bodies whose AABB reaches into q's box
   (order-preserving, 160 B/body: pose, velocities, 8 vertices, global id); counts by ``all_to_all``,
   payload by grouped point-to-point send/recv; received bodies become ghost rows behind the owned
   ones (local row order == global index order);
4. broadphase + GJK/EPA + contact list, locally: a pair is emitted by the owner of its lower-index
   body, so every pair of the global world is tested exactly once;
5. the exact-order solve as a pipeline over ranks: receive the post-solve velocities of the own
   boundary bodies from the lower ranks, solve, send the ghosts' velocities (32 B/body) to their owners;
6. ``integrate_velocities`` + vertex rebuild on the owned rows.

The result is bit-identical to stepping the whole world on one GPU, because in the reference's sweep
order (contacts sorted by lower body index) every contact that a lower rank applies to a body precedes
every contact its owner applies.  Requirement: a body may be a ghost on at most ONE lower rank (true
for slabs thicker than a body, i.e. only neighbouring ranks touch).  The solve is latency-bound and
serial across ranks; everything else scales with 1/ranks.

The protocol (this file) is backend-agnostic: ``CudaEngine`` drives libnans_b200.so (device side in
``csrc/slab.cu``, buffers stay on the GPU, NCCL moves them); tests drive the same protocol with a CPU
engine over gloo.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .scenes import Scene

HALO_FLOATS = 40   # 10 float4 per body: pos, vel, angvel, 6 x verts, (global id, -, -, -)
VEL_FLOATS = 8     # 2 float4 per body: vel, angvel


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
    """A rank's share of one global world; the per-step exchange protocol.

    ``engine`` provides the local stepping primitives and (de)serialisation of halo / velocity records
    as torch tensors on ``engine.device``; ``dist`` is an initialised torch.distributed module."""

    def __init__(self, engine, rank: int, world_size: int, dist):
        self.e, self.rank, self.size, self.dist = engine, rank, world_size, dist
        self.n_ghosts = 0
        self.halo_bytes = 0
        self.lower_peers = 0

    def step(self, dt):
        import torch
        e, dist, R, r = self.e, self.dist, self.size, self.rank
        dev = e.device
        with e.stream_ctx():
            e.set_ghosts(0)
            e.integrate_forces(dt)
            # 2. bounds of every rank's owned bodies
            mine = torch.from_numpy(e.bounds()).to(dev)
            allb = torch.empty(R * 6, dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(allb, mine)
            allb = allb.cpu().numpy().reshape(R, 6)
            # 3. halos: my owned bodies that reach into a lower rank's box go to that rank
            halos = [e.pack_halo(allb[q], q) for q in range(r)]           # [cnt_q, 40] each, row order
            send_cnt = torch.tensor([len(h) for h in halos] + [0] * (R - r), dtype=torch.int64, device=dev)
            recv_cnt_t = torch.empty_like(send_cnt)
            dist.all_to_all_single(recv_cnt_t, send_cnt)
            recv_cnt = recv_cnt_t.cpu().numpy()
            ghosts = {p: torch.empty((int(recv_cnt[p]), HALO_FLOATS), dtype=torch.float32, device=dev)
                      for p in range(r + 1, R) if recv_cnt[p]}
            ops = [dist.P2POp(dist.isend, halos[q], q) for q in range(r) if len(halos[q])]
            ops += [dist.P2POp(dist.irecv, ghosts[p], p) for p in sorted(ghosts)]
            if ops:
                for req in dist.batch_isend_irecv(ops):
                    req.wait()
            order = sorted(ghosts)                                          # ascending rank = ascending global id
            e.set_ghost_rows([ghosts[p] for p in order])
            self.n_ghosts = int(sum(recv_cnt))
            self.halo_bytes = int(sum(len(h) for h in halos) + self.n_ghosts) * HALO_FLOATS * 4
            self.lower_peers = max(self.lower_peers, sum(1 for h in halos if len(h)))
            # 4. detection on owned + ghosts (pairs are emitted by the owner of the lower-index body)
            e.detect()
            # 5. exact-order solve, pipelined over ranks
            for q in range(r):                                              # ascending: the sweep order
                if len(halos[q]):
                    buf = torch.empty((len(halos[q]), VEL_FLOATS), dtype=torch.float32, device=dev)
                    dist.recv(buf, q)
                    e.unpack_owned_vel(q, buf)
            e.solve(dt)
            off = 0
            for p in order:
                n = int(recv_cnt[p])
                dist.send(e.pack_ghost_vel(off, n), p)
                off += n
            # 6. positions, angles, vertices
            e.integrate_velocities(dt)


class CudaEngine:
    """libnans_b200.so as the slab engine: everything stays in device memory."""

    def __init__(self, scene: Scene, rank: int, world_size: int, device: int, ghost_frac: float = 1.25):
        import torch
        from . import _lib
        from .world import World
        self._lib, self.torch = _lib, torch
        self.ranges = partition(scene.n_cubes, world_size)
        self.lo, self.hi = self.ranges[rank]
        self.n_owned = self.hi - self.lo
        # ghosts are selected by the lower rank's bounding box, so one stray body can pull in a whole
        # extra layer: leave room for more than a slab's worth
        self.ghost_cap = int(max(1024, ghost_frac * max(self.n_owned, 1)))
        self.device = torch.device("cuda", device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.world = World(local_scene(scene, self.lo, self.hi, self.ghost_cap), device=device,
                           stream=self.stream.cuda_stream)
        self.L = _lib.lib()
        self.n_ghosts = 0
        self.send_buf = {}
        self.list_off = {}
        self.ghost_ids = None
        self.set_ghosts(0)

    def stream_ctx(self):
        return self.torch.cuda.stream(self.stream)

    def _ck(self, rc):
        self._lib.check(rc)

    def set_ghosts(self, n):
        self.n_ghosts = n
        self._ck(self.L.nans_world_set_partition(self.world._h, self.n_owned, n))
        if n == 0:
            self._next_off = 0

    def rebuild_vertices(self):
        self.set_ghosts(0)
        self.world.rebuild_vertices()

    def integrate_forces(self, dt): self.world.integrate_forces(dt)
    def detect(self): self.world.detect_collisions()
    def solve(self, dt): self.world.solve_constraints(dt)
    def integrate_velocities(self, dt): self.world.integrate_velocities(dt)

    def bounds(self) -> np.ndarray:
        from .world import _fp
        b = np.zeros(6, np.float32)
        if self.n_owned == 0:
            b[:3], b[3:] = np.inf, -np.inf
            return b
        self._ck(self.L.nans_world_bounds(self.world._h, _fp(b)))
        return b

    def pack_halo(self, box, q):
        from .world import _fp
        torch = self.torch
        if q not in self.send_buf:
            self.send_buf[q] = torch.empty((self.ghost_cap, HALO_FLOATS), dtype=torch.float32, device=self.device)
        cnt = C.c_int32(0)
        box = np.ascontiguousarray(box, np.float32)
        self.list_off[q] = self._next_off
        self._ck(self.L.nans_slab_pack_halo(self.world._h, _fp(box), self.lo, self.send_buf[q].data_ptr(),
                                            self.ghost_cap, self._next_off, C.byref(cnt)))
        self._next_off += cnt.value
        return self.send_buf[q][:cnt.value]

    def set_ghost_rows(self, tensors):
        n = int(sum(len(t) for t in tensors))
        if n > self.ghost_cap:
            raise self._lib.NansError(f"{n} ghosts exceed the ghost capacity {self.ghost_cap}")
        row = self.n_owned
        for t in tensors:
            self._ck(self.L.nans_slab_unpack_halo(self.world._h, t.data_ptr(), len(t), row))
            row += len(t)
        self.ghost_ids = (self.torch.cat([t[:, 36] for t in tensors]).contiguous().view(self.torch.int32)
                          if tensors else None)
        self._keep = tensors            # the unpack kernels read them asynchronously
        self.n_ghosts = n
        self._ck(self.L.nans_world_set_partition(self.world._h, self.n_owned, n))

    def unpack_owned_vel(self, q, buf):
        self._ck(self.L.nans_slab_unpack_owned_vel(self.world._h, self.list_off[q], len(buf), buf.data_ptr()))
        self._keep_vel = buf

    def pack_ghost_vel(self, ghost_off, n):
        out = self.torch.empty((n, VEL_FLOATS), dtype=self.torch.float32, device=self.device)
        self._ck(self.L.nans_slab_pack_ghost_vel(self.world._h, self.n_owned + ghost_off, n, out.data_ptr()))
        return out

    # -- results ----------------------------------------------------------------------------------
    def download_owned(self, fields=("pos", "vel", "ang", "angvel", "verts")) -> Scene:
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
            gid[self.n_owned:] = self.ghost_ids.cpu().numpy()
        c["a"] = gid[c["a"]]
        cc = c["type"] == 0
        c["b"][cc] = gid[c["b"][cc]]
        return c

    def close(self):
        self.world.close()
