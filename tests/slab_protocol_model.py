"""TEST INFRASTRUCTURE: a CPU model of the ordering argument behind csrc/slab.cu.

The product path (nans_slab_step in libnans_b200.so) exchanges halos with NCCL and hands ghost velocities to their
owner by peer stores from inside the solve kernel.  What makes that EXACT is an ordering property of the reference's
sweep (contacts sorted by lower body index, code/nans.cpp:1355-1396, 1539-1548): with index-range ownership and
ghosts appended behind the owned rows, every contact a lower rank applies to a body precedes every contact the body's
owner applies.  This model runs that property with the CPU oracle as the per-rank engine over a gloo process group:
halo selection against the lower ranks' boxes, pairs owned by the lower-index body's owner, the solve as a pipeline
over ranks (receive the boundary bodies' post-solve velocities, solve, pass the ghosts' on).  The decomposed world
must stay bit-identical to the world stepped as one piece (tests/test_slab_cpu.py)."""
from __future__ import annotations

import numpy as np

HALO_FLOATS = 40   # 10 float4 per body: pos, vel, angvel, 6 x verts, (global id, -, -, -)
VEL_FLOATS = 8     # 2 float4 per body: vel, angvel


class SlabProtocolModel:
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


