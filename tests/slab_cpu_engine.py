"""CPU engine for the slab protocol model (tests only): the oracle restatement as the per-rank stepping
primitive, halo / velocity records as CPU torch tensors (tests/slab_protocol_model.py, over gloo)."""
import contextlib

import numpy as np
import torch

from nans_projekat_b200.slab import partition
from slab_protocol_model import HALO_FLOATS, VEL_FLOATS


class OracleEngine:
    device = torch.device("cpu")

    def __init__(self, O, scene, rank, world_size, ghost_cap=4096):
        self.O = O
        self.scene = scene
        self.lo, self.hi = partition(scene.n_cubes, world_size)[rank]
        self.n_owned = self.hi - self.lo
        self.ghost_cap = ghost_cap
        self.w = self._make_world(0)
        for f in ("pos", "vel", "force", "ang", "angvel", "torque", "scale"):
            getattr(self.w, f)[:] = getattr(scene, f)[self.lo:self.hi]
        self.w.mass[:] = scene.mass[self.lo:self.hi]
        self.w.moi[:] = scene.moi[self.lo:self.hi]
        self.w.verts[:] = scene.verts[self.lo:self.hi]
        self.n_ghosts = 0
        self.lists = {}
        self.contacts = None
        self.ghost_gid = np.zeros(0, np.int32)

    def _make_world(self, n_ghosts):
        w = self.O.World(self.n_owned + n_ghosts, 0, self.scene.n_statics)
        for f in ("st_pos", "st_ang", "st_scale", "st_mass", "st_moi", "st_verts"):
            getattr(w, f)[...] = getattr(self.scene, f)
        return w

    def stream_ctx(self):
        return contextlib.nullcontext()

    def _resize(self, n_ghosts):
        old = self.w
        w = self._make_world(n_ghosts)
        n = self.n_owned
        for f in ("pos", "vel", "force", "ang", "angvel", "torque", "scale", "mass", "moi", "verts"):
            getattr(w, f)[:n] = getattr(old, f)[:n]
        self.w, self.n_ghosts = w, n_ghosts

    def set_ghosts(self, n):
        self._resize(n)
        if n == 0:
            self.lists = {}

    def rebuild_vertices(self):
        self.w.rebuild_vertices()

    def integrate_forces(self, dt): self.w.integrate_forces(dt)

    def _aabbs(self):
        v = self.w.verts[:self.n_owned]
        lo, hi = v.min(1), v.max(1)
        m = np.float32(1e-3) + np.float32(1e-5) * np.maximum(np.abs(lo), np.abs(hi))
        return lo - m, hi + m

    def bounds(self):
        if self.n_owned == 0:
            return np.array([np.inf] * 3 + [-np.inf] * 3, np.float32)
        lo, hi = self._aabbs()
        return np.concatenate([lo.min(0), hi.max(0)]).astype(np.float32)

    def pack_halo(self, box, q):
        lo, hi = self._aabbs()
        sel = np.nonzero(((lo <= box[3:]) & (box[:3] <= hi)).all(1))[0]
        self.lists[q] = sel
        rec = np.zeros((len(sel), HALO_FLOATS), np.float32)
        w = self.w
        rec[:, 0:3] = w.pos[sel]; rec[:, 3] = w.mass[sel]
        rec[:, 4:7] = w.vel[sel]; rec[:, 8:11] = w.angvel[sel]; rec[:, 11] = w.moi[sel]
        rec[:, 12:36] = w.verts[sel].reshape(len(sel), 24)
        rec[:, 36] = (self.lo + sel).astype(np.int32).view(np.float32)
        return torch.from_numpy(rec)

    def set_ghost_rows(self, tensors):
        n = int(sum(len(t) for t in tensors))
        self._resize(n)
        if n:
            rec = torch.cat(tensors).numpy()
            w, o = self.w, self.n_owned
            w.pos[o:] = rec[:, 0:3]; w.mass[o:] = rec[:, 3]
            w.vel[o:] = rec[:, 4:7]; w.angvel[o:] = rec[:, 8:11]; w.moi[o:] = rec[:, 11]
            w.verts[o:] = rec[:, 12:36].reshape(n, 8, 3)
            self.ghost_gid = rec[:, 36].copy().view(np.int32)
        else:
            self.ghost_gid = np.zeros(0, np.int32)

    def detect(self):
        c = self.w.detect(prefilter=True)
        self.contacts = c[c["a"] < self.n_owned]     # a pair belongs to the owner of its lower-index body

    def unpack_owned_vel(self, q, buf):
        rows, b = self.lists[q], buf.numpy()
        self.w.vel[rows] = b[:, 0:3]
        self.w.angvel[rows] = b[:, 4:7]

    def solve(self, dt):
        self.w.solve(dt, self.contacts)

    def pack_ghost_vel(self, off, n):
        o = self.n_owned + off
        b = np.zeros((n, VEL_FLOATS), np.float32)
        b[:, 0:3] = self.w.vel[o:o + n]
        b[:, 4:7] = self.w.angvel[o:o + n]
        return torch.from_numpy(b)

    def integrate_velocities(self, dt):
        self.w.integrate_velocities(dt)
        self.w.rebuild_vertices()

    def contacts_global(self):
        c = self.contacts.copy()
        gid = np.concatenate([np.arange(self.lo, self.hi, dtype=np.int32), self.ghost_gid])
        c["a"] = gid[c["a"]]
        cc = c["type"] == 0
        c["b"][cc] = gid[c["b"][cc]]
        return c
