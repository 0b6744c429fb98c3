"""The ordering argument of the multi-GPU slab decomposition (csrc/slab.cu) on CPU: world_size-2, -3
and -4 gloo process groups, the oracle as the per-rank engine (tests/slab_protocol_model.py).  The decomposed
world must stay BIT-IDENTICAL to the same world stepped as one piece.  Scene: the slab-major numbered pile the
GPU path uses (index ranges = spatial x-slabs)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, size, port, steps, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from oracle import oracle as O
    from nans_projekat_b200 import scenes
    from slab_protocol_model import SlabProtocolModel as SlabWorld
    from slab_cpu_engine import OracleEngine
    from helpers import world_from_scene
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=size)
    scene = scenes.cube_pile_slabs(n_slabs=size, side_x=3, ny=6, nz=5, seed=3, jitter=0.02)
    whole = world_from_scene(O, scene); whole.rebuild_vertices()
    scene.verts[:] = whole.verts; scene.st_verts[:] = whole.st_verts
    eng = OracleEngine(O, scene, rank, size)
    sw = SlabWorld(eng, rank, size, dist)
    dt = np.float32(1 / 60.)
    bad, max_ghosts, ncontacts = 0, 0, 0
    for k in range(steps):
        sw.step(dt)
        ref_c = whole.step(dt, prefilter=True)
        st = {f: getattr(eng.w, f)[:eng.n_owned] for f in ("pos", "vel", "ang", "angvel", "verts")}
        for f, a in st.items():
            if not np.array_equal(a.view(np.uint32), getattr(whole, f)[eng.lo:eng.hi].view(np.uint32)):
                bad += 1
        parts = [None] * size
        dist.all_gather_object(parts, eng.contacts_global())
        allc = np.concatenate(parts)
        allc = allc[np.lexsort((allc["b"], allc["a"], allc["type"] != 0))]
        if allc.tobytes() != ref_c.tobytes():
            bad += 1
        max_ghosts = max(max_ghosts, sw.n_ghosts)
        ncontacts = max(ncontacts, len(ref_c))
    q.put((rank, bad, max_ghosts, ncontacts, sw.lower_peers))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("size", [2, 3, 4])
def test_slab_protocol_is_exact(oracle, size):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + size + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, size, port, 45, q)) for r in range(size)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] == 0 for r in res), f"slab world diverged from the single world: {res}"
    assert max(r[2] for r in res) > 0, "no ghosts were ever exchanged"
    assert max(r[3] for r in res) > 100, "scene never developed contacts"
    assert max(r[4] for r in res) <= 1, "a body was a ghost on more than one lower rank"


def test_partition_and_local_scene():
    from nans_projekat_b200 import scenes
    from nans_projekat_b200.slab import partition, local_scene
    assert partition(10, 3) == [(0, 4), (4, 8), (8, 10)]
    assert partition(5, 8)[-1] == (5, 5)
    s = scenes.cube_pile(n_side=4, layers=4)
    ls = local_scene(s, 16, 32, 8)
    assert ls.n_cubes == 24 and np.array_equal(ls.pos[:16], s.pos[16:32]) and (ls.pos[16:, 1] < -1e5).all()
    assert local_scene(s, 16, 32, 8, capacity=40).n_cubes == 40


def test_slab_major_pile_numbering():
    """cube_pile_slabs: contiguous index ranges are x-slabs, and a rank can build its own slab alone."""
    from nans_projekat_b200 import scenes
    full = scenes.cube_pile_slabs(n_slabs=3, side_x=4, ny=3, nz=5, seed=11)
    m = 4 * 3 * 5
    assert full.n_cubes == 3 * m
    for r in range(3):
        part = scenes.cube_pile_slabs(n_slabs=3, side_x=4, ny=3, nz=5, seed=11, slab=r)
        assert np.array_equal(part.pos, full.pos[r * m:(r + 1) * m])
        assert np.array_equal(part.st_pos, full.st_pos) and np.array_equal(part.st_scale, full.st_scale)
        x = full.pos[r * m:(r + 1) * m, 0]
        assert x.min() > 0.5 + (4 * r - 0.5) * 1.02 and x.max() < 0.5 + (4 * r + 3.5) * 1.02
