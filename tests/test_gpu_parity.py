"""GPU parity tests: the CUDA path, called through the C ABI (libnans_b200.so), against the CPU
oracle on the same inputs and against golden vectors produced by the reference's own nans.so.

Bars (BASELINE.json north_star): GJK flags and pair/contact sets bit-exact; EPA normal/depth and
per-step body state within 1e-4 relative.  Because every kernel reproduces the reference's fp32
operation order (and glibc's sinf/cosf), these tests assert the stronger property: BIT-EXACT floats.
Steps always start from the oracle's/golden state (trajectories diverge chaotically otherwise).
"""
import os

import numpy as np
import pytest

from helpers import assert_bit_equal, load_stage_records, rel_err, world_from_scene, world_from_stage_record

pytestmark = pytest.mark.gpu

STATE = ("pos", "vel", "force", "ang", "angvel", "torque", "verts")
DT = np.float32(1 / 60.)


@pytest.fixture(scope="module")
def nb200():
    from nans_projekat_b200 import _lib, world
    assert _lib.lib().nans_device_count() > 0, "no CUDA device: the product path has no CPU fallback"
    return world


def scene_from_oracle_world(w):
    from nans_projekat_b200.scenes import Scene
    s = Scene(w.n_cubes, w.n_spheres, w.n_statics)
    for f in s.ARRAYS:
        getattr(s, f)[...] = getattr(w, f)
    return s


# ------------------------------------------------------------------------------------ narrowphase
@pytest.mark.parametrize("tag", ["rot", "axis"])
def test_narrowphase_golden(nb200, golden_dir, tag):
    """CheckCollision against outputs of the reference binary (all five pair types)."""
    z = np.load(os.path.join(golden_dir, "narrowphase.npz"))
    g = lambda k: z[f"{tag}_{k}"]
    r = nb200.check_collision(g("type"), g("pos_a"), g("verts_a"), g("rad_a"), g("pos_b"), g("verts_b"), g("rad_b"))
    assert np.array_equal(r["hit"], g("ref_hit")), "hit flags must be bit-exact"
    h = g("ref_hit") == 1
    for k in ("N", "PA", "PB"):
        assert_bit_equal(r[k][h], g(f"ref_{k}")[h], f"{tag} {k}")
    depth_ref = np.einsum("ij,ij->i", g("ref_PA")[h] - g("ref_PB")[h], g("ref_N")[h])
    depth = np.einsum("ij,ij->i", r["PA"][h] - r["PB"][h], r["N"][h])
    assert rel_err(depth, depth_ref) <= 1e-4 and rel_err(r["N"][h], g("ref_N")[h]) <= 1e-4


def test_narrowphase_emptied_polytope_golden(nb200, golden_dir):
    """EPA runs that empty the reference's triangle list (it then reads the stale Triangle[0], code/nans.cpp:807-866;
    found by the 1 M-cube parity test of round 2): flags and contacts as nans.so reports them, through both the
    stand-alone batch path and (below, test_headline_gpu.py) the world kernel."""
    z = np.load(os.path.join(golden_dir, "epa_emptied.npz"))
    r = nb200.check_collision(z["type"], z["pos_a"], z["verts_a"], z["rad_a"], z["pos_b"], z["verts_b"], z["rad_b"])
    assert np.array_equal(r["hit"], z["ref_hit"])
    h = z["ref_hit"] == 1
    for k in ("N", "PA", "PB"):
        assert_bit_equal(r[k][h], z[f"ref_{k}"][h], f"emptied {k}")


@pytest.mark.parametrize("rotated,n", [(True, 1 << 18), (False, 1 << 17)])
def test_narrowphase_vs_oracle(nb200, oracle, rotated, n):
    """Config C3 at reduced size (CC/CS/SS strata 8:7:1): flags, GJK results and contacts bit-exact."""
    from nans_projekat_b200 import scenes
    p = scenes.narrowphase_pairs(n, seed=1234, rotated=rotated)
    args = (p["type"], p["pos_a"], p["verts_a"], p["rad_a"], p["pos_b"], p["verts_b"], p["rad_b"])
    g = nb200.check_collision(*args)
    o = oracle.check_collision_batch(*args)
    assert np.array_equal(g["gjk"], o["gjk"]), "GJK evolve_result flags"
    assert np.array_equal(g["hit"], o["hit"]), "hit flags"
    h = o["hit"] == 1
    assert 0.3 < h.mean() < 0.7
    for k in ("N", "PA", "PB"):
        assert_bit_equal(g[k][h], o[k][h], k)
    assert g["hit"][p["type"] == 3].sum() == 0   # SS never hits in the reference


def test_narrowphase_mixed_types_and_value_equal_vertices(nb200, oracle):
    """All five pair types shuffled (every warp of the GJK kernel feeds several class lists of the split batch
    path) and axis-aligned boxes on a lattice, where different box-vertex pairs give Minkowski vertices that
    are equal BY VALUE (the reference cancels horizon edges by value, code/nans.h:251-254)."""
    from nans_projekat_b200 import scenes
    rng = np.random.default_rng(77)
    n = 1 << 17
    p = scenes.narrowphase_pairs(n, seed=78, types=rng.integers(0, 5, n).astype(np.int32))
    c8 = scenes.CORNERS.astype(np.float32)
    m = n // 2                                                   # second half: lattice boxes
    pos_a = rng.integers(-2, 3, (m, 3)).astype(np.float32) * np.float32(0.25)
    pos_b = pos_a + rng.integers(-4, 5, (m, 3)).astype(np.float32) * np.float32(0.25)
    sa = rng.choice([0.5, 1.0, 2.0], (m, 1, 3)).astype(np.float32)
    sb = rng.choice([0.5, 1.0, 2.0], (m, 1, 3)).astype(np.float32)
    p["pos_a"][m:], p["pos_b"][m:] = pos_a, pos_b
    p["verts_a"][m:] = c8[None] * sa + pos_a[:, None]
    p["verts_b"][m:] = c8[None] * sb + pos_b[:, None]
    args = (p["type"], p["pos_a"], p["verts_a"], p["rad_a"], p["pos_b"], p["verts_b"], p["rad_b"])
    g = nb200.check_collision(*args)
    o = oracle.check_collision_batch(*args)
    assert np.array_equal(g["gjk"], o["gjk"]) and np.array_equal(g["hit"], o["hit"])
    h = o["hit"] == 1
    assert h[:m].any() and h[m:].any()
    for k in ("N", "PA", "PB"):
        assert_bit_equal(g[k][h], o[k][h], k)


def test_narrowphase_edge_cases(nb200, oracle):
    """Degenerate inputs: coincident shapes, exactly aligned cubes, NaN/inf vertices, zero radius,
    touching faces.  Flags must still match the oracle bit for bit."""
    from nans_projekat_b200 import scenes
    C = scenes.CORNERS
    cases = []

    def add(t, pa, va, ra, pb, vb, rb):
        cases.append((t, np.asarray(pa, np.float32), np.asarray(va, np.float32), np.float32(ra),
                      np.asarray(pb, np.float32), np.asarray(vb, np.float32), np.float32(rb)))
    z3 = (0, 0, 0)
    add(0, z3, C, 0, z3, C, 0)                                   # coincident cubes (frame-0 state)
    add(0, z3, C, 0, (0, 0.9, 0), C + (0, 0.9, 0), 0)            # exactly aligned: collinearity degeneracy
    add(0, z3, C, 0, (0.05, 0.9, 0.02), C + np.float32((0.05, 0.9, 0.02)), 0)  # SURVEY known answer: hit
    add(0, z3, C, 0, (0.05, 1.1, 0.02), C + np.float32((0.05, 1.1, 0.02)), 0)  # no hit
    add(0, z3, C, 0, (1.0, 0.01, 0.02), C + np.float32((1.0, 0.01, 0.02)), 0)  # touching faces
    add(1, z3, C, 0, (0.1, 0.7, 0), C, 0.25)                     # cube vs sphere, known hit
    add(1, z3, C, 0, (0.1, 0.7, 0), C, 0.0)                      # zero radius
    add(3, z3, C, 0.5, (0.3, 0.2, 0.1), C, 0.5)                  # overlapping spheres: reference says no
    add(4, (0.3, 0.6, 0.1), C, 0.3, z3, C, 0)                    # sphere vs floor-type box
    nanv = C.copy(); nanv[3, 1] = np.nan
    add(0, z3, nanv, 0, (0.2, 0.5, 0.1), C + np.float32((0.2, 0.5, 0.1)), 0)
    infv = C.copy(); infv[0, 0] = np.inf
    add(2, z3, C, 0, (0.2, -0.5, 0.1), infv, 0)
    add(0, (np.nan, 0, 0), C, 0, (0.2, 0.5, 0.1), C, 0)
    cols = list(zip(*cases))
    args = (np.array(cols[0], np.int32), np.stack(cols[1]), np.stack(cols[2]), np.array(cols[3], np.float32),
            np.stack(cols[4]), np.stack(cols[5]), np.array(cols[6], np.float32))
    g = nb200.check_collision(*args)
    o = oracle.check_collision_batch(*args)
    assert np.array_equal(g["hit"], o["hit"]) and np.array_equal(g["gjk"], o["gjk"])
    assert list(o["hit"][:4]) == [0, 0, 1, 0]
    h = o["hit"] == 1
    for k in ("N", "PA", "PB"):
        assert_bit_equal(g[k][h], o[k][h], k)
    # empty batch is a no-op
    e = nb200.check_collision(np.zeros(0, np.int32), np.zeros((0, 3)), np.zeros((0, 8, 3)), np.zeros(0),
                              np.zeros((0, 3)), np.zeros((0, 8, 3)), np.zeros(0))
    assert len(e["hit"]) == 0


# ------------------------------------------------------------------------------------ stages
def test_stages_golden(nb200, oracle, golden_dir):
    """Each stage function from the reference binary's recorded states (<=16+16 bodies, one floor)."""
    recs, dt = load_stage_records(os.path.join(golden_dir, "stages.npz"))
    total = 0
    for rec in recs:
        w = world_from_stage_record(oracle, rec)
        gw = nb200.World(scene_from_oracle_world(w))
        gw.integrate_forces(dt)
        d = gw.download()
        for k in ("vel", "angvel", "force", "torque"):
            assert_bit_equal(getattr(d, k), rec["s1"][k], f"IntegrateForces {k}")
        gw.detect_collisions()
        c = gw.contacts()
        assert c.tobytes() == rec["contacts"].tobytes(), "DetectCollisions: contact list (order, indices, floats)"
        total += len(c)
        gw.solve_constraints(dt)
        d = gw.download()
        for k in ("vel", "angvel"):
            assert_bit_equal(getattr(d, k), rec["s2"][k], f"SolveConstraints {k}")
        gw.integrate_velocities(dt)
        d = gw.download()
        for k in ("pos", "ang"):
            assert_bit_equal(getattr(d, k), rec["s3"][k], f"IntegrateVelocities {k}")
        # vertex rebuild vs the restatement (itself pinned to the binary's Model rebuild)
        w.pos[...] = rec["s3"]["pos"]; w.ang[...] = rec["s3"]["ang"]
        w.rebuild_vertices()
        assert_bit_equal(d.verts, w.verts, "rebuilt vertices")
        assert gw.stats()["overflow"] == 0
        gw.close()
    assert total > 500


def test_demo_trajectory_golden(nb200, oracle, golden_dir):
    """Config C1: the reference's Init scene, 1000 steps recorded from nans.so through
    SimUpdateAndRender; every GPU step starts from the recorded previous state."""
    from nans_projekat_b200 import scenes
    z = np.load(os.path.join(golden_dir, "demo_traj.npz"))
    s = scenes.demo_scene()
    s.st_verts[0] = z["floor_verts"]
    gw = nb200.World(s)
    n_steps = len(z["pos"])
    checked = 0
    zero = np.zeros_like(s.force)
    for k in range(1, n_steps):
        if k == 201:
            continue  # ShootSphere frame: the game layer teleports sphere 0 (covered by the plugin test)
        for f in ("pos", "vel", "ang", "angvel", "verts"):
            getattr(s, f)[...] = z[f][k - 1]
        s.force[...] = zero; s.torque[...] = zero
        gw.upload(s, fields=STATE)
        gw.step(DT)
        d = gw.download()
        assert gw.stats()["n_contacts"] == z["ncontacts"][k], f"step {k}: contact count"
        for f in ("pos", "vel", "ang", "angvel", "verts"):
            assert_bit_equal(getattr(d, f), z[f][k], f"step {k} {f}")
        checked += 1
    assert checked >= 990
    gw.close()


# ------------------------------------------------------------------------------------ larger worlds
def _settle(oracle, s, steps):
    w = world_from_scene(oracle, s)
    w.rebuild_vertices()
    for _ in range(steps):
        w.step(DT, prefilter=True)
    return w


@pytest.mark.gpu
def test_deferred_force_upload_is_visible_to_every_entry_point(nb200, oracle):
    """A force/torque-only upload_async is unpacked lazily (after the step's detection phase, so the copy
    overlaps broadphase + narrowphase).  Every other entry point must see it as if it had been applied
    at once: download, add_force, a second upload of the same field, the stage calls, snapshot."""
    from nans_projekat_b200 import scenes
    s = scenes.cube_drop(n=600, dims=(9, 8, 9), spacing=1.05, jitter=0.04)
    s.pos[:, 1] -= 0.4
    w = world_from_scene(oracle, s); w.rebuild_vertices()
    base = scene_from_oracle_world(w)
    rng = np.random.default_rng(21)
    f1 = rng.normal(0, 3, (base.nb, 3)).astype(np.float32); t1 = rng.normal(0, 0.5, (base.nb, 3)).astype(np.float32)
    f2 = rng.normal(0, 3, (base.nb, 3)).astype(np.float32)
    io1 = scenes.Scene(base.n_cubes, 0, base.n_statics); io1.force[...], io1.torque[...] = f1, t1
    io2 = scenes.Scene(base.n_cubes, 0, base.n_statics); io2.force[...] = f2
    ga, gb = nb200.World(base), nb200.World(base)
    for g in (ga, gb):                       # settle into contact, graphs captured
        for _ in range(25):
            g.step(DT)
    # 1) download right after the deferred upload
    gb.upload_async(io1, fields=("force", "torque"))
    d = gb.download(fields=("force", "torque"))
    assert_bit_equal(d.force, f1, "deferred force visible to download"); assert_bit_equal(d.torque, t1, "torque")
    # 2) a second upload of one field overrides the first, add_force adds on top, then a step
    gb.upload_async(io1, fields=("force", "torque"))
    gb.upload_async(io2, fields=("force",))
    gb.add_force(3, (1.0, 2.0, 3.0), (0.5, 0.0, -0.5))
    gb.step(DT)
    ga.upload(io1, fields=("force", "torque")); ga.upload(io2, fields=("force",))
    ga.add_force(3, (1.0, 2.0, 3.0), (0.5, 0.0, -0.5))
    ga.step(DT)
    da, db = ga.download(), gb.download()
    for fld in ("pos", "vel", "ang", "angvel", "force", "torque", "verts"):
        assert_bit_equal(getattr(db, fld), getattr(da, fld), f"after step: {fld}")
    # 3) stage calls and snapshot/restore with an upload pending
    gb.upload_async(io1, fields=("force", "torque")); gb.snapshot()
    gb.integrate_forces(DT); gb.detect_collisions(); gb.solve_constraints(DT); gb.integrate_velocities(DT)
    ga.upload(io1, fields=("force", "torque"))
    ga.integrate_forces(DT); ga.detect_collisions(); ga.solve_constraints(DT); ga.integrate_velocities(DT)
    da, db = ga.download(), gb.download()
    for fld in ("pos", "vel", "ang", "angvel", "verts"):
        assert_bit_equal(getattr(db, fld), getattr(da, fld), f"after stage calls: {fld}")
    gb.restore()
    d = gb.download(fields=("force", "torque"))
    assert_bit_equal(d.force, f1, "snapshot holds the deferred force")
    # 4) several frames of upload_async -> step -> synchronous download (the e2e loop of bench.py)
    ga.upload(gb.download(), fields=("pos", "vel", "force", "ang", "angvel", "torque", "verts"))
    for k in range(6):
        io1.force[...] = rng.normal(0, 3, (base.nb, 3)).astype(np.float32)
        gb.upload_async(io1, fields=("force", "torque")); gb.step(DT); db = gb.download(fields=("pos", "ang"))
        ga.upload(io1, fields=("force", "torque")); ga.step(DT); da = ga.download(fields=("pos", "ang"))
        for fld in ("pos", "ang"):
            assert_bit_equal(getattr(db, fld), getattr(da, fld), f"e2e frame {k}: {fld}")
    ga.close(); gb.close()


@pytest.mark.parametrize("n,dims,steps", [(1500, (12, 11, 12), 40)])
def test_drop_scene_steps_vs_oracle(nb200, oracle, n, dims, steps):
    """Config C2 at reduced size (cubes dropped into the static box: floor + 4 wall slabs).
    The oracle runs the reference's all-pairs DetectCollisions; the GPU runs broadphase + GJK/EPA.
    Contact lists (order included) and post-step state must be bit-identical."""
    from nans_projekat_b200 import scenes
    s = scenes.cube_drop(n=n, dims=dims, spacing=1.02, jitter=0.04)
    s.pos[:, 1] -= 0.45   # start close to the floor so contacts appear within a few steps
    w = _settle(oracle, s, 6)
    gw = nb200.World(scene_from_oracle_world(w))
    max_contacts = 0
    for step in range(steps):
        gw.upload(w, fields=STATE)
        gw.step(DT)
        oc = w.step(DT, prefilter=(step % 4 != 0))   # every 4th step: the true all-pairs reference loop
        gc = gw.contacts()
        st = gw.stats()
        assert st["overflow"] == 0
        assert gc.tobytes() == oc.tobytes(), f"step {step}: contact list {len(gc)} vs {len(oc)}"
        d = gw.download()
        for f in ("pos", "vel", "ang", "angvel", "verts"):
            assert_bit_equal(getattr(d, f), getattr(w, f), f"step {step} {f}")
        max_contacts = max(max_contacts, len(oc))
    assert max_contacts > n // 2, "scene never developed contacts"
    gw.close()


def test_mixed_world_with_spheres(nb200, oracle):
    """Cubes + spheres + several statics, incl. the reference's CS index-swap de-duplication quirk
    (code/nans.cpp:1479-1489) which only fires when cube and sphere indices overlap."""
    from nans_projekat_b200 import scenes
    rng = np.random.default_rng(5)
    s = scenes.Scene(40, 40, 3)
    p = rng.uniform(0, 3.2, (80, 3)); p[:, 1] = rng.uniform(0.3, 2.5, 80)
    for i in range(40):
        s.set_cube(i, p[i], ang=rng.uniform(-180, 180, 3))
        s.set_sphere(i, p[40 + i], radius=float(rng.uniform(0.2, 0.5)))
    s.vel[:] = rng.normal(0, 1, (80, 3)); s.angvel[:] = rng.normal(0, 2, (80, 3))
    s.set_static(0, (1.2, -0.5, 1.0), (100.0, 1.0, 100.0))
    s.set_static(1, (-0.6, 2.0, 1.5), (1.0, 6.0, 20.0), size_for_moi=20)
    s.set_static(2, (3.9, 2.0, 1.5), (1.0, 6.0, 20.0), size_for_moi=20)
    w = world_from_scene(oracle, s)
    w.rebuild_vertices()
    gw = nb200.World(scene_from_oracle_world(w))
    dropped = 0
    for step in range(25):
        gw.upload(w, fields=STATE)
        gw.step(DT)
        before = w.copy()
        oc = w.step(DT, prefilter=False)
        gc = gw.contacts()
        assert gc.tobytes() == oc.tobytes(), f"step {step}"
        d = gw.download()
        for f in ("pos", "vel", "ang", "angvel", "verts"):
            assert_bit_equal(getattr(d, f), getattr(w, f), f"step {step} {f}")
        types = set(oc["type"].tolist())
        cs = oc[oc["type"] == 1]
        have = set(zip(cs["a"].tolist(), cs["b"].tolist()))
        dropped += sum(1 for (a, b) in have if (b, a) in have and b < a)
        del before
    assert {0, 1, 2, 4} <= types or len(oc) > 0
    assert dropped == 0   # a surviving (a,b),(b,a) couple would contradict the reference's quirk
    gw.close()


def test_broadphase_superset_and_order(nb200, oracle):
    """Pair set = every pair whose (inflated) AABBs overlap — a superset of the reference's hit
    set — emitted in reference list order; checked against a brute-force AABB oracle."""
    from nans_projekat_b200 import scenes
    rng = np.random.default_rng(9)
    s = scenes.Scene(300, 100, 2)
    p = rng.uniform(0, 9, (400, 3))
    for i in range(300):
        s.set_cube(i, p[i], ang=rng.uniform(-180, 180, 3))
    for j in range(100):
        s.set_sphere(j, p[300 + j], radius=float(rng.uniform(0.1, 0.5)))
    s.set_static(0, (4.5, -0.5, 4.5), (100.0, 1.0, 100.0))
    s.set_static(1, (-0.5, 4.0, 4.5), (1.0, 10.0, 30.0))
    w = world_from_scene(oracle, s)
    w.rebuild_vertices()
    gw = nb200.World(scene_from_oracle_world(w))
    gw.detect_collisions()
    a, b = gw.pairs()
    got = list(zip(a.tolist(), b.tolist()))
    assert len(set(got)) == len(got), "duplicate candidate pairs"
    # every reference contact must be among the candidates
    oc = w.detect(prefilter=False)
    nc = s.n_cubes

    def rows(c):
        t, x, y = int(c["type"]), int(c["a"]), int(c["b"])
        ra = x + nc if t in (3, 4) else x
        rb = -(y + 1) if t in (2, 4) else (y + nc if t in (1, 3) else y)
        return ra, rb
    want = [rows(c) for c in oc]
    assert set(want) <= set(got)
    # order: the candidate list restricted to hits is exactly the reference list
    hits = set(want)
    assert [g for g in got if g in hits] == want
    gc = gw.contacts()
    assert gc.tobytes() == oc.tobytes()
    gw.close()


def test_empty_and_tiny_worlds(nb200, oracle):
    from nans_projekat_b200 import scenes
    s = scenes.Scene(0, 0, 1)
    s.set_static(0, (1.2, -0.5, 1.0), (100.0, 1.0, 100.0))
    gw = nb200.World(s)
    gw.step(DT)
    assert gw.stats()["n_contacts"] == 0
    gw.close()
    s = scenes.Scene(1, 0, 1)
    s.set_cube(0, (0.3, 0.4, 0.2))
    s.set_static(0, (1.2, -0.5, 1.0), (100.0, 1.0, 100.0))
    w = world_from_scene(oracle, s)
    w.rebuild_vertices()
    gw = nb200.World(scene_from_oracle_world(w))
    gw.step(DT)
    oc = w.step(DT)
    assert gw.contacts().tobytes() == oc.tobytes() and len(oc) == 1
    d = gw.download()
    for f in ("pos", "vel", "ang", "angvel", "verts"):
        assert_bit_equal(getattr(d, f), getattr(w, f), f)
    # dt = 0 (the reference's first frame): identity on velocities
    gw.upload(w, fields=STATE)
    gw.integrate_forces(np.float32(0.0))
    d = gw.download()
    assert_bit_equal(d.vel, w.vel, "dt=0 vel")
    gw.close()


def test_capacity_overflow_is_reported(nb200, oracle):
    from nans_projekat_b200 import scenes
    from nans_projekat_b200._lib import NansError
    s = scenes.Scene(64, 0, 1)
    rng = np.random.default_rng(2)
    for i in range(64):
        s.set_cube(i, rng.uniform(0, 1.0, 3))      # all overlapping
    s.set_static(0, (1.2, -0.5, 1.0), (100.0, 1.0, 100.0))
    w = world_from_scene(oracle, s); w.rebuild_vertices()
    gw = nb200.World(scene_from_oracle_world(w), max_pairs=100, max_contacts=50)
    gw.detect_collisions()
    with pytest.raises(NansError):
        gw.stats()
    assert gw.stats(strict=False)["overflow"] & 1
    gw.close()


def test_batched_independent_worlds(nb200, oracle):
    """Config C4 at reduced size: many independent worlds in one device world (world_id): bodies of
    different worlds never pair up, and every world evolves exactly as if stepped alone by the oracle."""
    from nans_projekat_b200 import scenes
    n_worlds, cpw, spw = 12, 10, 4
    s = scenes.batched_worlds(n_worlds=n_worlds, cubes_per=cpw, spheres_per=spw, seed=3)
    s.pos[:, 1] -= 0.25
    nc = s.n_cubes
    whole = world_from_scene(oracle, s); whole.rebuild_vertices()
    s.verts[:] = whole.verts; s.st_verts[:] = whole.st_verts
    # one oracle world per independent world
    subs = []
    for wd in range(n_worlds):
        rows = np.nonzero(s.world_id == wd)[0]
        cu, sp = rows[rows < nc], rows[rows >= nc]
        o = oracle.World(len(cu), len(sp), 1)
        r = np.concatenate([cu, sp])
        for f in ("pos", "vel", "force", "ang", "angvel", "torque", "scale"):
            getattr(o, f)[:] = getattr(s, f)[r]
        o.mass[:], o.moi[:], o.radius[:] = s.mass[r], s.moi[r], s.radius[r]
        o.verts[:] = s.verts[cu]
        for f in ("st_pos", "st_ang", "st_scale", "st_mass", "st_moi", "st_verts"):
            getattr(o, f)[...] = getattr(s, f)
        subs.append((r, cu, sp, o))
    gw = nb200.World(s)
    total = 0
    for step in range(40):
        gw.step(DT)
        gc = gw.contacts()
        d = gw.download()
        assert gw.stats()["overflow"] == 0
        for wd, (r, cu, sp, o) in enumerate(subs):
            oc = o.step(DT, prefilter=False)
            for f in ("pos", "vel", "ang", "angvel"):
                assert_bit_equal(getattr(d, f)[r], getattr(o, f), f"step {step} world {wd} {f}")
            assert_bit_equal(d.verts[cu], o.verts, f"step {step} world {wd} verts")
            # this world's contacts out of the global list, re-indexed locally
            cube_local = {int(g): i for i, g in enumerate(cu)}
            sph_local = {int(g - nc): i for i, g in enumerate(sp)}
            mine = []
            for c in gc:
                t, a, b = int(c["type"]), int(c["a"]), int(c["b"])
                a_sph = t in (3, 4)
                if (a_sph and a in sph_local) or (not a_sph and a in cube_local):
                    cc = c.copy()
                    cc["a"] = sph_local[a] if a_sph else cube_local[a]
                    if t in (0,): cc["b"] = cube_local[b]
                    elif t in (1, 3): cc["b"] = sph_local[b]
                    mine.append(cc)
            mine = np.array(mine, dtype=gc.dtype) if mine else np.zeros(0, gc.dtype)
            assert mine.tobytes() == oc.tobytes(), f"step {step} world {wd}: contacts {len(mine)} vs {len(oc)}"
            total += len(oc)
    assert total > 1000
    gw.close()


def test_c2_full_size_drop_scene(nb200, oracle):
    """Config C2 at FULL size: 10 000 cubes dropped into the static box (floor + 4 wall slabs).
    The GPU free-runs to a contact-rich state; from that identical state GPU and oracle are stepped
    side by side (re-seeded each step): contact lists and body state bit-identical.  Plus size-independent
    properties over the whole run: finite state, contact list strictly ordered, no capacity overflow."""
    from nans_projekat_b200 import scenes
    s = scenes.cube_drop(n=10000)
    w = world_from_scene(oracle, s)
    w.rebuild_vertices()
    gw = nb200.World(scene_from_oracle_world(w))
    max_contacts = 0
    for step in range(150):
        gw.step(DT)
        if step % 25 == 24:
            st = gw.stats()
            assert st["overflow"] == 0
            c = gw.contacts()
            key = c["type"].astype(np.int64) * 0  # order: CC, CF, SF, CS, SS then (a, b)
            seg = np.array([0, 3, 1, 4, 2])[c["type"]]
            k = (seg.astype(np.int64) << 50) | (c["a"].astype(np.int64) << 25) | c["b"].astype(np.int64)
            assert (np.diff(k) > 0).all(), "contact list must be strictly increasing in reference order"
            max_contacts = max(max_contacts, len(c))
    d = gw.download()
    assert all(np.isfinite(getattr(d, f)).all() for f in ("pos", "vel", "ang", "angvel", "verts"))
    assert max_contacts > 5000, "the drop never developed a contact-rich state"
    for f in STATE:
        getattr(w, f)[...] = getattr(d, f)
    for step in range(4):
        gw.upload(w, fields=STATE)
        gw.step(DT)
        oc = w.step(DT, prefilter=True)
        gc = gw.contacts()
        assert gc.tobytes() == oc.tobytes(), f"step {step}: contact list {len(gc)} vs {len(oc)}"
        d = gw.download()
        for f in ("pos", "vel", "ang", "angvel", "verts"):
            assert_bit_equal(getattr(d, f), getattr(w, f), f"step {step} {f}")
    gw.close()


def test_model_matrices_instanced_draw_data(nb200, oracle):
    """nans_world_models: the draw section's Model = T*Rx*Ry*Rz*S (code/nans.cpp:1870-1881, 1913-1941, 1971-1990)
    for every cube, sphere and static, bit-identical to the oracle's (which is pinned to nans.so's Models)."""
    from nans_projekat_b200 import scenes
    rng = np.random.default_rng(3)
    s = scenes.random_small_world(rng, 16, 16)
    s.scale[3] = (0.5, 1.0, 0.5)                 # the reference's squashed debug box
    big = scenes.batched_worlds(n_worlds=64, cubes_per=48, spheres_per=16, seed=2)
    big.ang[:] = rng.uniform(-400, 400, big.ang.shape).astype(np.float32)
    for sc in (s, big):
        gw = nb200.World(sc)
        m = gw.models()
        assert m.shape == (sc.nb + sc.n_statics, 16)
        idx = np.arange(sc.nb) if sc.nb < 100 else rng.choice(sc.nb, 400, replace=False)
        for i in idx:
            scale = sc.scale[i] if i < sc.n_cubes else (sc.radius[i],) * 3
            want = oracle.model_vertices(sc.pos[i], sc.ang[i], scale)[0].reshape(16)
            assert_bit_equal(m[i], want, f"Model of body {i}")
        for k in range(sc.n_statics):
            want = oracle.model_vertices(sc.st_pos[k], sc.st_ang[k], sc.st_scale[k])[0].reshape(16)
            assert_bit_equal(m[sc.nb + k], want, f"Model of static {k}")
        gw.close()


def test_pipelined_io_matches_synchronous_io(nb200, oracle):
    """upload_async / download_async / wait (copies overlapping the step on their own streams) must give
    exactly the states the synchronous upload / step / download sequence gives; snapshot/restore returns
    the world to the snapshot bit for bit."""
    from nans_projekat_b200 import scenes
    s = scenes.cube_drop(n=2000, dims=(13, 12, 13), spacing=1.05, jitter=0.04)
    s.pos[:, 1] -= 0.4
    w = world_from_scene(oracle, s); w.rebuild_vertices()
    base = scene_from_oracle_world(w)
    rng = np.random.default_rng(8)
    ga, gb = nb200.World(base), nb200.World(base)
    gb.snapshot()
    outs = [scenes.Scene(base.n_cubes, 0, base.n_statics) for _ in range(2)]
    ticket, sync_states = None, []
    forces = [(rng.normal(0, 3, (base.nb, 3)).astype(np.float32), rng.normal(0, 0.5, (base.nb, 3)).astype(np.float32))
              for _ in range(12)]
    io = scenes.Scene(base.n_cubes, 0, base.n_statics)
    for k, (f, t) in enumerate(forces):
        io.force[...], io.torque[...] = f, t
        ga.upload(io, fields=("force", "torque")); ga.step(DT)
        sync_states.append(ga.download(fields=("pos", "ang")))
    for k, (f, t) in enumerate(forces):
        io2 = scenes.Scene(base.n_cubes, 0, base.n_statics)      # a fresh host buffer per frame (stays valid)
        io2.force[...], io2.torque[...] = f, t
        gb.upload_async(io2, fields=("force", "torque")); gb.step(DT)
        tk = gb.download_async(outs[k & 1], fields=("pos", "ang"))
        if ticket is not None:
            gb.wait(ticket)
            for fld in ("pos", "ang"):
                assert_bit_equal(getattr(outs[(k - 1) & 1], fld), getattr(sync_states[k - 1], fld), f"frame {k-1} {fld}")
        ticket = tk
    gb.wait(-1)
    for fld in ("pos", "ang"):
        assert_bit_equal(getattr(outs[(len(forces) - 1) & 1], fld), getattr(sync_states[-1], fld), f"last frame {fld}")
    gb.restore(); gb.synchronize()
    d = gb.download()
    for fld in ("pos", "vel", "ang", "angvel", "verts"):
        assert_bit_equal(getattr(d, fld), getattr(base, fld), f"restore {fld}")
    ga.close(); gb.close()


def test_c4_full_size_batched_worlds_properties(nb200):
    """Config C4 at FULL size (4096 worlds x 64 bodies = 262 144 bodies in one device world):
    size-independent properties — no contact ever joins two different worlds, the list stays in
    reference order, the state stays finite, and an identical second world reproduces it bit for bit."""
    from nans_projekat_b200 import scenes
    s = scenes.batched_worlds(n_worlds=4096, cubes_per=48, spheres_per=16, seed=1)
    s.pos[:, 1] -= 0.25
    nc = s.n_cubes
    ga, gb = nb200.World(s), nb200.World(s)
    ga.rebuild_vertices(); gb.rebuild_vertices()
    for step in range(30):
        ga.step(DT); gb.step(DT)
    assert ga.stats()["overflow"] == 0
    c = ga.contacts()
    assert len(c) > 100_000
    wid = s.world_id
    a_row = np.where(np.isin(c["type"], (3, 4)), c["a"] + nc, c["a"])
    dyn = np.isin(c["type"], (0, 1, 3))
    b_row = np.where(np.isin(c["type"], (1, 3)), c["b"] + nc, c["b"])
    assert (wid[a_row[dyn]] == wid[b_row[dyn]]).all(), "a contact joins two independent worlds"
    seg = np.array([0, 3, 1, 4, 2])[c["type"]]
    k = (seg.astype(np.int64) << 50) | (c["a"].astype(np.int64) << 25) | c["b"].astype(np.int64)
    assert (np.diff(k) > 0).all()
    da, db = ga.download(), gb.download()
    for f in ("pos", "vel", "ang", "angvel", "verts"):
        assert np.isfinite(getattr(da, f)).all()
        assert_bit_equal(getattr(da, f), getattr(db, f), f"determinism {f}")
    assert gb.contacts().tobytes() == c.tobytes()
    ga.close(); gb.close()
