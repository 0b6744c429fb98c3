"""GPU parity tests of the exact-order solver on SYNTHETIC contact lists (nans_set_contacts ->
nans_solve_constraints) against the oracle's sequential SolveConstraints on the same list: arbitrary
dependency graphs (long chains, hub bodies, duplicate pairs, all five contact types), degenerate
normals, NaN/inf inputs, and agreement of the three solver implementations.

The reference applies contacts strictly in list order (code/nans.cpp:1539-1548); the CUDA solvers run
the same list as a dependency graph and must reproduce the sequential result bit for bit.
"""
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import assert_bit_equal, bits, world_from_scene

pytestmark = pytest.mark.gpu

DT = np.float32(1 / 60.)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CC, CS, CF, SS, SF = 0, 1, 2, 3, 4      # contact types (include/nans_b200.h, oracle/nans_oracle.h)


@pytest.fixture(scope="module")
def nb200():
    from nans_projekat_b200 import _lib, world
    assert _lib.lib().nans_device_count() > 0, "no CUDA device: the product path has no CPU fallback"
    return world


def cloud(n_cubes, n_spheres, seed):
    """Bodies scattered in a box with random velocities / spins / masses (no geometry needed: the
    contact lists below are synthetic)."""
    from nans_projekat_b200.scenes import Scene
    r = np.random.default_rng(seed)
    s = Scene(n_cubes, n_spheres, 1)
    nb = n_cubes + n_spheres
    s.pos[...] = r.uniform(-20, 20, (nb, 3)).astype(np.float32)
    s.vel[...] = r.normal(0, 2, (nb, 3)).astype(np.float32)
    s.angvel[...] = r.normal(0, 1, (nb, 3)).astype(np.float32)
    s.vel[::7] = 0.0            # exact zeros: signed-zero handling of the accumulation
    s.angvel[::5] = 0.0
    s.mass[...] = r.uniform(0.5, 4.0, nb).astype(np.float32)
    s.moi[...] = r.uniform(0.05, 1.0, nb).astype(np.float32)
    s.radius[n_cubes:] = 0.5
    s.st_pos[0] = (0.0, -30.0, 0.0)
    s.st_mass[0] = 1.0e5
    s.st_moi[0] = 1.0e9
    return s


def synthetic_contacts(scene, n, seed, hub=None, chain=None):
    """A contact list in the reference's block order (CC, CF, SF, CS, SS; each block sorted by (a, b)),
    with random contact points near the bodies and random (some zero) normals."""
    from nans_projekat_b200.world import CONTACT_DTYPE
    r = np.random.default_rng(seed)
    nc, ns = scene.n_cubes, scene.n_spheres
    blocks = []

    def block(t, a, b):
        order = np.lexsort((b, a))
        c = np.zeros(len(a), CONTACT_DTYPE)
        c["type"], c["a"], c["b"] = t, a[order], b[order]
        blocks.append(c)

    k = n // (3 if ns == 0 else 6)
    a = r.integers(0, nc - 1, 2 * k); b = r.integers(0, nc, 2 * k)
    keep = a != b
    a, b = np.minimum(a, b)[keep], np.maximum(a, b)[keep]
    if hub is not None:          # one cube touched by very many others
        others = r.choice(np.delete(np.arange(nc), hub), size=min(nc - 1, 600), replace=False)
        a = np.concatenate([a, np.minimum(others, hub)]); b = np.concatenate([b, np.maximum(others, hub)])
    if chain is not None:        # i - i+1 - i+2 ...: a dependency chain as long as the list
        i = np.arange(chain[0], chain[1])
        a = np.concatenate([a, i]); b = np.concatenate([b, i + 1])
    block(CC, a.astype(np.int32), b.astype(np.int32))
    cf = np.unique(r.integers(0, nc, k)).astype(np.int32)
    block(CF, cf, np.zeros_like(cf))
    if ns:
        sf = np.unique(r.integers(0, ns, k // 2)).astype(np.int32)
        block(SF, sf, np.zeros_like(sf))
        block(CS, r.integers(0, nc, k).astype(np.int32), r.integers(0, ns, k).astype(np.int32))
        sa = r.integers(0, ns - 1, k); sb = r.integers(0, ns, k)
        ok = sa != sb
        block(SS, np.minimum(sa, sb)[ok].astype(np.int32), np.maximum(sa, sb)[ok].astype(np.int32))
    c = np.concatenate(blocks)
    m = len(c)
    row_a = np.where((c["type"] == SF) | (c["type"] == SS), nc + c["a"], c["a"])
    c["point_a"] = scene.pos[row_a] + r.normal(0, 0.5, (m, 3)).astype(np.float32)
    c["point_b"] = c["point_a"] + r.normal(0, 0.05, (m, 3)).astype(np.float32)
    nrm = r.normal(0, 1, (m, 3)).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True).astype(np.float32)
    nrm[::13] = (0.0, 1.0, 0.0)  # axis-aligned: exact zeros in T1/T2
    nrm[::17] = (-1.0, 0.0, 0.0)
    c["n"] = nrm
    return c


def solve_both(nb200, oracle, scene, contacts):
    ow = world_from_scene(oracle, scene)
    ow.solve(DT, contacts)
    gw = nb200.World(scene, max_contacts=len(contacts) + 16)
    gw.set_contacts(contacts)
    gw.solve_constraints(DT)
    d = gw.download(fields=("vel", "angvel"))
    st = gw.stats()
    gw.close()
    return ow, d, st


@pytest.mark.parametrize("n_cubes,n_spheres,n_contacts,seed", [(64, 0, 400, 1), (3000, 0, 20000, 2),
                                                               (2000, 500, 30000, 3), (20000, 0, 150000, 4)])
def test_synthetic_contact_lists(nb200, oracle, n_cubes, n_spheres, n_contacts, seed):
    scene = cloud(n_cubes, n_spheres, seed)
    c = synthetic_contacts(scene, n_contacts, seed + 100)
    ow, d, st = solve_both(nb200, oracle, scene, c)
    assert_bit_equal(d.vel, ow.vel, "vel after SolveConstraints")
    assert_bit_equal(d.angvel, ow.angvel, "angvel after SolveConstraints")
    assert st["n_contacts"] == len(c) and st["solver_levels"] >= 2


def test_hub_body_and_long_chain(nb200, oracle):
    """One cube in 600 contacts (its sequence is strictly serial) and a 4000-contact chain i - i+1."""
    scene = cloud(5000, 0, 7)
    c = synthetic_contacts(scene, 6000, 8, hub=2500, chain=(100, 4100))
    ow, d, st = solve_both(nb200, oracle, scene, c)
    assert_bit_equal(d.vel, ow.vel, "vel")
    assert_bit_equal(d.angvel, ow.angvel, "angvel")
    assert st["solver_levels"] >= 4000


def test_repeated_pair_and_single_contact(nb200, oracle):
    from nans_projekat_b200.world import CONTACT_DTYPE
    scene = cloud(8, 0, 9)
    one = synthetic_contacts(scene, 12, 10)[:1]
    ow, d, _ = solve_both(nb200, oracle, scene, one)
    assert_bit_equal(d.vel, ow.vel, "single contact vel")
    rep = np.zeros(200, CONTACT_DTYPE)          # the same pair 200 times: a pure chain on both bodies
    rep[:] = one[0]
    rep["type"], rep["a"], rep["b"] = CC, 2, 5
    ow, d, st = solve_both(nb200, oracle, scene, rep)
    assert_bit_equal(d.vel, ow.vel, "repeated pair vel")
    assert_bit_equal(d.angvel, ow.angvel, "repeated pair angvel")
    assert st["solver_levels"] == 200


def test_nan_and_inf_inputs_take_the_literal_path(nb200, oracle):
    """NaN / inf contact data (the 70-iteration loop then runs the reference's compares literally).
    NaN payloads differ between SSE and the GPU, so NaNs are compared by position, everything else by bits."""
    scene = cloud(400, 0, 11)
    c = synthetic_contacts(scene, 3000, 12)
    c["n"][5] = (np.nan, 0.0, 0.0)
    c["n"][9] = 0.0             # normalize(0) is NaN, so the (0,0,0) fallback of :1115-1119 never triggers
    c["point_a"][40] = (np.inf, 0.0, 0.0)
    c["point_b"][77] = (1.0e30, -1.0e30, 1.0e30)
    c["n"][120] = (1.0e-30, 0.0, 0.0)
    scene.vel[int(c["a"][200])] = (np.nan, 0.0, 0.0)
    ow, d, _ = solve_both(nb200, oracle, scene, c)
    for name, got, want in (("vel", d.vel, ow.vel), ("angvel", d.angvel, ow.angvel)):
        gn, wn = np.isnan(got), np.isnan(want)
        assert (gn == wn).all(), f"{name}: NaN positions differ"
        assert gn.any(), "the poisoned inputs should reach some velocity"
        assert (bits(got)[~gn] == bits(want)[~wn]).all(), f"{name}: finite values differ"


def sweep_position(n):
    """Host copy of csrc/solver.cu's mix_bits: the position of contact c in the SHUFFLED sweep."""
    k = 1 if n <= 2 else int(n - 1).bit_length()
    m = np.uint64((1 << k) - 1)
    s_ = np.uint64((k + 1) // 2)
    x = np.arange(n, dtype=np.uint64)
    x = (x * np.uint64(0x9E3779B1)) & m
    x ^= x >> s_
    x = (x * np.uint64(0x85EBCA6B)) & m
    x ^= x >> s_
    return x


@pytest.mark.parametrize("n_cubes,n_spheres,n_contacts,seed", [(64, 0, 400, 21), (3000, 500, 30000, 22), (20000, 0, 150000, 23)])
def test_shuffled_sweep_is_the_sequential_sweep_in_its_documented_order(nb200, oracle, n_cubes, n_spheres, n_contacts, seed):
    """The throughput mode (NANS_SOLVER_SHUFFLED) is still ONE Gauss-Seidel pass of the reference's Constraint
    (code/nans.cpp:1021-1329), only in another order: the oracle's sequential SolveConstraints over the contact
    list permuted by the same bijective hash must give the GPU's velocities bit for bit.  It is NOT the
    reference's list order: the deviation from the exact sweep is reported, not asserted to be small."""
    scene = cloud(n_cubes, n_spheres, seed)
    c = synthetic_contacts(scene, n_contacts, seed + 100)
    pos = sweep_position(len(c))
    assert len(np.unique(pos)) == len(c), "the sweep order must be a bijection"
    order = np.argsort(pos, kind="stable")
    ow = world_from_scene(oracle, scene)
    ow.solve(DT, c[order])
    ex = world_from_scene(oracle, scene)
    ex.solve(DT, c)
    gw = nb200.World(scene, max_contacts=len(c) + 16)
    gw.set_solver("shuffled")
    gw.set_contacts(c)
    gw.solve_constraints(DT)
    d = gw.download(fields=("vel", "angvel"))
    st_shuffled = gw.stats()
    gw.set_solver("exact")
    gw.upload(scene, fields=("vel", "angvel"))
    gw.set_contacts(c)
    gw.solve_constraints(DT)
    d_exact = gw.download(fields=("vel", "angvel"))
    st_exact = gw.stats()
    gw.close()
    assert_bit_equal(d.vel, ow.vel, "shuffled sweep vel")
    assert_bit_equal(d.angvel, ow.angvel, "shuffled sweep angvel")
    assert_bit_equal(d_exact.vel, ex.vel, "exact sweep vel (after switching back)")
    assert st_shuffled["solver_levels"] <= st_exact["solver_levels"]
    dev = np.abs(d.vel.astype(np.float64) - ex.vel) / np.maximum(np.abs(ex.vel), 1.0)
    print(f"shuffled vs exact sweep: levels {st_shuffled['solver_levels']} vs {st_exact['solver_levels']}, "
          f"max relative velocity deviation {dev.max():.3g}")


def test_shuffled_sweep_on_a_pile_world(nb200, oracle):
    """Whole steps in throughput mode on a small pile: finite, deterministic, and each step equal to the oracle's
    step whose contact list is swept in the documented shuffled order."""
    from nans_projekat_b200 import scenes
    s = scenes.cube_pile(n_side=16, layers=8, seed=5)
    s.pos[:, 1] *= np.float32(0.985)
    w = world_from_scene(oracle, s); w.rebuild_vertices()
    sc = s.copy(); sc.verts[:] = w.verts; sc.st_verts[:] = w.st_verts
    gw = nb200.World(sc); gw.set_solver("shuffled")
    fields = ("pos", "vel", "force", "ang", "angvel", "torque", "verts")
    total = 0
    for step in range(6):
        gw.upload(w, fields=fields)
        gw.step(DT)
        gc = gw.contacts()
        # the oracle's step with the solve in shuffled order: forces, detect, permuted solve, velocities, rebuild
        w.integrate_forces(DT)
        oc = w.detect(prefilter="grid")
        assert gc.tobytes() == oc.tobytes()
        w.solve(DT, oc[np.argsort(sweep_position(len(oc)), kind="stable")])
        w.integrate_velocities(DT); w.rebuild_vertices()
        d = gw.download(fields=fields)
        for f in ("pos", "vel", "ang", "angvel", "verts"):
            assert_bit_equal(getattr(d, f), getattr(w, f), f"step {step} {f}")
        total += len(oc)
    assert total > 2000
    gw.close()
