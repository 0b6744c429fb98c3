"""The solver's per-contact arithmetic (nans_projekat_b200/csrc/solver_constraint.cuh: constraint_prepare +
constraint_apply, with the accumulation form the kernel runs) compiled for the host and swept over a contact list
in the reference's order, against the oracle's SolveConstraints: velocities bit for bit.  The parallel schedule
(which contact may run when) is what the GPU tests add."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import assert_bit_equal, world_from_scene
from test_solver_gpu import cloud, synthetic_contacts

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
f32p = C.POINTER(C.c_float)
DT = np.float32(1 / 60.)


@pytest.fixture(scope="module")
def sweep():
    if not os.path.exists("/usr/local/cuda/include/cuda_runtime.h"):
        pytest.skip("CUDA headers not found")
    os.makedirs(BUILD, exist_ok=True)
    out = os.path.join(BUILD, "libsolver_host.so")
    src = os.path.join(HERE, "solver_host_shim.cpp")
    hdrs = [os.path.join(HERE, "..", "nans_projekat_b200", "csrc", h) for h in ("solver_constraint.cuh", "solver_accum.cuh")]
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(f) for f in [src] + hdrs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
                               "-I/usr/local/cuda/include", "-D__device__=",
                               "-D__forceinline__=inline __attribute__((always_inline))", "-o", out, src])
    lib = C.CDLL(out)

    def run(scene, contacts):
        fp = lambda a: np.ascontiguousarray(a, np.float32)
        vel, angvel = fp(scene.vel).copy(), fp(scene.angvel).copy()
        keep = [fp(scene.pos), fp(scene.mass), fp(scene.moi), fp(scene.st_pos), fp(scene.st_mass), fp(scene.st_moi)]
        c = np.ascontiguousarray(contacts)
        lib.solver_host_sweep(scene.n_cubes, scene.n_spheres, scene.n_statics, keep[0].ctypes.data_as(f32p),
                              vel.ctypes.data_as(f32p), angvel.ctypes.data_as(f32p), keep[1].ctypes.data_as(f32p),
                              keep[2].ctypes.data_as(f32p), keep[3].ctypes.data_as(f32p), keep[4].ctypes.data_as(f32p),
                              keep[5].ctypes.data_as(f32p), c.ctypes.data_as(C.c_void_p), len(c), C.c_float(DT))
        return vel, angvel
    return run


def _check(sweep, oracle, scene, contacts, what):
    ow = world_from_scene(oracle, scene)
    ow.solve(DT, contacts)
    vel, angvel = sweep(scene, contacts)
    for got, ref, name in ((vel, ow.vel, "vel"), (angvel, ow.angvel, "angvel")):
        nan = np.isnan(got) & np.isnan(ref)          # NaN payload / sign is not part of the contract
        assert_bit_equal(np.where(nan, 0, got), np.where(nan, 0, ref), f"{what}: {name}")
    return int(np.isnan(ow.vel).any(1).sum())


@pytest.mark.parametrize("n_cubes,n_spheres,n_contacts,seed", [(64, 0, 400, 1), (3000, 0, 20000, 2), (2000, 500, 30000, 3)])
def test_synthetic_contact_lists(sweep, oracle, n_cubes, n_spheres, n_contacts, seed):
    scene = cloud(n_cubes, n_spheres, seed)
    _check(sweep, oracle, scene, synthetic_contacts(scene, n_contacts, seed + 100), "synthetic")


def test_hub_chain_and_degenerate_normals(sweep, oracle):
    scene = cloud(1500, 0, 11)
    c = synthetic_contacts(scene, 6000, 12, hub=700, chain=(100, 1400))
    _check(sweep, oracle, scene, c, "hub + chain")
    # a zero normal: normalize(0) is NaN, so the reference's `N == 0` rescue (:1115-1119) never fires and the
    # NaN spreads along the dependency chains -- the same bodies must go NaN on both sides
    scene2 = cloud(400, 0, 13)
    c2 = synthetic_contacts(scene2, 900, 14)
    c2["n"][5::40] = 0.0
    n_nan = _check(sweep, oracle, scene2, c2, "zero normals")
    assert 0 < n_nan < scene2.nb


def test_contacts_of_a_stepped_scene(sweep, oracle):
    """Real contact lists: a drop scene stepped by the oracle, the solve of every 5th frame re-done on the host."""
    from nans_projekat_b200 import scenes
    s = scenes.cube_drop(n=300, dims=(7, 7, 7), spacing=1.05, jitter=0.05)
    ow = world_from_scene(oracle, s); ow.rebuild_vertices()
    n_checked = 0
    for step in range(40):
        ow.integrate_forces(DT)
        contacts = ow.detect(prefilter=True)
        if step % 5 == 4 and len(contacts):
            snap = scenes.Scene(ow.n_cubes, ow.n_spheres, ow.n_statics)
            for f in snap.ARRAYS:
                getattr(snap, f)[...] = getattr(ow, f)
            ref = ow.copy(); ref.solve(DT, contacts)
            vel, angvel = sweep(snap, contacts)
            assert_bit_equal(vel, ref.vel, f"step {step}: vel"); assert_bit_equal(angvel, ref.angvel, f"step {step}: angvel")
            n_checked += len(contacts)
        ow.solve(DT, contacts)
        ow.integrate_velocities(DT)
    assert n_checked > 200
