"""The integrators' arithmetic (nans_projekat_b200/csrc/integrate.cuh: RK4, Model = T*Rx*Ry*Rz*S -> vertices with
the glibc-exact sinf/cosf) compiled for the host, against the oracle: velocities, poses and vertices bit for bit."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import assert_bit_equal, world_from_scene

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
f32p = C.POINTER(C.c_float)
P = lambda a: a.ctypes.data_as(f32p)


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists("/usr/local/cuda/include/cuda_runtime.h"):
        pytest.skip("CUDA headers not found")
    os.makedirs(BUILD, exist_ok=True)
    out = os.path.join(BUILD, "libintegrate_host.so")
    src = os.path.join(HERE, "integrate_host_shim.cpp")
    hdrs = [os.path.join(HERE, "..", "nans_projekat_b200", "csrc", h) for h in ("integrate.cuh", "glibc_sincosf.cuh")]
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(f) for f in [src] + hdrs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
                               "-I/usr/local/cuda/include", "-D__device__=",
                               "-D__forceinline__=inline __attribute__((always_inline))", "-o", out, src])
    return C.CDLL(out)


@pytest.mark.parametrize("dt", [0.0, 1 / 60., 0.004])
def test_integrators_vs_oracle(lib, oracle, dt):
    from nans_projekat_b200 import scenes
    rng = np.random.default_rng(5)
    nc, ns = 3000, 500
    s = scenes.Scene(nc, ns, 1)
    nb = nc + ns
    s.pos[:] = rng.uniform(-50, 50, (nb, 3)); s.vel[:] = rng.normal(0, 3, (nb, 3)); s.angvel[:] = rng.normal(0, 4, (nb, 3))
    s.ang[:] = rng.uniform(-720, 720, (nb, 3))                       # degrees (the reference feeds them to glm::radians)
    s.ang[::11] = 0.0; s.ang[1::13, 0] = 90.0; s.ang[2::17, 1] = -180.0
    s.force[:] = rng.normal(0, 20, (nb, 3)) * (rng.random((nb, 1)) < 0.5); s.torque[:] = rng.normal(0, 2, (nb, 3))
    s.mass[:] = rng.uniform(0.3, 5, nb); s.moi[:] = rng.uniform(0.02, 2, nb)
    s.scale[:nc] = rng.choice([0.5, 1.0, 2.5], (nc, 3)); s.radius[nc:] = 0.3
    dt = np.float32(dt)
    ow = world_from_scene(oracle, s)
    ow.integrate_forces(dt)
    vel, angvel, force, torque = (np.ascontiguousarray(getattr(s, f), np.float32).copy() for f in ("vel", "angvel", "force", "torque"))
    lib.integrate_forces_host(nb, P(vel), P(angvel), P(force), P(torque), P(np.ascontiguousarray(s.mass, np.float32)),
                              P(np.ascontiguousarray(s.moi, np.float32)), C.c_float(dt))
    assert_bit_equal(vel, ow.vel, "RK4 V"); assert_bit_equal(angvel, ow.angvel, "RK4 W")
    assert not force.any() and not torque.any() and not ow.force.any()
    ow.integrate_velocities(dt)
    ow.rebuild_vertices()                   # the draw section's Model rebuild + UpdateVertices, from the new pose
    pos, ang = np.ascontiguousarray(s.pos, np.float32).copy(), np.ascontiguousarray(s.ang, np.float32).copy()
    verts = np.zeros((nc, 8, 3), np.float32)
    lib.integrate_velocities_host(nb, nc, P(pos), P(ang), P(vel), P(angvel), P(np.ascontiguousarray(s.scale, np.float32)),
                                  P(verts), C.c_float(dt))
    assert_bit_equal(pos, ow.pos, "Position"); assert_bit_equal(ang, ow.ang, "Angles")
    assert_bit_equal(verts, ow.verts, "vertices")
