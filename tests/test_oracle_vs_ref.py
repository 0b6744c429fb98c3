"""Live pinning of the CPU restatement against the reference's own prebuilt nans.so
(oracle/_ref, through the dlopen harness).  Fresh random inputs every run of the seed list;
set NANS_FUZZ_SCALE=20 for the long campaign quoted in oracle/README.md."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import assert_bit_equal

SCALE = int(os.environ.get("NANS_FUZZ_SCALE", "1"))


@pytest.mark.parametrize("rotated", [True, False])
def test_check_collision_matches_binary(oracle, ref, rotated):
    from nans_projekat_b200 import scenes
    n = 4000 * SCALE
    types = np.tile(np.arange(5, dtype=np.int32), n // 5)
    p = scenes.narrowphase_pairs(n, seed=99 + rotated, rotated=rotated, types=types)
    args = (p["type"], p["pos_a"], p["verts_a"], p["rad_a"], p["pos_b"], p["verts_b"], p["rad_b"])
    r = oracle.ref_check_collision_batch(*args)
    o = oracle.check_collision_batch(*args)
    assert np.array_equal(r["hit"], o["hit"])
    h = r["hit"] == 1
    for k in ("N", "PA", "PB"):
        assert_bit_equal(o[k][h], r[k][h], k)


def test_stages_match_binary(oracle, ref):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_golden import scene_to_ref_state
    from nans_projekat_b200 import scenes
    rng = np.random.default_rng(555)
    dt = np.float32(1 / 60)
    ncontacts = 0
    for trial in range(150 * SCALE):
        nc, ns = int(rng.integers(1, 17)), int(rng.integers(0, 17))
        s = scenes.random_small_world(rng, nc, ns, spread=float(rng.choice([1.0, 2.0, 3.0])))
        st = scene_to_ref_state(s)
        w = oracle.world_from_ref_state(st)
        ref.nansref_integrate_forces(oracle._sp(st), C.c_float(dt)); w.integrate_forces(dt)
        pr = oracle.ref_detect(st, dt); co = w.detect()
        assert oracle.contacts_from_ref_pairs(pr).tobytes() == co.tobytes()
        ncontacts += len(co)
        ref.nansref_solve_constraints(oracle._sp(st), C.c_float(dt), pr.ctypes.data_as(C.c_void_p), len(pr))
        w.solve(dt, co)
        ref.nansref_integrate_velocities(oracle._sp(st), C.c_float(dt)); w.integrate_velocities(dt)
        w2 = oracle.world_from_ref_state(st)
        for k in ("pos", "vel", "force", "ang", "angvel", "torque"):
            assert_bit_equal(getattr(w, k), getattr(w2, k), f"trial {trial} {k}")
    assert ncontacts > 1000


def test_model_rebuild_matches_binary(oracle, ref):
    """Model = T*Rx*Ry*Rz*S and the 8 vertices, through the exported SimUpdateAndRender
    (draw section, code/nans.cpp:1870-1881,1913-1941)."""
    rng = np.random.default_rng(3)
    perm = np.zeros(1 << 20, np.uint8)
    st = perm[:oracle.STATE_DTYPE.itemsize].view(oracle.STATE_DTYPE)
    mem = oracle.MemoryStruct(perm.ctypes.data, perm.nbytes, perm.ctypes.data, perm.nbytes, 1)
    inp = np.zeros(1, oracle.INPUT_DTYPE); ren = np.zeros(256, np.uint8)
    for t in range(300 * SCALE):
        s = st[0]
        ref.nansref_init(oracle._sp(st)); s["SphereCount"] = 0
        for i in range(4):
            c = s["Cubes"][i]
            c["Position"] = rng.uniform(-50, 50, 3)
            c["Angles"] = rng.uniform(-1, 1, 3) * rng.choice([1, 30, 400, 5000])
            c["Size"] = rng.uniform(.3, 2)
        s["Floor"]["Angles"] = rng.uniform(-10, 10, 3) * (t % 2)
        ref.nansref_sim_update_and_render(C.byref(mem), oracle._sp(inp), ren.ctypes.data_as(C.c_void_p),
                                          C.c_float(0.0))
        for i in range(4):
            ref.nansref_update_vertices(oracle._sp(st), i)
        ref.nansref_floor_update_vertices(oracle._sp(st))
        for i in range(4):
            c = s["Cubes"][i]
            h = np.float32(0.5)
            sc = (c["Size"],) * 3 if i < 3 else (c["Size"] * h, c["Size"] * np.float32(1), c["Size"] * h)
            m, v = oracle.model_vertices(c["Position"], c["Angles"], sc)
            assert_bit_equal(m.reshape(16), c["Model"], "cube Model")
            assert_bit_equal(v, c["Vertices"], "cube Vertices")
        f = s["Floor"]
        m, v = oracle.model_vertices(f["Position"], f["Angles"], (f["Size"], 1.0, f["Size"]))
        assert_bit_equal(m.reshape(16), f["Model"], "floor Model")
        assert_bit_equal(v, f["Vertices"], "floor Vertices")
