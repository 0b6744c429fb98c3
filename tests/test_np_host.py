"""The DEVICE narrowphase source (nans_projekat_b200/csrc/narrowphase.cuh, the code the CUDA kernels run)
compiled as host C++ (tests/np_host_shim.cpp: one thread, __shared__ = static storage, __f*_rn = plain IEEE
fp32 with -ffp-contract=off) and checked bit for bit against the oracle.  It proves the algorithm of the
kernel source without a GPU; the GPU tests prove the compiled kernels.  Checked: GJK + EPA as the world kernel
runs them, and the capped-then-resumed GJK of the split batch path (config C3)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from nans_projekat_b200 import scenes

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
f32p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_int32)


def _build(name, flags):
    os.makedirs(BUILD, exist_ok=True)
    out = os.path.join(BUILD, f"libnp_host_{name}.so")
    src = os.path.join(HERE, "np_host_shim.cpp")
    hdr = os.path.join(HERE, "..", "nans_projekat_b200", "csrc", "narrowphase.cuh")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
                               "-I/usr/local/cuda/include", "-D__device__=",
                               "-D__forceinline__=inline __attribute__((always_inline))", *flags, "-o", out, src])
    return C.CDLL(out)


@pytest.fixture(scope="module", params=[("default", []), ("gjk_capped", ["-DNANS_NP_GJK_CAPPED=8"]),
                                        ("gjk_capped1", ["-DNANS_NP_GJK_CAPPED=1"])],
                ids=lambda p: p[0])
def host_np(request):
    if not os.path.exists("/usr/local/cuda/include/cuda_runtime.h"):
        pytest.skip("CUDA headers not found")
    return _build(*request.param)


def _run(lib, p):
    n = len(p["type"])
    hit, gjk = np.zeros(n, np.int32), np.zeros(n, np.int32)
    N, PA, PB = (np.zeros((n, 3), np.float32) for _ in range(3))
    mf = np.zeros(1, np.int32)
    fp = lambda a: np.ascontiguousarray(a, np.float32).ctypes.data_as(f32p)
    keep = [np.ascontiguousarray(p[k], np.float32) for k in ("pos_a", "verts_a", "rad_a", "pos_b", "verts_b", "rad_b")]
    t = np.ascontiguousarray(p["type"], np.int32)
    ovf = lib.np_host_check_collision_batch(n, t.ctypes.data_as(i32p), *[a.ctypes.data_as(f32p) for a in keep],
                                            hit.ctypes.data_as(i32p), gjk.ctypes.data_as(i32p), fp(N), fp(PA), fp(PB),
                                            mf.ctypes.data_as(i32p))
    # fp() of a fresh contiguous copy would not write back: N/PA/PB are already contiguous float32
    return dict(hit=hit, gjk=gjk, N=N, PA=PA, PB=PB, ovf=ovf)


def _check(lib, oracle, p, what):
    g = _run(lib, p)
    o = oracle.check_collision_batch(p["type"], p["pos_a"], p["verts_a"], p["rad_a"], p["pos_b"], p["verts_b"], p["rad_b"])
    assert g["ovf"] == 0, what
    assert np.array_equal(g["hit"], o["hit"]) and np.array_equal(g["gjk"], o["gjk"]), f"{what}: flags differ"
    h = o["hit"] == 1
    assert h.any(), what
    for k in ("N", "PA", "PB"):
        assert np.array_equal(g[k][h].view(np.uint32), o[k][h].view(np.uint32)), f"{what}: {k} differs"


def _exact_grid(n, seed, types=None):
    """Axis-aligned boxes on a 0.25 lattice: many Minkowski vertices coincide BY VALUE although they come from
    different box-vertex pairs (the reference cancels horizon edges by value, code/nans.h:251-254)."""
    rng = np.random.default_rng(seed)
    c8 = scenes.CORNERS.astype(np.float32)
    t = np.zeros(n, np.int32) if types is None else types
    pos_a = rng.integers(-2, 3, (n, 3)).astype(np.float32) * np.float32(0.25)
    pos_b = pos_a + rng.integers(-4, 5, (n, 3)).astype(np.float32) * np.float32(0.25)
    sa = rng.choice([0.5, 1.0, 2.0], (n, 1, 3)).astype(np.float32)
    sb = rng.choice([0.5, 1.0, 2.0], (n, 1, 3)).astype(np.float32)
    r = rng.uniform(0.1, 0.5, n).astype(np.float32)
    return dict(type=t, pos_a=pos_a, verts_a=np.ascontiguousarray(c8[None] * sa + pos_a[:, None], np.float32), rad_a=r,
                pos_b=pos_b, verts_b=np.ascontiguousarray(c8[None] * sb + pos_b[:, None], np.float32), rad_b=r.copy())


def test_random_pairs_all_types(host_np, oracle):
    _check(host_np, oracle, scenes.narrowphase_pairs(60000, seed=5), "random CC/CS/SS")
    _check(host_np, oracle, scenes.narrowphase_pairs(30000, seed=6, rotated=False), "axis-aligned")
    types = np.random.default_rng(1).integers(0, 5, 30000).astype(np.int32)
    _check(host_np, oracle, scenes.narrowphase_pairs(30000, seed=7, types=types), "all five pair types")


def test_value_equal_vertices(host_np, oracle):
    _check(host_np, oracle, _exact_grid(60000, 3), "lattice boxes")
    types = np.random.default_rng(2).integers(0, 5, 20000).astype(np.int32)
    _check(host_np, oracle, _exact_grid(20000, 4, types), "lattice, all types")


def test_degenerate_boxes(host_np, oracle):
    n = 40000
    p = scenes.narrowphase_pairs(n, seed=9, mix=(1, 0, 0))
    rng = np.random.default_rng(10)
    m = rng.random(n) < 0.3; p["verts_a"][m, :, 1] = p["verts_a"][m, 0:1, 1]          # flat in y
    m = rng.random(n) < 0.2; p["verts_b"][m, 4:] = p["verts_b"][m, :4]                  # duplicated vertices
    m = rng.random(n) < 0.05; p["verts_b"][m] = p["verts_b"][m, 0:1]                    # a point
    m = rng.random(n) < 0.02; p["verts_a"][m, rng.integers(0, 8), rng.integers(0, 3)] = np.nan
    m = rng.random(n) < 0.02; p["verts_b"][m, rng.integers(0, 8), rng.integers(0, 3)] = np.inf
    m = rng.random(n) < 0.02; p["pos_b"][m] = p["pos_a"][m]                              # coincident centres
    _check(host_np, oracle, p, "degenerate boxes")


def test_settled_pile_pairs(host_np, oracle):
    """Candidate pairs of a small settled pile (the bench workload's geometry: resting face contacts)."""
    s = scenes.cube_pile(n_side=8, layers=8, seed=7)
    ow = oracle.World(s.n_cubes, s.n_spheres, s.n_statics)
    for f in s.ARRAYS:
        getattr(ow, f)[...] = getattr(s, f)
    ow.rebuild_vertices()
    for _ in range(46):
        ow.step(np.float32(1 / 60.), prefilter=True)
    v = ow.verts; lo, hi = v.min(1), v.max(1)
    pa, pb = [], []
    for i in range(len(v)):
        j = np.nonzero(np.all((lo[i] <= hi[i + 1:]) & (lo[i + 1:] <= hi[i]), axis=1))[0] + i + 1
        pa.append(np.full(len(j), i)); pb.append(j)
    pa, pb = np.concatenate(pa), np.concatenate(pb)
    z = np.zeros(len(pa), np.float32)
    p = dict(type=np.zeros(len(pa), np.int32), pos_a=ow.pos[pa].copy(), verts_a=v[pa].copy(), rad_a=z,
             pos_b=ow.pos[pb].copy(), verts_b=v[pb].copy(), rad_b=z)
    assert len(pa) > 300
    _check(host_np, oracle, p, "pile pairs")


def test_emptied_polytope_golden(host_np, oracle):
    """The pair that the 1 M-cube headline parity test caught in round 2: the first EPA point sees all four faces
    of the start tetrahedron and every horizon edge cancels, so the reference's triangle vector is EMPTY and its
    next iteration reads the stale Triangle[0] (code/nans.cpp:807-866) -- which holds the last face.  nans.so
    reports a hit with that face's normal (tests/golden/epa_emptied.npz holds its outputs); the device source
    and the oracle must both reproduce it."""
    z = np.load(os.path.join(HERE, "golden", "epa_emptied.npz"))
    p = {k: z[k] for k in ("type", "pos_a", "verts_a", "rad_a", "pos_b", "verts_b", "rad_b")}
    g = _run(host_np, p)
    o = oracle.check_collision_batch(p["type"], p["pos_a"], p["verts_a"], p["rad_a"], p["pos_b"], p["verts_b"], p["rad_b"])
    assert z["ref_hit"].sum() >= 3 and len(z["ref_hit"]) >= 40
    for r, who in ((g, "device source"), (o, "oracle")):
        assert np.array_equal(r["hit"], z["ref_hit"]), who
        h = z["ref_hit"] == 1
        for k in ("N", "PA", "PB"):
            assert np.array_equal(r[k][h].view(np.uint32), z[f"ref_{k}"][h].view(np.uint32)), f"{who}: {k}"


def test_vertex_hash_respects_float_equality(host_np):
    """EPA's class scan skips an earlier vertex whose stored 16-bit hash differs from the new vertex's
    (csrc/narrowphase.cuh, epa_store_vertex): sound only if vectors that COMPARE equal hash equal.  The one case of
    different bits comparing equal is +0 / -0 (NaN compares equal to nothing and is handled before the scan)."""
    lib = host_np
    lib.np_host_hash16.restype = C.c_uint
    lib.np_host_hash16.argtypes = [C.c_float] * 3
    rng = np.random.default_rng(3)
    vals = [0.0, -0.0, 1.0, -1.0, 1e-45, -1e-45, 3.5, np.float32(np.inf), np.float32(-np.inf)]
    vecs = [(a, b, c) for a in vals for b in vals for c in vals]
    vecs += [tuple(rng.normal(0, 3, 3).astype(np.float32)) for _ in range(2000)]
    h = {}
    for v in vecs:
        key = tuple(np.float32(x) + np.float32(0.0) for x in v)       # the equality class: -0 -> +0
        got = lib.np_host_hash16(*[C.c_float(float(x)) for x in v])
        assert 0 <= got < 65536
        assert h.setdefault(key, got) == got, f"{v}: equal vectors, different hashes"
    # and it does discriminate: random vectors rarely collide
    rnd = [lib.np_host_hash16(*[C.c_float(float(x)) for x in v]) for v in vecs[-2000:]]
    assert len(set(rnd)) > 1900
