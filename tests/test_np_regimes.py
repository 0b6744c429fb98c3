"""CheckCollision (code/nans.cpp:907-966) outside the unit-cube regime of config C3: boxes scaled over six decades
per axis (plates, rods), pairs 10..10^6 units from the origin (the fp32 grid of the coordinates is coarse there),
faces touching to within +-1e-7..1e-3 (GJK's >= on the support dot, :528), and shapes 1e-6..1e-2 across (GJK's
absolute 1e-4 direction cut-off, :742 -- most of those never intersect, by the reference's own arithmetic).
Three-way, bit for bit, on the CPU: the reference's own binary (oracle/_ref/nans.so) == the oracle restatement ==
the DEVICE narrowphase source compiled for the host (tests/np_host_shim.cpp).  A 4.7e7-pair campaign of the same
generator found no difference (profiles/README.md, fourth session)."""
import numpy as np
import pytest

from nans_projekat_b200 import scenes
from test_np_host import _build, _run

REGIMES = ("scaled", "far", "touch", "tiny")


def regime_pairs(rng, n, regime):
    types = rng.integers(0, 5, n).astype(np.int32)
    c8 = scenes.CORNERS.astype(np.float64)
    sa = sb = np.ones((n, 1, 3))
    base = np.zeros((n, 3))
    if regime == "scaled":
        sa, sb = np.exp(rng.uniform(-3, 3, (n, 1, 3))), np.exp(rng.uniform(-3, 3, (n, 1, 3)))
        off = rng.uniform(-1, 1, (n, 3)) * (sa[:, 0] + sb[:, 0]) * 0.6
    elif regime == "far":
        base = rng.uniform(-1, 1, (n, 3)) * 10.0 ** rng.integers(1, 7, (n, 1))
        off = rng.uniform(-1.2, 1.2, (n, 3))
    elif regime == "touch":
        off = np.zeros((n, 3))
        off[np.arange(n), rng.integers(0, 3, n)] = 1.0 + rng.choice(
            [0, 1e-7, -1e-7, 1e-6, -1e-6, 1e-4, -1e-4, 5e-4, 1e-3, -1e-3], n)
        off += rng.uniform(-0.3, 0.3, (n, 3)) * (rng.random((n, 1)) < 0.5)
    else:
        s = 10.0 ** rng.uniform(-6, -2, (n, 1, 1))
        sa = sb = s * np.ones((n, 1, 3))
        off = rng.uniform(-1.2, 1.2, (n, 3)) * s[:, 0]
    eye = np.eye(3)[None]
    rot_on = (rng.random(n) < (0.0 if regime == "touch" else 0.7))[:, None, None]
    ra = np.where(rot_on, scenes._rand_rot(rng, n), eye)
    rb = np.where(rot_on, scenes._rand_rot(rng, n), eye)
    if regime == "touch":      # a third of the touching pairs a milliradian out of alignment
        a = rng.uniform(-1e-3, 1e-3, n) * (rng.random(n) < 0.3)
        z, o = np.zeros(n), np.ones(n)
        rb = np.stack([np.stack([np.cos(a), -np.sin(a), z], -1), np.stack([np.sin(a), np.cos(a), z], -1),
                       np.stack([z, z, o], -1)], 1)
    pos_a, pos_b = base.astype(np.float32), (base + off).astype(np.float32)
    va = (np.einsum("nij,nkj->nki", ra, c8[None] * sa) + pos_a[:, None, :].astype(np.float64)).astype(np.float32)
    vb = (np.einsum("nij,nkj->nki", rb, c8[None] * sb) + pos_b[:, None, :].astype(np.float64)).astype(np.float32)
    scale = np.minimum(sa[:, 0].mean(1), sb[:, 0].mean(1))
    return dict(type=types, pos_a=pos_a, verts_a=np.ascontiguousarray(va),
                rad_a=(rng.uniform(0.05, 0.8, n) * scale).astype(np.float32), pos_b=pos_b,
                verts_b=np.ascontiguousarray(vb), rad_b=(rng.uniform(0.05, 0.8, n) * scale).astype(np.float32))


def _same(a, b, keys, what):
    assert np.array_equal(a["hit"], b["hit"]), f"{what}: hit flags differ"
    h = a["hit"] == 1
    for k in keys:
        assert np.array_equal(a[k][h].view(np.uint32), b[k][h].view(np.uint32)), f"{what}: {k} differs"


@pytest.mark.parametrize("regime", REGIMES)
def test_three_way_parity_outside_the_unit_cube_regime(oracle, regime):
    p = regime_pairs(np.random.default_rng(REGIMES.index(regime) + 41), 30000, regime)
    args = (p["type"], p["pos_a"], p["verts_a"], p["rad_a"], p["pos_b"], p["verts_b"], p["rad_b"])
    o = oracle.check_collision_batch(*args)
    assert (o["hit"] == 1).sum() > 30
    g = _run(_build("default", []), p)
    assert g["ovf"] == 0
    assert np.array_equal(g["gjk"], o["gjk"]), "device source vs oracle: GJK flags differ"
    _same(g, o, ("N", "PA", "PB"), "device source vs oracle")
    if oracle.ref() is not None:      # the reference's own binary, when staged (oracle/Makefile)
        _same(oracle.ref_check_collision_batch(*args), o, ("N", "PA", "PB"), "reference binary vs oracle")
