"""The oracle's spatial prefilter (prefilter == "grid", test infrastructure for 10^6-body parity) must give
the SAME contact list, in the SAME order, as the reference's all-pairs loop (code/nans.cpp:1352-1536) that
the restatement is pinned to nans.so with.  CPU only."""
import numpy as np
import pytest

from nans_projekat_b200 import scenes
from helpers import world_from_scene


def _rebuild(O, w):
    w.rebuild_vertices()
    return w


@pytest.mark.parametrize("seed", range(12))
def test_grid_equals_all_pairs_small_worlds(oracle, seed):
    """<= 16 + 16 bodies, cubes and spheres mixed (all five pair types, incl. the live CS 'Exists' quirk)."""
    rng = np.random.default_rng(100 + seed)
    for _ in range(40):
        s = scenes.random_small_world(rng, int(rng.integers(1, 17)), int(rng.integers(0, 17)),
                                      spread=float(rng.uniform(0.6, 2.5)))
        w = _rebuild(oracle, world_from_scene(oracle, s))
        a = w.detect(prefilter=False)
        g = w.detect(prefilter="grid")
        assert a.tobytes() == g.tobytes(), f"{len(a)} vs {len(g)} contacts"


def test_grid_equals_aabb_filter_10k_drop(oracle):
    """10 000 cubes in the static box (config C2), mid-fall and in contact: grid == all-pairs AABB filter
    (itself equal to the unfiltered loop: tests/test_gpu_parity.py, test_oracle_vs_ref.py)."""
    s = scenes.cube_drop(n=10000, seed=1)
    w = _rebuild(oracle, world_from_scene(oracle, s))
    dt = np.float32(1 / 60.)
    for _ in range(3):
        w.step(dt, prefilter="grid")
    w.pos[:, 1] *= np.float32(0.8)          # squeeze the lattice: plenty of contacts at once
    w.rebuild_vertices()
    a = w.detect(prefilter=True)
    g = w.detect(prefilter="grid")
    assert len(a) > 5000
    assert a.tobytes() == g.tobytes()


def test_grid_whole_step_equals_filtered_step(oracle):
    s = scenes.cube_pile(n_side=12, layers=6, seed=3)
    s.pos[:, 1] *= np.float32(0.97)
    wa = _rebuild(oracle, world_from_scene(oracle, s))
    wb = wa.copy()
    dt = np.float32(1 / 60.)
    for _ in range(4):
        ca = wa.step(dt, prefilter=False)
        cb = wb.step(dt, prefilter="grid")
        assert ca.tobytes() == cb.tobytes()
        for f in ("pos", "vel", "ang", "angvel", "verts"):
            assert getattr(wa, f).tobytes() == getattr(wb, f).tobytes(), f
    assert len(ca) > 0


def test_grid_far_outliers_and_nonfinite(oracle):
    """Bodies flung far away (an exploding pile) and a non-finite one: clamped cells stay correct."""
    rng = np.random.default_rng(5)
    s = scenes.random_small_world(rng, 16, 8, spread=1.0)
    s.pos[3] = (1e6, 2.0, -3e5)
    s.pos[5] = (np.nan, 1.0, 1.0)
    w = _rebuild(oracle, world_from_scene(oracle, s))
    a = w.detect(prefilter=True)
    g = w.detect(prefilter="grid")
    assert a.tobytes() == g.tobytes()
