"""Parity AT THE HEADLINE SIZE: the 1 M-cube piles bench.py times, compared with the CPU oracle.

bench.py's state preparation is repeated here (the pile free-runs on the GPU for `settle` steps), then GPU and
oracle are stepped side by side from that identical state: the contact list (reference order) and the body
state after the step must be bit-identical.  The oracle visits the same pairs in the same order as the
reference's all-pairs loop through its uniform-grid prefilter (tests/test_oracle_grid.py), which is what
makes 10^6 bodies checkable in seconds.  Both shapes: SURVEY.md §8(d)'s 100 x 100 x 100 (config C5) and the
flat 250 x 250 x 16 pile of round 1.  Reference path: code/nans.cpp:1758-1762.
"""
import numpy as np
import pytest

from helpers import assert_bit_equal, world_from_scene

pytestmark = pytest.mark.gpu

STATE = ("pos", "vel", "force", "ang", "angvel", "torque", "verts")
DT = np.float32(1 / 60.)


@pytest.mark.parametrize("side,layers,settle", [(100, 100, 80), (250, 16, 40)])
def test_one_million_cube_pile_steps_bit_exact(oracle, side, layers, settle):
    from nans_projekat_b200 import scenes
    from nans_projekat_b200.world import World
    s = scenes.cube_pile(n_side=side, layers=layers, seed=7)
    assert s.n_cubes == 1_000_000
    gw = World(s)
    gw.rebuild_vertices()
    for _ in range(settle):
        gw.step(DT)
    assert gw.stats()["overflow"] == 0
    d = gw.download(fields=STATE)
    w = world_from_scene(oracle, s)
    w.rebuild_vertices()                     # statics' vertices (bit-exact on both sides: test_gpu_parity)
    for f in STATE:
        getattr(w, f)[...] = getattr(d, f)
    n_min = 150_000 if layers == 100 else 800_000
    for step in range(2):
        gw.upload(w, fields=STATE)
        gw.step(DT)
        oc = w.step(DT, prefilter="grid", cap=8_000_000)
        st = gw.stats()
        assert st["overflow"] == 0
        gc = gw.contacts()
        assert len(oc) > n_min, f"the pile is not in contact ({len(oc)} contacts)"
        assert gc.tobytes() == oc.tobytes(), f"step {step}: contact list {len(gc)} vs {len(oc)}"
        d = gw.download(fields=STATE)
        for f in ("pos", "vel", "ang", "angvel", "verts"):
            assert_bit_equal(getattr(d, f), getattr(w, f), f"{side}x{side}x{layers} step {step} {f}")
    gw.close()
