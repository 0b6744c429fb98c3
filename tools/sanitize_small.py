"""Small run for compute-sanitizer (memcheck / racecheck / initcheck): every kernel of the step on a
mixed world (cubes + spheres + statics), the stand-alone narrowphase, snapshot/restore, transfers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nans_projekat_b200 import scenes
from nans_projekat_b200.world import World, check_collision

rng = np.random.default_rng(5)
s = scenes.Scene(300, 60, 3)
p = rng.uniform(0, 6, (360, 3)); p[:, 1] = rng.uniform(0.3, 4.0, 360)
for i in range(300):
    s.set_cube(i, p[i], ang=rng.uniform(-180, 180, 3))
for j in range(60):
    s.set_sphere(j, p[300 + j], radius=float(rng.uniform(0.2, 0.5)))
s.set_static(0, (3, -0.5, 3), (100.0, 1.0, 100.0))
s.set_static(1, (-0.6, 2.0, 3), (1.0, 8.0, 20.0), size_for_moi=20)
s.set_static(2, (6.6, 2.0, 3), (1.0, 8.0, 20.0), size_for_moi=20)
w = World(s)
w.rebuild_vertices()
dt = np.float32(1 / 60.)
w.snapshot()
for k in range(12):
    w.step(dt)          # eager, then captured graph, then replays
w.restore()
for k in range(3):
    w.step_profiled(dt)
w.set_solver("shuffled")          # throughput sweep order (solve_versioned_kernel<true, false>)
for k in range(3):
    w.step(dt)
w.set_solver("exact")
w.step(np.float32(1 / 50.))       # another dt: the captured graph is replayed, dt comes from device memory
m = w.models()
st = w.stats()
c = w.contacts(); a, b = w.pairs(); d = w.download()
q = scenes.narrowphase_pairs(2048, seed=3)
r = check_collision(q["type"], q["pos_a"], q["verts_a"], q["rad_a"], q["pos_b"], q["verts_b"], q["rad_b"])
# a cube-only world: the two-kernel narrowphase (GJK kernel, deferred stragglers + EPA kernel) instead of the one-kernel form
s2 = scenes.cube_drop(n=400, dims=(8, 7, 8), spacing=1.02, jitter=0.08)
w2 = World(s2)
w2.rebuild_vertices()
for k in range(14):
    w2.step(dt)
st2 = w2.stats()
c2 = w2.contacts()
w2.close()
print("cube-only world ok:", st2, "contacts", len(c2))
print("sanitize run ok:", st, "contacts", len(c), "pairs", len(a), "hits", int(r["hit"].sum()))
w.close()
