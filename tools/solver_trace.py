"""Debug: timeline of the exact-order solver (run with NANS_SOLVER_TRACE=1).
usage: NANS_SOLVER_TRACE=1 python tools/solver_trace.py [bodies side settle]"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nans_projekat_b200 import scenes, _lib
from nans_projekat_b200.world import World
bodies = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
side = int(sys.argv[2]) if len(sys.argv) > 2 else 250
settle = int(sys.argv[3]) if len(sys.argv) > 3 else 45
layers = max(1, (bodies + side * side - 1) // (side * side))
w = World(scenes.cube_pile(n_side=side, layers=layers, n=bodies, seed=7)); w.rebuild_vertices()
dt = np.float32(1 / 60.)
for _ in range(settle): w.step(dt)
ms = w.step_profiled(dt); st = w.stats(); n = st["n_contacts"]
t4 = np.zeros((n, 4), np.uint64); lv = np.zeros(n, np.int32)
_lib.check(_lib.lib().nans_debug_solver_trace(w._h, t4.ctypes.data, lv.ctypes.data, n))
t0 = t4[:, 0].min()
t = (t4[:, 0] - t0).astype(np.float64) / 1e3
apply_us = (t4[:, 1] - t4[:, 0]).astype(np.float64) / 1e3
sync_us = (t4[:, 2] - t4[:, 1]).astype(np.float64) / 1e3
chain = t4[:, 3] == 1
print("stage ms", ms, "contacts", n, "levels", st["solver_levels"], "span us", t.max())
c = w.contacts()
print("type counts", np.bincount(c["type"], minlength=5))
for L in sorted(set(lv.tolist()))[:200]:
    m = lv == L
    if L <= 20 or L % 10 == 0 or m.sum() > 5000:
        print(f"level {L:3d} n={m.sum():7d} start min/med/max us = {t[m].min():8.1f} {np.median(t[m]):8.1f} {t[m].max():8.1f}"
              f" | apply med {np.median(apply_us[m]):5.2f} sync med {np.median(sync_us[m]):5.2f} chained {chain[m].mean()*100:3.0f}%")
