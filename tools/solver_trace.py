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
versioned = os.environ.get("NANS_SOLVER", "versioned") not in ("flow", "levels")
c = w.contacts()
print("stage ms", ms, "contacts", n, "levels", st["solver_levels"])
print("type counts", np.bincount(c["type"], minlength=5))
if versioned:
    # slots: fire, stored, ticket, polls
    t0 = t4[:, 2].min()
    fire = (t4[:, 0] - t0).astype(np.float64) / 1e3
    done = (t4[:, 1] - t0).astype(np.float64) / 1e3
    ticket = (t4[:, 2] - t0).astype(np.float64) / 1e3
    cyc = t4[:, 3].astype(np.int64)
    print(f"apply cycles (clock64): med {np.median(cyc):.0f} p10 {np.percentile(cyc, 10):.0f} p90 {np.percentile(cyc, 90):.0f}")
    # hop latency along the deepest chains: fire time of a level-L contact minus the latest fire of level L-1 is
    # not available per edge, so report the slope of the per-level minimum fire time instead
    lvs = np.array(sorted(set(lv.tolist())))
    fmin = np.array([fire[lv == L].min() for L in lvs])
    if len(lvs) > 20:
        k = len(lvs) // 2
        print(f"per-level slope (min fire time): first half {(fmin[k] - fmin[0]) / (lvs[k] - lvs[0]):.2f} us/level, "
              f"second half {(fmin[-1] - fmin[k]) / (lvs[-1] - lvs[k]):.2f} us/level")
    print(f"span us {done.max():.1f}; apply med {np.median(done - fire):.2f} us; wait med {np.median(fire - ticket):.2f} "
          f"p90 {np.percentile(fire - ticket, 90):.2f} max {np.max(fire - ticket):.1f} us; last ticket at {ticket.max():.1f} us")
    idx = np.arange(n)
    for q in range(0, 100, 10):
        m = (idx >= n * q // 100) & (idx < n * (q + 10) // 100)
        print(f"list {q:3d}-{q+10:3d}%: ticket med {np.median(ticket[m]):7.1f} fire med {np.median(fire[m]):7.1f} max {fire[m].max():7.1f}"
              f" | level max {lv[m].max():3d} | wait med {np.median((fire - ticket)[m]):6.2f}")
    for L in sorted(set(lv.tolist())):
        m = lv == L
        if L <= 12 or L % 8 == 0:
            print(f"level {L:3d} n={m.sum():7d} fire min/med/max us = {fire[m].min():8.1f} {np.median(fire[m]):8.1f} {fire[m].max():8.1f}")
    # critical path: walk back from the contact that finished last, always to the predecessor (previous
    # contact on body A or body B) that finished later
    a = c["a"].astype(np.int64).copy(); b = c["b"].astype(np.int64).copy()
    nc_ = w.n_cubes
    ty = c["type"]
    a[(ty == 3) | (ty == 4)] += nc_                       # SS, SF: a is a sphere
    b[(ty == 1) | (ty == 3)] += nc_                       # CS, SS: b is a sphere
    b[(ty == 2) | (ty == 4)] = -1                         # statics are no dependency
    last = {}
    pa = np.full(n, -1, np.int64); pb = np.full(n, -1, np.int64)
    for i in range(n):
        pa[i] = last.get(a[i], -1); last[a[i]] = i
        if b[i] >= 0:
            pb[i] = last.get(b[i], -1); last[b[i]] = i
    head = np.ones(n, bool); head[1:] = a[1:] != a[:-1]
    chunk = (np.cumsum(head) - 1) // 32                   # 32 runs per warp ticket
    cur = int(np.argmax(done)); hops = []; 
    while True:
        cand = [p for p in (pa[cur], pb[cur]) if p >= 0]
        if not cand: break
        p = max(cand, key=lambda k: done[k])
        hops.append((fire[cur] - done[p], done[cur] - fire[cur], chunk[cur] == chunk[p], (p == cur - 1) and not head[cur]))
        cur = int(p)
    h = np.array(hops, dtype=np.float64)
    print(f"critical path: {len(h)} hops, ends at {done.max():.1f} us, starts at {fire[cur]:.1f} us")
    for name, m in (("same lane (run)", h[:, 3] == 1), ("same warp ticket", (h[:, 2] == 1) & (h[:, 3] == 0)), ("other warp", h[:, 2] == 0)):
        if m.any():
            print(f"  {name:17s}: {int(m.sum()):4d} hops, detect latency med {np.median(h[m, 0]):.2f} us (sum {h[m, 0].sum():.1f}), "
                  f"apply med {np.median(h[m, 1]):.2f} us (sum {h[m, 1].sum():.1f})")
    sys.exit(0)
t0 = t4[:, 0].min()
t = (t4[:, 0] - t0).astype(np.float64) / 1e3
apply_us = (t4[:, 1] - t4[:, 0]).astype(np.float64) / 1e3
sync_us = (t4[:, 2] - t4[:, 1]).astype(np.float64) / 1e3
chain = t4[:, 3] == 1
print("span us", t.max())
for L in sorted(set(lv.tolist()))[:200]:
    m = lv == L
    if L <= 20 or L % 10 == 0 or m.sum() > 5000:
        print(f"level {L:3d} n={m.sum():7d} start min/med/max us = {t[m].min():8.1f} {np.median(t[m]):8.1f} {t[m].max():8.1f}"
              f" | apply med {np.median(apply_us[m]):5.2f} sync med {np.median(sync_us[m]):5.2f} chained {chain[m].mean()*100:3.0f}%")
