"""Exploration helper: step a cube pile on the GPU and print stage timings + counters.
usage: python tools/explore_pile.py n_side layers steps [print_every]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nans_projekat_b200 import scenes
from nans_projekat_b200.world import World, kernel_launches

n_side, layers, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
every = int(sys.argv[4]) if len(sys.argv) > 4 else 10
t0 = time.time()
s = scenes.cube_pile(n_side=n_side, layers=layers)
print(f"scene {s.n_cubes} cubes built in {time.time()-t0:.1f}s; arena {World.arena_bytes(s)/2**30:.2f} GiB", flush=True)
w = World(s)
w.rebuild_vertices()
dt = np.float32(1/60.)
for k in range(steps):
    ms = w.step_profiled(dt)
    if k % every == 0 or k == steps - 1:
        st = w.stats(strict=False)
        d = w.download(fields=("pos",))
        print(f"step {k:4d} " + " ".join(f"{a[:5]}={b:7.3f}" for a, b in ms.items()) +
              f" | pairs {st['n_pairs']} contacts {st['n_contacts']} levels {st['solver_levels']} ovf {st['overflow']}"
              f" epaF {st['max_epa_faces']} | y[min,mean,max]={d.pos[:,1].min():.2f},{d.pos[:,1].mean():.2f},{d.pos[:,1].max():.2f}"
              f" nan={int(np.isnan(d.pos).any())}", flush=True)
print("launches", kernel_launches())
