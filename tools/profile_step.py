"""Profiling driver: prepare the bench workload, then run `steps` steps between cudaProfilerStart/Stop
(use with `ncu --profile-from-start off`).  usage: python tools/profile_step.py [bodies] [side] [settle] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from nans_projekat_b200 import scenes
from nans_projekat_b200.world import World

bodies = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
side = int(sys.argv[2]) if len(sys.argv) > 2 else 250
settle = int(sys.argv[3]) if len(sys.argv) > 3 else 45
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
layers = max(1, (bodies + side * side - 1) // (side * side))
s = scenes.cube_pile(n_side=side, layers=layers, n=bodies, seed=7)
w = World(s)
w.rebuild_vertices()
dt = np.float32(1 / 60.)
for _ in range(settle):
    w.step(dt)
w.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(steps):
    w.step(dt)
w.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print(w.stats())
