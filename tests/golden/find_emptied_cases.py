"""Search for EPA runs that EMPTY the reference's triangle list (every face dissolved, every horizon edge cancelled;
the next iteration then reads the stale Triangle[0], code/nans.cpp:807-866) among the neighbour pairs of the settling
100^3 cube pile (config C5), using the oracle's `emptied` statistic.  ~5 minutes of CPU.  Writes the INPUT shapes to
/tmp/emptied_cases.npz; `make_golden.py epa_emptied` then runs the reference binary on them.  The first case is the pair
the 1 M-cube GPU parity test of round 2 caught (cubes 40349 / 50349, step 81)."""
import time, numpy as np, sys
sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo')
from nans_projekat_b200 import scenes
from oracle import oracle as O
from helpers import world_from_scene
s=scenes.cube_pile(n_side=100, layers=100, seed=7)
w=world_from_scene(O,s); w.rebuild_vertices()
dt=np.float32(1/60.)
found=[]
import os
if os.path.exists('/root/repo/gpurun_out/diag_100_100_1.npz'):
    z=np.load('/root/repo/gpurun_out/diag_100_100_1.npz')
    found.append((z['pos_a'],z['verts_a'],z['pos_b'],z['verts_b']))
for k in range(100):
    w.step(dt,prefilter="grid", cap=16_000_000)
    if k>=50 and k%4==0:
        n=w.nb
        i=np.arange(n)
        for off in (1,100,10000):
            a=i[:n-off]; b=a+off
            # keep only plausible neighbours
            m=np.abs(w.pos[a]-w.pos[b]).max(1)<1.2
            a=a[m]; b=b[m]
            zf=np.zeros(len(a),np.float32)
            r=O.check_collision_batch(np.zeros(len(a),np.int32), w.pos[a], w.verts[a], zf, w.pos[b], w.verts[b], zf, want_stats=True)
            e=np.nonzero(r['stats']['emptied']>0)[0]
            for q in e:
                found.append((w.pos[a[q]].copy(), w.verts[a[q]].copy(), w.pos[b[q]].copy(), w.verts[b[q]].copy()))
            print(k, off, len(a), 'emptied', len(e), 'hits among them', int(r['hit'][e].sum()), flush=True)
np.savez('/tmp/emptied_cases.npz', pos_a=np.array([f[0] for f in found]), verts_a=np.array([f[1] for f in found]), pos_b=np.array([f[2] for f in found]), verts_b=np.array([f[3] for f in found]))
print(len(found))
