"""Diagnostic for a 1M-cube parity mismatch: find the first differing contact, dump the bodies involved."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from nans_projekat_b200 import scenes
from nans_projekat_b200.world import World
from oracle import oracle as O
from helpers import world_from_scene
STATE = ("pos", "vel", "force", "ang", "angvel", "torque", "verts")
DT = np.float32(1 / 60.)
side, layers, settle, nsteps = [int(x) for x in sys.argv[1:5]]
s = scenes.cube_pile(n_side=side, layers=layers, seed=7)
gw = World(s); gw.rebuild_vertices()
for _ in range(settle):
    gw.step(DT)
d = gw.download(fields=STATE)
w = world_from_scene(O, s); w.rebuild_vertices()
for f in STATE:
    getattr(w, f)[...] = getattr(d, f)
found = 0
for step in range(nsteps):
    pre = {f: getattr(w, f).copy() for f in ("pos", "verts")}
    gw.upload(w, fields=STATE)
    gw.step(DT)
    oc = w.step(DT, prefilter="grid", cap=8_000_000)
    gc = gw.contacts()
    pa, pb = gw.pairs()
    print("step", step, "gpu", len(gc), "oracle", len(oc), "pairs", len(pa), "stats", gw.stats(strict=False), flush=True)
    if gc.tobytes() != oc.tobytes():
        n = min(len(gc), len(oc))
        key = lambda c: np.stack([c["type"][:n], c["a"][:n], c["b"][:n]], 1)
        diff = np.nonzero((key(gc) != key(oc)).any(1))[0]
        i = int(diff[0]) if len(diff) else n
        print("first index with different (type,a,b):", i)
        for k in range(max(0, i - 2), min(n, i + 3)):
            print(" gpu", k, gc[k], "\n orc", k, oc[k])
        if i < len(oc):
            c = oc[i]
            a, b = int(c["a"]), int(c["b"])
            pset = set(zip(pa.tolist(), pb.tolist()))
            print("oracle contact bodies", a, b, "type", c["type"], "in GPU candidate pairs:", (a, b) in pset)
            np.savez(os.path.join(ROOT, "gpurun_out", f"diag_{side}_{layers}_{step}.npz"), a=a, b=b, contact=oc[i:i+1],
                     pos_a=pre["pos"][a], pos_b=pre["pos"][b], verts_a=pre["verts"][a], verts_b=pre["verts"][b])
        # float mismatches with equal keys
        if not len(diff) and len(gc) == len(oc):
            for fld in ("point_a", "point_b", "n"):
                bad = np.nonzero((gc[fld].view(np.uint32) != oc[fld].view(np.uint32)).any(1))[0]
                print(fld, "differs at", bad[:10])
        found += 1
        if found >= 3:
            break
    dd = gw.download(fields=STATE)
    for f in ("pos", "vel", "ang", "angvel", "verts"):
        if getattr(dd, f).tobytes() != getattr(w, f).tobytes():
            bad = np.nonzero((getattr(dd, f).view(np.uint32) != getattr(w, f).view(np.uint32)).reshape(len(dd.pos), -1).any(1))[0]
            print("  state", f, "differs on", len(bad), "bodies, first", bad[:8])
gw.close()
