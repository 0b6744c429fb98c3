#!/usr/bin/env python
"""CPU model of the narrowphase's lane utilisation: per-pair GJK / EPA iteration counts from the oracle, grouped into
32-pair chunks the way the kernels take them, -> mean / chunk-maximum (a chunk runs as long as its slowest lane).
Cost-weighted: an EPA iteration over a polytope of k iterations' age costs ~ (5 + k) (faces grow by ~2 per iteration).
usage: python tools/np_lane_model.py [pairs=400000]      (TEST/ANALYSIS infrastructure: uses oracle/)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nans_projekat_b200 import scenes
from oracle import oracle as O


def util(x, weight=None):
    n = len(x) // 32 * 32
    if n == 0:
        return float("nan"), float("nan")
    ch = x[:n].reshape(-1, 32)
    w = weight or (lambda k: k)
    return ch.mean() / ch.max(1).mean(), w(ch).mean() / w(ch.max(1)).mean()


def report(name, p):
    r = O.check_collision_batch(p["type"], p["pos_a"], p["verts_a"], p["rad_a"], p["pos_b"], p["verts_b"], p["rad_b"], want_stats=True)
    st, t = r["stats"], p["type"]
    for tn, tt in (("CC", 0), ("CS", 1), ("CF", 2)):
        m = t == tt
        if not m.any():
            continue
        found = st["gjk_result"][m] == 1
        gi, ei = st["gjk_iters"][m], st["epa_iters"][m][found]
        ug, _ = util(gi)
        ue, uc = util(ei, lambda k: k * (5 + k))
        print(f"{name} {tn}: {m.sum()} pairs, {100 * found.mean():.0f} % intersect | GJK evolutions mean {gi.mean():.1f}, "
              f"lanes busy {100 * ug:.0f} % | EPA iterations mean {ei.mean() if len(ei) else 0:.1f} p99 "
              f"{np.percentile(ei, 99) if len(ei) else 0:.0f} max {ei.max() if len(ei) else 0}, lanes busy over the compacted hits "
              f"{100 * ue:.0f} % (cost-weighted {100 * uc:.0f} %)")


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
    report("C3 random pairs", scenes.narrowphase_pairs(n, seed=1234))
    # settled pile: candidate pairs of a small pile stepped by the oracle
    s = scenes.cube_pile(n_side=14, layers=16, seed=7)
    ow = O.World(s.n_cubes, s.n_spheres, s.n_statics)
    for f in s.ARRAYS:
        getattr(ow, f)[...] = getattr(s, f)
    ow.rebuild_vertices()
    for _ in range(46):
        ow.step(np.float32(1 / 60.), prefilter=True)
    v = ow.verts; lo, hi = v.min(1), v.max(1)
    pa, pb = [], []
    for i in range(len(v)):
        j = np.nonzero(np.all((lo[i] <= hi[i + 1:]) & (lo[i + 1:] <= hi[i]), axis=1))[0] + i + 1
        pa.append(np.full(len(j), i)); pb.append(j)
    pa, pb = np.concatenate(pa), np.concatenate(pb)
    z = np.zeros(len(pa), np.float32)
    report("settled pile (14x14x16)", dict(type=np.zeros(len(pa), np.int32), pos_a=ow.pos[pa].copy(), verts_a=v[pa].copy(), rad_a=z,
                                           pos_b=ow.pos[pb].copy(), verts_b=v[pb].copy(), rad_b=z))
