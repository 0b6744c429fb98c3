"""Shared helpers for the parity tests."""
import numpy as np


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_bit_equal(a, b, what=""):
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    bad = bits(a) != bits(b)
    # +0/-0 and NaN payloads are distinct bit patterns; report them as mismatches too
    assert not bad.any(), f"{what}: {int(bad.sum())} of {bad.size} floats differ, first at " \
                          f"{np.argwhere(bad)[0]}: {a[bad][0]!r} vs {b[bad][0]!r}"


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    scale = np.maximum(np.maximum(np.abs(a), np.abs(b)), 1.0)
    return float(np.max(np.abs(a - b) / scale)) if a.size else 0.0


def world_from_scene(O, s):
    """oracle.World holding the same arrays as a product Scene (copies)."""
    w = O.World(s.n_cubes, s.n_spheres, s.n_statics)
    for f in s.ARRAYS:
        getattr(w, f)[...] = getattr(s, f)
    return w


def load_stage_records(path):
    z = np.load(path)
    n = int(z["n"])
    recs = []
    for i in range(n):
        rec = {}
        pre = f"{i}_"
        for k in z.files:
            if k.startswith(pre):
                parts = k[len(pre):].split("_", 1)
                if parts[0] in ("s0", "s1", "s2", "s3"):
                    rec.setdefault(parts[0], {})[parts[1]] = z[k]
                else:
                    rec[k[len(pre):]] = z[k]
        recs.append(rec)
    return recs, np.float32(z["dt"])


def world_from_stage_record(O, rec, snap="s0"):
    nc, ns = int(rec["nc"]), int(rec["ns"])
    w = O.World(nc, ns, 1)
    for k in ("pos", "vel", "force", "ang", "angvel", "torque", "verts"):
        getattr(w, k)[...] = rec[snap][k]
    for k in ("mass", "moi", "radius", "scale", "st_pos", "st_scale", "st_mass", "st_moi", "st_verts"):
        getattr(w, k)[...] = rec[k]
    return w
