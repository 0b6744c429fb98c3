"""Generate tests/golden/*.npz from the reference's own prebuilt build/nans.so.

Run in the authoring container (where /root/reference exists):
    python tests/golden/make_golden.py
The fixtures pin the CPU restatement (oracle/) and the CUDA path on boxes where the
reference binary is absent.  Everything here is OUTPUT OF THE REFERENCE BINARY; inputs are
stored explicitly so no RNG/libm difference can change them.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from nans_projekat_b200 import scenes  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
H = O.ref()
assert H is not None, "reference binary not available"


def scene_to_ref_state(s):
    """Write a <=16+16+1 Scene into a reference sdl_state (Model rebuilt from pose)."""
    st = O.ref_new_state()
    r = st[0]
    nc, ns = s.n_cubes, s.n_spheres
    r["CubeCount"], r["SphereCount"] = nc, ns
    for i in range(nc):
        c = r["Cubes"][i]
        c["Position"], c["V"], c["Forces"] = s.pos[i], s.vel[i], s.force[i]
        c["Angles"], c["W"], c["Torque"] = s.ang[i], s.angvel[i], s.torque[i]
        c["Size"], c["Mass"], c["MOI"] = s.scale[i][0], s.mass[i], s.moi[i]
        m, v = O.model_vertices(s.pos[i], s.ang[i], s.scale[i])
        c["Model"], c["Vertices"] = m.reshape(16), v
    for j in range(ns):
        b = nc + j
        c = r["Spheres"][j]
        c["Position"], c["V"], c["Forces"] = s.pos[b], s.vel[b], s.force[b]
        c["Angles"], c["W"], c["Torque"] = s.ang[b], s.angvel[b], s.torque[b]
        c["Radius"], c["Mass"], c["MOI"] = s.radius[b], s.mass[b], s.moi[b]
    f = r["Floor"]
    f["Position"], f["Angles"] = s.st_pos[0], s.st_ang[0]
    f["Size"], f["Mass"], f["MOI"] = s.st_scale[0][0], s.st_mass[0], s.st_moi[0]
    m, v = O.model_vertices(s.st_pos[0], s.st_ang[0], s.st_scale[0])
    f["Model"], f["Vertices"] = m.reshape(16), v
    return st


def snap(st):
    w = O.world_from_ref_state(st)
    return {k: getattr(w, k).copy() for k in ("pos", "vel", "force", "ang", "angvel", "torque", "verts")}


def gen_narrowphase():
    per = 300
    types = np.concatenate([np.full(per, t) for t in (0, 1, 2, 3, 4)]).astype(np.int32)
    out = {}
    for tag, rot in (("rot", True), ("axis", False)):
        p = scenes.narrowphase_pairs(len(types), seed=4321 + rot, rotated=rot, types=types)
        r = O.ref_check_collision_batch(p["type"], p["pos_a"], p["verts_a"], p["rad_a"],
                                        p["pos_b"], p["verts_b"], p["rad_b"])
        for k, v in p.items():
            out[f"{tag}_{k}"] = v
        for k, v in r.items():
            out[f"{tag}_ref_{k}"] = v
        print("narrowphase", tag, "hit rate per type",
              [float(r["hit"][types == t].mean()) for t in range(5)])
    np.savez_compressed(os.path.join(OUT, "narrowphase.npz"), **out)


def gen_stages():
    rng = np.random.default_rng(20261017)
    dt = np.float32(1 / 60)
    recs = []
    for trial in range(48):
        nc, ns = int(rng.integers(1, 17)), int(rng.integers(0, 17))
        s = scenes.random_small_world(rng, nc, ns, spread=float(rng.choice([1.0, 2.0, 3.0])))
        st = scene_to_ref_state(s)
        rec = {"nc": nc, "ns": ns, "mass": s.mass, "moi": s.moi, "radius": s.radius, "scale": s.scale,
               "st_pos": s.st_pos, "st_scale": s.st_scale, "st_mass": s.st_mass, "st_moi": s.st_moi,
               "st_verts": st[0]["Floor"]["Vertices"].copy()[None]}
        rec["s0"] = snap(st)
        H.nansref_integrate_forces(O._sp(st), C.c_float(dt)); rec["s1"] = snap(st)
        pr = O.ref_detect(st, dt); rec["contacts"] = O.contacts_from_ref_pairs(pr)
        H.nansref_solve_constraints(O._sp(st), C.c_float(dt), pr.ctypes.data_as(C.c_void_p), len(pr))
        rec["s2"] = snap(st)
        H.nansref_integrate_velocities(O._sp(st), C.c_float(dt)); rec["s3"] = snap(st)
        recs.append(rec)
    flat = {"n": np.int32(len(recs)), "dt": dt}
    for i, rec in enumerate(recs):
        for k, v in rec.items():
            if isinstance(v, dict):
                for kk, vv in v.items():
                    flat[f"{i}_{k}_{kk}"] = vv
            else:
                flat[f"{i}_{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(OUT, "stages.npz"), **flat)
    print("stages", len(recs), "worlds, contacts", sum(len(r["contacts"]) for r in recs))


def gen_demo():
    """Config C1: Init scene through the exported SimUpdateAndRender, scripted shot
    (SURVEY.md §8d): dt=0 on step 0 then 1/60; aim at step 200, shoot at step 201."""
    perm = np.zeros(1 << 20, np.uint8)
    st = perm[:O.STATE_DTYPE.itemsize].view(O.STATE_DTYPE)
    mem = O.MemoryStruct(perm.ctypes.data, perm.nbytes, perm.ctypes.data, perm.nbytes, 0)
    inp = np.zeros(1, O.INPUT_DTYPE)
    ren = np.zeros(256, np.uint8)
    steps = 1000
    keys = ("pos", "vel", "ang", "angvel", "verts")
    traj = {k: [] for k in keys}
    ncontacts, cam = [], []
    for k in range(steps):
        inp[0]["Sensitivity"] = 0.5
        inp[0]["Buttons"][4][1] = 1 if k == 201 else 0
        if k == 200:
            inp[0]["XRel"], inp[0]["YRel"] = 127, -44
        dt = 0.0 if k == 0 else 1 / 60.
        H.nansref_sim_update_and_render(C.byref(mem), O._sp(inp), ren.ctypes.data_as(C.c_void_p), C.c_float(dt))
        # refresh Vertices from the Model the draw section just rebuilt (what the next frame's
        # IntegrateForces would do, code/nans.cpp:995): state = end-of-frame pose + its vertices
        for i in range(4):
            H.nansref_update_vertices(O._sp(st), i)
        H.nansref_floor_update_vertices(O._sp(st))
        sn = snap(st)
        for kk in keys:
            traj[kk].append(sn[kk])
        pairs = st[0]["Pairs"]
        ncontacts.append(int((pairs[1] - pairs[0]) // 96))
        c = st[0]["Camera"]
        cam.append(np.concatenate([c["Position"], c["Front"], [c["Yaw"], c["Pitch"]]]))
    out = {k: np.stack(v) for k, v in traj.items()}
    out["ncontacts"] = np.array(ncontacts, np.int32)
    out["camera"] = np.stack(cam).astype(np.float32)
    out["floor_verts"] = st[0]["Floor"]["Vertices"].copy()
    np.savez_compressed(os.path.join(OUT, "demo_traj.npz"), **out)
    print("demo: final cube y", out["pos"][-1][:4, 1], "contacts last", ncontacts[-1],
          "max contacts", max(ncontacts))


def gen_epa_emptied(cases="/tmp/emptied_cases.npz"):
    """EPA runs that empty the triangle list (the reference then reads the stale Triangle[0]): inputs found by
    tests/golden/find_emptied_cases.py (or the previously committed fixture), outputs = nans.so's CheckCollision."""
    dst = os.path.join(OUT, "epa_emptied.npz")
    z = np.load(cases if os.path.exists(cases) else dst)
    n = len(z["pos_a"])
    p = dict(type=np.zeros(n, np.int32), pos_a=z["pos_a"].astype(np.float32), verts_a=z["verts_a"].astype(np.float32),
             rad_a=np.zeros(n, np.float32), pos_b=z["pos_b"].astype(np.float32), verts_b=z["verts_b"].astype(np.float32),
             rad_b=np.zeros(n, np.float32))
    r = O.ref_check_collision_batch(p["type"], p["pos_a"], p["verts_a"], p["rad_a"], p["pos_b"], p["verts_b"], p["rad_b"])
    o = O.check_collision_batch(p["type"], p["pos_a"], p["verts_a"], p["rad_a"], p["pos_b"], p["verts_b"], p["rad_b"],
                                want_stats=True)
    assert (o["stats"]["emptied"] > 0).all(), "every stored case must empty the triangle list"
    np.savez_compressed(dst, **p, ref_hit=r["hit"], ref_N=r["N"], ref_PA=r["PA"], ref_PB=r["PB"])
    print("epa_emptied:", n, "pairs,", int(r["hit"].sum()), "hits in the reference binary")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "epa_emptied":
        gen_epa_emptied()
        sys.exit(0)
    gen_narrowphase()
    gen_stages()
    gen_demo()
    gen_epa_emptied()
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")
