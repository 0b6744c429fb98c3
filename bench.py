#!/usr/bin/env python
"""bench.py — body-steps/s of the rigid-body step (BASELINE.json metric) on N B200s.

    python bench.py --gpus 1 --steps 20 --warmup 5            # our arm
    python bench.py --impl reference --steps 3 --warmup 1      # the reference's CPU path
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of IntegrateForces -> DetectCollisions (broadphase + GJK/EPA) ->
SolveConstraints -> IntegrateVelocities (+ vertex rebuild) over the whole world
(reference code/nans.cpp:1758-1762).  Workload at N=1: the 1M-cube pile (BASELINE.json metric:
"body-steps/sec at 1M cubes") in SURVEY.md 8(d)'s shape, 100 x 100 x 100 cubes, one world resident on
one GPU; the flat 250 x 250 x 16 pile of round 1 is timed beside it (`alt_shapes`).  The same line
carries the second half of the metric (config C3: GJK+EPA pairs/s on 16 Mi random pairs, flags
checked against the reference binary), configs C2 and C4 as sub-records, and an in-run parity check
of the headline state against the CPU oracle.  At N>1 every rank steps its own independent 1M-cube
world (north_star: "independent batched worlds ... shard embarrassingly across GPUs with no
communication"): weak scaling, no data-path collective; the `slab` sub-record is ONE world of
N x 1M cubes in spatial x-slabs with an NCCL halo exchange (config C5).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DT = np.float32(1 / 60.)
# ALGORITHMIC bytes per unit (restated in DESIGN.md §5)
BYTES = {"integrate_forces": 104, "integrate_velocities": 168, "aabb_key": 128, "radix_sort": 64,
         "pair_emit": 8, "narrowphase": 264, "solver": 184}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bodies", type=int, default=1_000_000, help="cubes per world (per rank)")
    ap.add_argument("--side", type=int, default=100, help="pile footprint: side x side cubes per layer "
                                                           "(100 -> 100x100x100, SURVEY 8d; 250 -> 250x250x16)")
    ap.add_argument("--settle", type=int, default=-1, help="untimed scene-preparation steps before warm-up "
                                                            "(-1: 80 for the 100^3 pile, 40 otherwise)")
    ap.add_argument("--window", type=int, default=20, help="steps between snapshot restores")
    ap.add_argument("--cpu-bodies", type=int, default=2048, help="size of the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--workload", default="pile", choices=["pile", "drop10k", "worlds4096"],
                    help="pile = 1M-cube pile (BASELINE metric; default); drop10k = config C2; worlds4096 = config C4")
    ap.add_argument("--c3", type=int, default=-1, metavar="PAIRS",
                    help="GJK+EPA microbench (config C3) on this many random pairs; -1 = 16777216 at N=1, off at N>1; 0 = off")
    ap.add_argument("--c3-check", type=int, default=1 << 20, help="pairs of C3 whose flags are checked on the CPU")
    ap.add_argument("--no-subrecords", action="store_true", help="skip alt shape, C2, C4 and the in-run parity check")
    a = ap.parse_args()
    if a.settle < 0:
        a.settle = 80 if (a.workload == "pile" and a.side == 100) else 40
    return a


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t_begin=None, t_end=None):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        rows = [r for (ts, r) in self.rows if t_begin is None or (t_begin - 0.05 <= ts <= t_end + 0.15)]
        if not rows:
            rows = [r for (_, r) in self.rows]
        for r in rows:
            try:
                sm.append(float(r[1])); mx = max(mx, float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa(gpu_index):
    """Pin this process to the CPUs local to its GPU (PCIe root / NUMA node) BEFORE any pinned buffer is allocated, so
    that host staging memory is first-touched on that node.  Round 1 left all 8 ranks on node 0 and lost 22 % of the
    end-to-end throughput at N = 8 to cross-socket copies.  Returns what was done (for the JSON line)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        path = f"/sys/bus/pci/devices/{dom[-4:].lower()}:{rest.lower()}"
        cpus = open(os.path.join(path, "local_cpulist")).read().strip()
        node = open(os.path.join(path, "numa_node")).read().strip()
        ids = set()
        for part in cpus.split(","):
            if "-" in part:
                a, b = part.split("-"); ids.update(range(int(a), int(b) + 1))
            elif part:
                ids.add(int(part))
        ids &= os.sched_getaffinity(0)
        if ids:
            os.sched_setaffinity(0, ids)
        return {"gpu": gpu_index, "pci": bus, "numa_node": node, "cpus": cpus, "bound": bool(ids)}
    except Exception as ex:
        return {"gpu": gpu_index, "bound": False, "why": f"{type(ex).__name__}: {ex}"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def build_pile(bodies, side, seed):
    from nans_projekat_b200 import scenes
    layers = max(1, (bodies + side * side - 1) // (side * side))
    return scenes.cube_pile(n_side=side, layers=layers, n=bodies, seed=seed), layers


def build_workload(args, seed):
    """(scene, layers, name): the scene the step is timed on."""
    from nans_projekat_b200 import scenes
    if args.workload == "drop10k":      # config C2: 10 000 cubes dropped into the static box
        return scenes.cube_drop(n=10000, seed=seed), 21, "cube_drop_10k (config C2)"
    if args.workload == "worlds4096":   # config C4: 4096 independent worlds x (48 cubes + 16 spheres)
        return scenes.batched_worlds(4096, 48, 16, seed=seed), 3, "batched_worlds_4096x64 (config C4)"
    scene, layers = build_pile(args.bodies, args.side, seed)
    return scene, layers, ("cube_pile_1M" if scene.n_cubes == 1_000_000 else f"cube_pile_{scene.n_cubes}") + f"_{args.side}x{layers}x{args.side}"


# --------------------------------------------------------------------------------------- CPU legs
def cpu_port_baseline(state_scene, n_sample, budget_s=15.0):
    """The oracle port (same all-pairs algorithm as the reference, 1 thread, -O2) on a bounded
    sample of the SAME workload: the first n_sample cubes of the settled pile + the floor."""
    from oracle import oracle as O
    n = min(n_sample, state_scene.n_cubes)
    w = O.World(n, 0, state_scene.n_statics)
    for f in ("pos", "vel", "force", "ang", "angvel", "torque", "scale"):
        getattr(w, f)[...] = getattr(state_scene, f)[:n]
    w.mass[...] = state_scene.mass[:n]; w.moi[...] = state_scene.moi[:n]
    w.verts[...] = state_scene.verts[:n]
    for f in ("st_pos", "st_ang", "st_scale", "st_mass", "st_moi", "st_verts"):
        getattr(w, f)[...] = getattr(state_scene, f)
    steps, t0 = 0, time.perf_counter()
    while True:
        w.step(DT, prefilter=False)
        steps += 1
        el = time.perf_counter() - t0
        if el > budget_s or steps >= 200:
            break
    full = state_scene.n_cubes
    return {"value": n * steps / el, "unit": "body-steps/s", "cores": 1, "kind": "port",
            "sample": f"first {n} cubes of the settled pile + floor, {steps} all-pairs steps "
                      f"(reference algorithm, O(N^2) GJK, oracle/nans_oracle.c -O2, 1 thread), {el:.1f} s",
            # the all-pairs step costs ~N^2, so body-steps/s falls as 1/N; NOT measured, the full
            # workload (5e11 pair tests per step at 1 M cubes) is not runnable on a CPU
            "extrapolated_to_workload": {"value": n * steps / el * n / full, "unit": "body-steps/s", "bodies": full,
                                         "how": f"measured value x {n} / {full} (extrapolated, O(N^2) per step)"}}


def run_reference(args):
    """Reference arm: the reference's OWN prebuilt plugin (oracle/_ref/nans.so, unmodified, -O0 as
    shipped) stepping its physics stage functions, single thread (the reference has no threads), at
    its hard cap of 16 cubes (MAX_CUBE_COUNT, code/nans.h:52): a 16-cube column sample of the pile.
    Falls back to the oracle port when the binary was not staged."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import ctypes as C
    from oracle import oracle as O
    from nans_projekat_b200 import scenes
    H = O.ref()
    # the same workload name as our arm's line (the reference can only run a sample of it: `sample` says which)
    layers = max(1, (args.bodies + args.side * args.side - 1) // (args.side * args.side))
    cfg = {"workload": ("cube_pile_1M" if args.bodies == 1_000_000 else f"cube_pile_{args.bodies}") + f"_{args.side}x{layers}x{args.side}",
           "bodies_per_world": args.bodies, "l2": "n/a (CPU)",
           "same_config": False,   # the reference binary is capped at 16 cubes (MAX_CUBE_COUNT, code/nans.h:52): it runs a SAMPLE
           }
    # 16-cube sample: a 2x2 footprint, 4 layers of the same lattice, settled on the floor
    s = scenes.cube_pile(n_side=2, layers=4, seed=7)
    window = 25     # world steps per window; every window restarts from the same prepared state
    if H is not None:
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        from make_golden import scene_to_ref_state
        st = scene_to_ref_state(s)
        kind, cores = "reference", 1
        what = "oracle/_ref/nans.so (the reference's shipped -O0 build), stage functions code/nans.cpp:1758-1762"

        def step():
            """one world step; returns the seconds spent INSIDE the reference binary"""
            t = time.perf_counter()
            H.nansref_physics_step(O._sp(st), C.c_float(float(DT)), None, 0)
            t = time.perf_counter() - t
            r = st[0]   # the draw section's Model rebuild (untimed here; not part of :1758-1762)
            for i in range(int(r["CubeCount"])):
                c = r["Cubes"][i]
                c["Model"] = O.model_vertices(c["Position"], c["Angles"], (c["Size"],) * 3)[0].reshape(16)
            return t
        save = lambda: st.copy()
        def restore(x): st[...] = x
    else:
        w = O.World(s.n_cubes, 0, 1)
        for f in s.ARRAYS:
            getattr(w, f)[...] = getattr(s, f)
        w.rebuild_vertices()
        kind, cores, what = "port", 1, "oracle/nans_oracle.c (-O2 restatement; reference binary not staged)"

        def step():
            t = time.perf_counter()
            w.step(DT, prefilter=False)
            return time.perf_counter() - t
        save = lambda: w.copy()
        def restore(x):
            for f in ("pos", "vel", "force", "ang", "angvel", "torque", "verts"):
                getattr(w, f)[...] = getattr(x, f)
    n = s.n_cubes
    for _ in range(40):          # same scene preparation as our arm: let the column come into contact
        step()
    prepared = save()
    inner = 40      # windows per bench step: one bench "step" = inner*window world steps (bounded sample)
    for _ in range(max(args.warmup, 1)):
        restore(prepared)
        for _ in range(window):
            step()
    el = 0.0
    for _ in range(args.steps):
        for _ in range(inner):
            restore(prepared)
            for _ in range(window):
                el += step()
    inner = inner * window
    value = n * inner * args.steps / el
    line = {"impl": "reference", "metric": "body-steps/s", "value": value, "unit": "body-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(cfg, sample=f"{n}-cube column of the pile lattice + floor (the reference's "
                                       f"MAX_CUBE_COUNT cap), {inner} world steps per bench step"),
            "cpu_baseline": {"value": value, "unit": "body-steps/s", "cores": cores, "kind": kind,
                             "sample": f"{what}; {n} cubes + floor, {inner * args.steps} steps, {el:.1f} s"},
            "e2e": {"value": value, "unit": "body-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "host": {"nproc": os.cpu_count()}}
    if not args.no_subrecords and args.workload == "pile":
        try:
            line["same_workload_grid_port"] = cpu_same_workload(args)
        except Exception as ex:   # an extra record must never lose the reference line
            line["same_workload_grid_port"] = {"unavailable": f"{type(ex).__name__}: {ex}"}
    print(json.dumps(line), flush=True)


def cpu_same_workload(args, timed_steps=3):
    """The arm's OWN workload (the whole pile, same seed, same settle steps) on the host cores: the oracle port's
    arithmetic behind its uniform-grid prefilter.  The reference cannot run it (16-cube cap, all-pairs loop), so this is
    NOT the reference arm's `value`: it is the same-config CPU figure to read beside it -- a CPU given this repo's
    candidate search.  GJK+EPA on min(nproc, 32) threads, grid / solver / integrators on one (oracle/nans_oracle.c)."""
    from oracle import oracle as O
    scene, layers = build_pile(args.bodies, args.side, seed=7)
    w = O.World(scene.n_cubes, scene.n_spheres, scene.n_statics)
    for f in scene.ARRAYS:
        getattr(w, f)[...] = getattr(scene, f)
    w.rebuild_vertices()
    cap = max(8 * scene.nb, 1 << 20)
    for _ in range(args.settle):                            # the same preparation as our arm (prepare_world)
        w.step(DT, prefilter="grid", cap=cap)
    t0, contacts = time.perf_counter(), 0
    for _ in range(timed_steps):
        contacts += len(w.step(DT, prefilter="grid", cap=cap))
    el = time.perf_counter() - t0
    return {"value": scene.nb * timed_steps / el, "unit": "body-steps/s", "bodies": scene.nb,
            "shape": f"{args.side}x{layers}x{args.side}", "cores": min(os.cpu_count() or 1, 32),
            "steps": timed_steps, "settle_steps": args.settle, "seconds_per_step": el / timed_steps,
            "contacts_per_step": contacts / timed_steps,
            "what": "oracle port + uniform-grid prefilter on the whole pile (not the reference's O(N^2) all-pairs loop)"}


def _c3_pairs_torch(n, seed, dev):
    """Config C3 inputs generated ON THE DEVICE (synthetic input, not the product): strata CC:CS:SS = 8:7:1, shape A
    at the origin, shape B centre U(-1.2,1.2)^3, unit cubes with Euler angles U(-pi,pi) (rotations in fp64, vertices
    rounded to fp32: world-space vertices are the input), radii U(0.1,0.5)."""
    import torch
    g = torch.Generator(device=dev); g.manual_seed(seed)
    n_cc, n_cs = int(round(n * 8 / 16)), int(round(n * 7 / 16))
    t = torch.cat([torch.full((n_cc,), 0), torch.full((n_cs,), 1), torch.full((n - n_cc - n_cs,), 3)]).to(torch.int32).to(dev)
    corners = torch.tensor([[.5, .5, .5], [.5, .5, -.5], [-.5, .5, .5], [-.5, .5, -.5],
                            [.5, -.5, .5], [.5, -.5, -.5], [-.5, -.5, .5], [-.5, -.5, -.5]], dtype=torch.float64, device=dev)
    out = {"type": t}
    pos_b = (torch.rand((n, 3), generator=g, device=dev, dtype=torch.float64) * 2.4 - 1.2).float()
    for side, pos in (("a", torch.zeros((n, 3), device=dev)), ("b", pos_b)):
        verts = torch.empty((n, 8, 3), dtype=torch.float32, device=dev)
        for o in range(0, n, 1 << 21):
            m = min(1 << 21, n - o)
            ang = (torch.rand((m, 3), generator=g, device=dev, dtype=torch.float64) * 2 - 1) * np.pi
            c, s_ = torch.cos(ang), torch.sin(ang)
            one, zero = torch.ones(m, device=dev, dtype=torch.float64), torch.zeros(m, device=dev, dtype=torch.float64)
            rx = torch.stack([one, zero, zero, zero, c[:, 0], -s_[:, 0], zero, s_[:, 0], c[:, 0]], 1).view(m, 3, 3)
            ry = torch.stack([c[:, 1], zero, s_[:, 1], zero, one, zero, -s_[:, 1], zero, c[:, 1]], 1).view(m, 3, 3)
            rz = torch.stack([c[:, 2], -s_[:, 2], zero, s_[:, 2], c[:, 2], zero, zero, zero, one], 1).view(m, 3, 3)
            r = rx @ ry @ rz
            verts[o:o + m] = (torch.einsum("nij,kj->nki", r, corners) + pos[o:o + m, None, :].double()).float()
        rad = (torch.rand(n, generator=g, device=dev, dtype=torch.float64) * 0.4 + 0.1).float()
        out["posrad_" + side] = torch.cat([pos, rad[:, None]], 1).contiguous()
        out["verts_" + side] = verts
    return out


def _ref_check_chunk(args):
    """worker (separate process): CheckCollision of one chunk through the reference's own nans.so, else the oracle port"""
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    t, pa, va, ra, pb, vb, rb = args
    if O.ref() is not None:
        return "reference", O.ref_check_collision_batch(t, pa, va, ra, pb, vb, rb)["hit"]
    return "port", O.check_collision_batch(t, pa, va, ra, pb, vb, rb)["hit"]


def run_c3(n_pairs, device, n_check, reps=3):
    """Config C3: GJK+EPA pairs/s on random cube/sphere pairs, device-resident inputs, CUDA events.  The hit flags of
    an evenly spaced subsample of `n_check` pairs are compared with CheckCollision of the reference's own binary
    (oracle/_ref/nans.so, code/nans.cpp:907-966) run on the host cores in worker processes."""
    import ctypes as C
    import torch
    from concurrent.futures import ProcessPoolExecutor
    import multiprocessing as mp
    from nans_projekat_b200 import _lib
    dev = torch.device("cuda", device)
    d = _c3_pairs_torch(n_pairs, 1234, dev)
    hit = torch.empty(n_pairs, dtype=torch.int32, device=dev)
    out = torch.empty((n_pairs, 12), dtype=torch.float32, device=dev)
    L = _lib.lib()
    stream = torch.cuda.current_stream(dev)
    call = lambda: _lib.check(L.nans_check_collision_device(n_pairs, d["type"].data_ptr(), d["posrad_a"].data_ptr(),
                                                             d["verts_a"].data_ptr(), d["posrad_b"].data_ptr(),
                                                             d["verts_b"].data_ptr(), hit.data_ptr(), out.data_ptr(),
                                                             stream.cuda_stream))
    for _ in range(3):
        call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(reps):
        call()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    ovf = C.c_int32(0)
    _lib.check(L.nans_check_collision_device_status(C.byref(ovf)))     # an EPA overflow would read as a miss
    res = {"pairs": n_pairs, "ms": ms, "pairs_per_s": n_pairs / (ms * 1e-3), "hit_rate": float(hit.float().mean().item()),
           "strata": "CC:CS:SS = 8:7:1, rotated unit cubes, r in [0.1,0.5], seed 1234, generated on the device",
           "alg_bytes_per_pair": 264, "hbm_frac": 264 * n_pairs / (ms * 1e-3) / 1e9 / load_peaks()[0],
           "bound": "fp32 issue / divergence (GJK+EPA), not HBM", "l2": "inputs larger than L2 (3.5 GB)"}
    m = min(n_check, n_pairs)
    if m > 0:
        t0 = time.perf_counter()
        idx = torch.linspace(0, n_pairs - 1, m, device=dev).long()
        host = lambda k: d[k][idx].cpu().numpy()
        t, pra, prb = host("type"), host("posrad_a"), host("posrad_b")
        va, vb = host("verts_a"), host("verts_b")
        h_gpu = hit[idx].cpu().numpy()
        workers = max(1, min(os.cpu_count() or 1, 32))
        cuts = np.linspace(0, m, 4 * workers + 1).astype(int)
        jobs = [(t[a:b], np.ascontiguousarray(pra[a:b, :3]), va[a:b], np.ascontiguousarray(pra[a:b, 3]),
                 np.ascontiguousarray(prb[a:b, :3]), vb[a:b], np.ascontiguousarray(prb[a:b, 3]))
                for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
        try:
            with ProcessPoolExecutor(workers, mp_context=mp.get_context("spawn")) as ex:
                got = list(ex.map(_ref_check_chunk, jobs))
            kinds = sorted({k for k, _ in got})
            h_ref = np.concatenate([h for _, h in got])
            res.update(checked=int(m), checker="oracle/_ref/nans.so CheckCollision" if kinds == ["reference"] else "oracle port",
                       flags_bit_exact=bool(np.array_equal(h_gpu, h_ref)), mismatches=int((h_gpu != h_ref).sum()),
                       check_seconds=time.perf_counter() - t0, check_workers=workers)
        except Exception as ex:  # the checker is optional here; the -m gpu tests are the gate
            res["flags_bit_exact"] = f"not checked ({type(ex).__name__}: {ex})"
    del d, hit, out
    torch.cuda.empty_cache()
    return res


# 1M cubes per GPU; steps [10, 25): 0.2-0.6 contacts per body (the pile compacting from the floor) and well before the
# blow-up on every slab: the CPU oracle stepping the same 8M-cube world sees every slab calm at step 30 and the first
# one gone by step 35 (WHEN a pile blows apart under the reference's solver is chaotic)
SLAB_SHAPE = dict(side_x=125, ny=16, nz=500, settle=10, window=15)


def slab_record(args, rank, local, size, dist, reduce_max, barrier, warmup):
    """Config C5: ONE world of size x 1M cubes in x-slabs (125 x 16 x 500 cubes per GPU), NCCL halo exchange +
    cross-GPU dataflow solve over NVLink peer memory (csrc/slab.cu), exact reference order.  Weak scaling of one
    world: efficiency = ms/step of one GPU stepping ONE slab alone (same scene, same window, measured here on
    every rank, max taken) / ms/step of N GPUs stepping the N-slab world.  The flat shape is used because how
    long a pile survives the reference's unstable one-pass solver is chaotic: among eight 100-layer
    slabs one blows apart by step ~50, which the exchange (rightly) refuses to run."""
    import torch
    from nans_projekat_b200 import scenes
    from nans_projekat_b200.slab import SlabWorld
    from nans_projekat_b200.world import World
    side_x, ny, nz, settle, window = (SLAB_SHAPE[k] for k in ("side_x", "ny", "nz", "settle", "window"))
    m = side_x * ny * nz
    owned = scenes.cube_pile_slabs(n_slabs=size, side_x=side_x, ny=ny, nz=nz, seed=7, slab=rank, centre=True)
    stream = torch.cuda.Stream()

    def measure(step, restore, sync):
        def run(n):
            for k in range(n):
                if k % window == 0:
                    restore()
                step()
        run(warmup)
        sync(); barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        run(args.steps)
        e1.record(stream)
        torch.cuda.synchronize(); barrier()
        return reduce_max(e0.elapsed_time(e1)) / args.steps

    # baseline: this rank's slab ALONE on its GPU (no neighbours, no exchange)
    alone = World(owned, device=local, stream=stream.cuda_stream)
    alone.rebuild_vertices()
    for _ in range(settle):
        alone.step(DT)
    alone.synchronize(); alone.snapshot()
    ms_alone = measure(lambda: alone.step(DT), alone.restore, alone.synchronize)
    alone.close()
    del alone
    torch.cuda.empty_cache()

    halo_cap = 3 * ny * nz          # three face layers (one is in use while the pile stands)
    sw = SlabWorld(owned, rank, size, dist, local, gid_base=rank * m, halo_cap=halo_cap, capacity=m + halo_cap,
                   stream=stream.cuda_stream)
    sw.rebuild_vertices()
    for _ in range(settle):
        sw.step(DT)
    sw.world.synchronize()
    sw.world.snapshot()
    ms = measure(lambda: sw.step(DT), sw.world.restore, sw.world.synchronize)
    try:                        # the exchange pattern violated (an exploded pile): the world would not be exact
        st, err = sw.status(), None
    except Exception as ex:
        st, err = {"ghosts": -1, "halo_message_bytes": 0}, str(ex)
    ws = sw.world.stats(strict=False)
    info = [None] * size
    dist.all_gather_object(info, {"ghosts": st["ghosts"], "contacts": ws["n_contacts"], "pairs": ws["n_pairs"],
                                  "solver_levels": ws["solver_levels"], "error": err})
    sw.close()
    torch.cuda.empty_cache()
    if any(i["error"] for i in info):
        return {"workload": f"cube_pile_{size}M_one_world", "error": [i["error"] for i in info if i["error"]][0],
                "note": "the timed numbers are withheld: a step whose exchange was refused is not a valid step"}
    return {"workload": f"cube_pile_{size}M_one_world ({size * side_x}x{ny}x{nz} cubes in {size} x-slabs, config C5)",
            "bodies": size * m, "n_gpus": size, "ms_per_step": ms, "body_steps_per_s": size * m / (ms * 1e-3),
            "scaling": "weak (one world grows with the GPU count: 1M cubes per GPU)",
            "one_gpu_one_slab_ms_per_step": ms_alone, "efficiency_vs_one_gpu_one_slab": ms_alone / ms,
            "halo_message_bytes": st["halo_message_bytes"], "per_rank": info, "settle_steps": settle, "window": window,
            "exchange": "per step: ncclAllGather of the slab boxes (32 B/rank), one fixed-capacity ncclSend/ncclRecv of halo "
                        "bodies to the lower / from the upper neighbour, boundary velocities handed to their owner by peer "
                        "stores over NVLink from inside the solve kernel; no host round trip; exact (bit-identical to one "
                        "GPU: tools/slab_check.py, profiles/)"}


# --------------------------------------------------------------------------------------- our arm
FULL = ("pos", "vel", "force", "ang", "angvel", "torque", "verts")


def prepare_world(scene, local, stream, settle):
    """World on the device, vertices rebuilt, `settle` free-running steps, snapshot taken."""
    from nans_projekat_b200.world import World
    world = World(scene, device=local, stream=stream.cuda_stream)
    world.rebuild_vertices()
    for _ in range(settle):        # scene preparation: let the pile come into contact
        world.step(DT)
    world.synchronize()
    world.snapshot()
    return world


def run_steps(world, n, window, step_fn):
    """n steps; every `window` steps the world returns to the snapshot (inside whatever is being timed)."""
    for k in range(n):
        if k % window == 0:
            world.restore()
        step_fn()


def time_steps(world, stream, steps, warmup, window, barrier=lambda: None):
    """ms for `steps` steps, CUDA events on the launching stream, after `warmup` untimed ones."""
    import torch
    run_steps(world, warmup, window, lambda: world.step(DT))
    world.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(); torch.cuda.synchronize()
    t_begin = time.time()
    e0.record(stream)
    run_steps(world, steps, window, lambda: world.step(DT))
    e1.record(stream)
    torch.cuda.synchronize(); barrier()
    return e0.elapsed_time(e1), t_begin, time.time()


def profile_stages(world, n_prof, window):
    """per-stage ms (CUDA events between the stages) and mean pair / contact / DAG-depth counts over n_prof steps"""
    stage, acc = {}, {"pairs": 0.0, "contacts": 0.0, "levels": 0.0}

    def prof_step():
        m = world.step_profiled(DT)
        for k, v in m.items():
            stage[k] = stage.get(k, 0.0) + v / n_prof
        s_ = world.stats()
        acc["pairs"] += s_["n_pairs"] / n_prof
        acc["contacts"] += s_["n_contacts"] / n_prof
        acc["levels"] += s_["solver_levels"] / n_prof
    run_steps(world, n_prof, window, prof_step)
    return stage, acc


def stage_roofline(stage, acc, nb, peak):
    alg = {"integrate_forces": nb * BYTES["integrate_forces"],
           "broadphase": nb * (BYTES["aabb_key"] + BYTES["radix_sort"]) + acc["pairs"] * BYTES["pair_emit"],
           "narrowphase": acc["pairs"] * BYTES["narrowphase"],
           "solver": acc["contacts"] * BYTES["solver"],
           "integrate_velocities": nb * BYTES["integrate_velocities"]}
    per = {k: {"ms": stage[k], "alg_bytes": alg[k], "gbs": alg[k] / (stage[k] * 1e-3) / 1e9,
               "frac": alg[k] / (stage[k] * 1e-3) / 1e9 / peak} for k in alg}
    return alg, per


def parity_in_run(world, scene, prepared):
    """One step of the headline state on the GPU and on the CPU oracle (grid prefilter = the reference's all-pairs
    list, tests/test_oracle_grid.py): contact list and post-step state must be bit-identical.  The oracle is the
    CHECKER here; nothing it computes enters a timed region."""
    from oracle import oracle as O
    t0 = time.perf_counter()
    w = O.World(scene.n_cubes, scene.n_spheres, scene.n_statics)
    for f in scene.ARRAYS:
        getattr(w, f)[...] = getattr(scene, f)
    w.rebuild_vertices()
    for f in FULL:
        getattr(w, f)[...] = getattr(prepared, f)
    world.restore()
    world.step(DT)
    gc = world.contacts()
    d = world.download(fields=FULL)
    t_or = time.perf_counter()
    oc = w.step(DT, prefilter="grid", cap=max(8 * scene.nb, 1 << 20))
    t_or = time.perf_counter() - t_or
    same_c = gc.tobytes() == oc.tobytes()
    bad = [f for f in ("pos", "vel", "ang", "angvel", "verts") if getattr(d, f).tobytes() != getattr(w, f).tobytes()]
    world.restore()
    return {"checked": "one step from the prepared state, GPU vs CPU oracle (oracle/nans_oracle.c, pinned to the "
                       "reference's nans.so), contact list in reference order + pos/vel/ang/angvel/verts of every body",
            "contacts": int(len(oc)), "contact_list_bit_exact": bool(same_c), "state_bit_exact": not bad,
            "fields_differing": bad, "seconds": time.perf_counter() - t0, "oracle_step_seconds": t_or}


def sub_record(scene, name, local, settle, window, steps, warmup, shard_note=None, barrier=lambda: None, reduce_max=None):
    """a secondary workload: device-timed body-steps/s + stage split (no e2e, no CPU leg)"""
    import torch
    stream = torch.cuda.Stream()
    world = prepare_world(scene, local, stream, settle)
    ms, _, _ = time_steps(world, stream, steps, warmup, window, barrier)
    if reduce_max is not None:
        ms = reduce_max(ms)
    stage, acc = profile_stages(world, min(steps, window), window)
    st = world.stats()
    world.close()
    torch.cuda.empty_cache()
    rec = {"workload": name, "bodies": scene.nb, "ms_per_step": ms / steps, "body_steps_per_s": scene.nb * steps / (ms * 1e-3),
           "settle_steps": settle, "window": window, "steps": steps, "pairs_per_step": acc["pairs"],
           "contacts_per_step": acc["contacts"], "contacts_per_body": acc["contacts"] / max(scene.nb, 1),
           "solver_dag_depth": acc["levels"], "stages_ms": stage, "overflow": st["overflow"]}
    if shard_note:
        rec["sharding"] = shard_note
    return rec


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist
    from nans_projekat_b200.world import kernel_launches
    from nans_projekat_b200.scenes import Scene
    from nans_projekat_b200 import scenes

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    affinity = bind_to_gpu_numa(local)
    torch.cuda.set_device(local)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world_size > 1:
            dist.barrier()

    def reduce_max(x):
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        if world_size > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    warmup = max(args.warmup, 3)
    # every rank steps the SAME world (seed 7): how long a pile survives the reference's unstable solver depends on
    # its jitter, and the timed window must lie before the blow-up on every rank
    scene, layers, workload_name = build_workload(args, seed=7)
    nb = scene.nb
    stream = torch.cuda.Stream()
    # The measured window is steps [settle, settle + window) of the simulation, the contact-rich phase
    # of the pile.  The reference's bug-compatible solver (minus sign in the angular JMJ term, one
    # Gauss-Seidel pass) eventually blows a large pile apart, so every `window` steps
    # the world returns to the prepared state by a device-to-device snapshot restore (an episode
    # reset, RL-style).  The restore is INSIDE the timed region; it is ~0.2 GB of D2D copy per window.
    world = prepare_world(scene, local, stream, args.settle)
    prepared = world.download(fields=FULL)
    window = args.window

    # ---- timed region: device-resident, CUDA events on the launching stream ----------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.5)          # let nvidia-smi start sampling before the timed region
    run_steps(world, warmup, window, lambda: world.step(DT))
    world.synchronize()
    launches0 = kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(); torch.cuda.synchronize()
    t_begin = time.time()
    e0.record(stream)
    run_steps(world, args.steps, window, lambda: world.step(DT))
    e1.record(stream)
    torch.cuda.synchronize(); barrier()
    t_end = time.time()
    ms = e0.elapsed_time(e1)
    launches = kernel_launches() - launches0
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    world.stats()          # raises on capacity overflow: a truncated step is not a valid step
    ms_max = reduce_max(ms)
    value = nb * world_size * args.steps / (ms_max * 1e-3)

    # ---- per-stage pass over the same window: CUDA events between the stages ----------------
    stage, acc = profile_stages(world, min(args.steps, 2 * window), window)
    pairs_acc, contacts_acc = acc["pairs"], acc["contacts"]
    peak, peak_src = load_peaks()
    alg, per_stage = stage_roofline(stage, acc, nb, peak)
    dom = max(alg, key=lambda k: stage[k])
    ach = alg[dom] / (stage[dom] * 1e-3) / 1e9
    traffic, traffic_src, pipes = None, None, None
    try:   # DRAM bytes per launch + pipe utilisation of the dominant kernel from the committed `ncu --set full` capture:
        # only when the capture was taken on THIS workload (else null: a capture of another scene says nothing here)
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        if tj.get("workload") == workload_name:
            # the narrowphase stage of a cube-only world is two kernels (GJK, then EPA over the intersecting pairs):
            # traffic = both, pipes = the EPA kernel's (80 % of the stage)
            knames = {"narrowphase": ["narrowphase_world_epa_kernel", "narrowphase_world_gjk_kernel"],
                      "solver": ["solve_versioned_kernel"], "integrate_forces": ["integrate_forces_kernel"],
                      "integrate_velocities": ["integrate_velocities_kernel"], "broadphase": ["pair_count_kernel"]}[dom]
            if dom == "narrowphase" and knames[0] not in tj["dram_bytes_per_launch"]:
                knames = ["narrowphase_world_kernel"]          # a world with spheres: one kernel
            got = [tj["dram_bytes_per_launch"].get(k) for k in knames]
            traffic = sum(got) if all(g is not None for g in got) else None
            traffic_src = tj["source"] + " : " + " + ".join(knames)
            pipes = tj.get("pipes", {}).get(knames[0])
    except Exception:
        pass
    limiter = {"narrowphase": "FP32 issue slots / divergence / local-memory latency of GJK+EPA (not HBM): see pipes_ncu",
               "solver": "dependency depth of the exact-order sweep (depth x store->poll->apply hop), not HBM",
               "broadphase": "L2 latency of the cell-hash probes + launch count", "integrate_forces": "HBM",
               "integrate_velocities": "HBM"}[dom]
    roofline = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "share_of_step": stage[dom] / stage["step"],
                "limited_by": limiter,
                "pipes_ncu": pipes,   # FMA / ALU pipe and issue-slot utilisation, active lanes per instruction (same capture)
                "per_stage": per_stage,
                "note": "`bound` names the accounting (algorithmic bytes / stage time / measured HBM copy peak, as the "
                        "bench contract asks), `limited_by` what actually limits the dominant stage; stage = all kernels "
                        "of that stage (CUDA events between stages on the launching stream); pairs/s = %.3g"
                        % (pairs_acc / (stage["narrowphase"] * 1e-3))}

    # ---- e2e: through the public API with HOST buffers, H2D + D2H inside the timed region ----
    e2e = None
    if not args.no_e2e:
        def pinned(shape):
            return torch.zeros(shape, dtype=torch.float32).pin_memory().numpy()
        io = Scene.__new__(Scene)
        io.n_cubes, io.n_spheres, io.n_statics, io.world_id = scene.n_cubes, scene.n_spheres, scene.n_statics, None
        io.force, io.torque, io.pos, io.ang = pinned((nb, 3)), pinned((nb, 3)), pinned((nb, 3)), pinned((nb, 3))

        def timed(step_fn):
            run_steps(world, warmup, window, step_fn)
            barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            run_steps(world, args.steps, window, step_fn)
            torch.cuda.synchronize()
            return nb * world_size * args.steps / reduce_max(time.perf_counter() - t0)

        def e2e_step():
            world.upload_async(io, fields=("force", "torque"))  # this frame's external forces/torques
            world.step(DT)                                      # detection first: the H2D overlaps it
            world.download_into(io, ("pos", "ang"))             # this frame's poses, synchronous
        def e2e_blocking_step():
            world.upload(io, fields=("force", "torque"))
            world.step(DT)
            world.download_into(io, ("pos", "ang"))
        v_block = timed(e2e_blocking_step)
        e2e = {"value": timed(e2e_step), "unit": "body-steps/s",
               "h2d_bytes_per_step": int(2 * nb * 12), "d2h_bytes_per_step": int(2 * nb * 12),
               "api": "World.upload_async(force,torque) -> World.step -> World.download(pos,ang) "
                      "(nans_world_upload_async / nans_step / nans_world_download): every frame's own poses are in "
                      "host memory when the frame's calls return; the force/torque copy overlaps the detection "
                      "phase, which reads neither; pinned host buffers, wall clock",
               "blocking_upload": {"value": v_block, "unit": "body-steps/s",
                                   "api": "World.upload(force,torque) -> World.step -> World.download(pos,ang), "
                                          "every call blocking"}}
        assert np.isfinite(io.pos).all(), "non-finite positions after the e2e loop"

        # the same loop through the pipelined I/O calls: H2D of frame k+1 and D2H of frame k overlap the
        # step on their own streams; the host consumes frame k's poses during frame k+1 (one frame late)
        outs = []
        for _ in range(2):
            o = Scene.__new__(Scene)
            o.n_cubes, o.n_spheres, o.n_statics, o.world_id = io.n_cubes, io.n_spheres, io.n_statics, None
            o.pos, o.ang = pinned((nb, 3)), pinned((nb, 3))
            outs.append(o)
        state = {"k": 0, "ticket": None, "sum": 0.0}

        def e2e_pipe_step():
            k = state["k"]
            world.upload_async(io, fields=("force", "torque"))
            world.step(DT)
            t = world.download_async(outs[k & 1], fields=("pos", "ang"))
            if state["ticket"] is not None:
                world.wait(state["ticket"])                      # frame k-1's poses are now in host memory
                state["sum"] += float(outs[(k - 1) & 1].pos[0, 1])  # touch the result
            state["ticket"], state["k"] = t, k + 1

        def run_pipe(n):
            for k in range(n):
                if k % window == 0:
                    world.wait(-1); state["ticket"] = None
                    world.restore()
                e2e_pipe_step()
            world.wait(-1)
        run_pipe(warmup)
        barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        run_pipe(args.steps)
        torch.cuda.synchronize()
        elp = reduce_max(time.perf_counter() - t0)
        e2e["pipelined"] = {"value": nb * world_size * args.steps / elp, "unit": "body-steps/s",
                            "api": "World.upload_async(force,torque) -> World.step -> World.download_async(pos,ang) -> "
                                   "World.wait(ticket of the previous frame): same bytes per step, copies on their own "
                                   "streams, poses consumed one frame late"}
        assert np.isfinite(outs[0].pos).all() and np.isfinite(outs[1].pos).all()

    single = rank == 0 and world_size == 1
    subs = not args.no_subrecords
    # ---- throughput solver mode: the same Gauss-Seidel pass in a shuffled sweep order (NOT the reference's results) ----
    solver_modes = None
    if subs:
        def vel_after_one_step():
            world.restore(); world.step(DT)
            d_ = world.download(fields=("vel", "angvel"))
            return d_.vel.astype(np.float64), d_.angvel.astype(np.float64)
        v_ex, w_ex = vel_after_one_step()
        world.set_solver("shuffled")
        v_sh, w_sh = vel_after_one_step()
        ms_sh, _, _ = time_steps(world, stream, args.steps, warmup, window, barrier)
        ms_sh = reduce_max(ms_sh)
        stage_sh, acc_sh = profile_stages(world, min(args.steps, window), window)
        world.set_solver("exact")
        world.restore()
        rel = lambda a, b: np.abs(a - b) / np.maximum(np.abs(b), 1.0)
        dv, dw = rel(v_sh, v_ex), rel(w_sh, w_ex)
        touched = (dv.max(1) > 0) | (dw.max(1) > 0)
        solver_modes = {
            "exact": {"ms_per_step": ms_max / args.steps, "solver_ms": stage["solver"], "dag_depth": acc["levels"],
                      "results": "bit-identical to the reference's sequential sweep (parity tests)"},
            "shuffled": {"ms_per_step": ms_sh / args.steps, "solver_ms": stage_sh["solver"], "dag_depth": acc_sh["levels"],
                         "body_steps_per_s": nb * world_size * args.steps / (ms_sh * 1e-3),
                         "results": "one Gauss-Seidel pass of the same Constraint in a fixed pseudo-random order "
                                    "(bit-identical to the oracle sweeping the list in that order: tests/test_solver_gpu.py); "
                                    "NOT the reference's list order",
                         "deviation_from_exact_after_one_step": {
                             "max_rel_vel": float(dv.max()), "max_rel_angvel": float(dw.max()),
                             "mean_rel_vel": float(dv.mean()), "bodies_differing_frac": float(touched.mean()),
                             "how": "|x_shuffled - x_exact| / max(|x_exact|, 1), same prepared state, one step"}},
            "speedup_step": (ms_max / args.steps) / (ms_sh / args.steps),
            "headline_uses": "exact"}
    parity = None
    if single and subs:
        try:
            parity = parity_in_run(world, scene, prepared)
        except Exception as ex:   # the -m gpu tests are the gate; a missing checker must not lose the bench line
            parity = {"checked": f"not checked ({type(ex).__name__}: {ex})"}
    cpu = None
    if single and not args.no_cpu_baseline and args.workload == "pile":
        state = prepared.copy()
        for f in ("scale", "mass", "moi", "st_pos", "st_ang", "st_scale", "st_mass", "st_moi"):
            getattr(state, f)[...] = getattr(scene, f)
        sv = Scene(0, 0, scene.n_statics)
        world.download_into(sv, ("st_verts",))
        state.st_verts[...] = sv.st_verts
        cpu = cpu_port_baseline(state, args.cpu_bodies)
        if parity and parity.get("oracle_step_seconds"):
            # the SAME workload on the host cores: the oracle port's arithmetic behind a uniform-grid prefilter.  NOT the
            # reference's algorithm (its all-pairs loop cannot run 10^6 bodies) -- a CPU with this repo's candidate
            # search; GJK+EPA on min(nproc, 32) threads, grid / solver / integrators on one (oracle/nans_oracle.c)
            cpu["same_workload_grid_port"] = {
                "value": nb / parity["oracle_step_seconds"], "unit": "body-steps/s", "bodies": nb,
                "cores": min(os.cpu_count() or 1, 32), "seconds_per_step": parity["oracle_step_seconds"],
                "what": "ONE step of the headline state by the oracle port with its uniform-grid prefilter (the step "
                        "parity_in_run checks the GPU against); not the reference's O(N^2) all-pairs loop"}
    world.close()
    del world
    torch.cuda.empty_cache()

    # ---- sub-records: the other configurations of BASELINE.json, same harness, shorter runs ----
    extra = {}
    if subs and args.workload == "pile":
        if single:
            alt_side = 250 if args.side == 100 else 100
            alt_scene, alt_layers = build_pile(args.bodies, alt_side, seed=7)
            extra["alt_shapes"] = [dict(sub_record(alt_scene, f"cube_pile_1M_{alt_side}x{alt_side}x{alt_layers}", local,
                                                   80 if alt_side == 100 else 40, window, args.steps, warmup),
                                        shape=f"{alt_side}x{alt_layers}x{alt_side}")]
            del alt_scene
            # steps [50, 70): the drop's contact-rich phase BEFORE the reference's one-pass solver blows it apart (from
            # step ~60 velocities grow without bound in the oracle and on the GPU alike; by step 150 the
            # cubes are a dispersed cloud 10^6 units wide, which is not a collision workload any more)
            extra["c2_drop10k"] = sub_record(scenes.cube_drop(n=10000, seed=1), "cube_drop_10k (config C2)", local, 50, 20,
                                             max(args.steps, 40), warmup)
        # config C4: a FIXED batch of 4096 worlds x (48 cubes + 16 spheres), 4096 / N worlds per GPU, no communication
        if 4096 % world_size == 0:
            per = 4096 // world_size
            c4 = scenes.batched_worlds(per, 48, 16, seed=1 + rank)   # per-world seeds = seed * 1000003 + world: all distinct
            r = sub_record(c4, f"batched_worlds_{per}x64 per GPU (config C4: 4096 worlds over {world_size} GPU)", local, 30, 20,
                           max(args.steps, 40), warmup, barrier=barrier, reduce_max=reduce_max,
                           shard_note=f"4096/{world_size} worlds per GPU, strong scaling of a fixed batch, no collective")
            r["body_steps_per_s"] = 4096 * 64 * r["steps"] / (r["ms_per_step"] * r["steps"] * 1e-3)
            r["bodies_total"] = 4096 * 64
            extra["c4_worlds4096"] = r
    if world_size > 1 and subs and args.workload == "pile":
        try:
            extra["slab"] = slab_record(args, rank, local, world_size, dist, reduce_max, barrier, warmup)
        except Exception as ex:
            if world_size > 1:
                raise               # a failed collective would hang the other ranks: fail the whole job loudly
    c3_pairs = args.c3 if args.c3 >= 0 else (16777216 if single else 0)
    c3 = run_c3(c3_pairs, local, args.c3_check) if (c3_pairs and rank == 0 and subs) else None
    if rank == 0:
        shape = f"{args.side}x{layers}x{args.side}"
        line = {"metric": "body-steps/s", "value": value, "unit": "body-steps/s", "n_gpus": world_size,
                "steps": args.steps, "warmup": warmup, "ms_per_step": ms_max / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": workload_name, "shape": shape,
                           "bodies_per_world": nb, "worlds": world_size, "footprint": f"{args.side}x{args.side}",
                           "layers": layers, "spacing": 1.02, "jitter": 0.005, "dt": float(DT), "settle_steps": args.settle,
                           "parallelism": "1 world per GPU, no collective" if world_size > 1 else "1 world on 1 GPU",
                           "l2": "inputs larger than L2 (>= 0.6 GB of world state touched per step vs 126 MB L2)",
                           "window": f"steps [{args.settle}, {args.settle + window}) of the simulation, restored from a "
                                     f"device snapshot every {window} steps inside the timed region",
                           "solver": "exact reference order (versioned body rows: value + version in one 128-bit row)",
                           "pairs_per_step": pairs_acc, "contacts_per_step": contacts_acc,
                           "contacts_per_body": contacts_acc / nb, "solver_dag_depth": acc["levels"]},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
                "clocks": clocks,
                "stages_ms": stage, "narrowphase_pairs_per_s": pairs_acc / (stage["narrowphase"] * 1e-3),
                "parity_in_run": parity, "solver_modes": solver_modes,
                "host": {"nproc": os.cpu_count(), "affinity": affinity}}
        line.update(extra)
        if c3:
            line["c3_narrowphase"] = c3
        print(json.dumps(line), flush=True)
    if world_size > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
