"""Multi-GPU equivalence check (run under torchrun, one rank per GPU): ONE world in x-slabs over N GPUs
(csrc/slab.cu: NCCL halo exchange + cross-GPU dataflow solve over NVLink peer memory) must be BIT-IDENTICAL,
step by step, to the same world on one GPU -- contact list (global ids, reference order) and body state.

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/slab_check.py [side_x] [ny] [nz] [steps] [settle] [every]
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from nans_projekat_b200 import scenes
from nans_projekat_b200.slab import SlabWorld
from nans_projekat_b200.world import World

arg = lambda i, d: int(sys.argv[i]) if len(sys.argv) > i else d
side_x, ny, nz, steps, settle, every = arg(1, 16), arg(2, 12), arg(3, 16), arg(4, 30), arg(5, 30), arg(6, 10)
# spacing below the cube size: lateral neighbours overlap from the first step on, so the slab faces are in contact
# (with the bench's 1.02 spacing cross-slab contacts only appear once the pile starts to shear)
spacing = float(sys.argv[7]) if len(sys.argv) > 7 else 0.998
rank, size, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
m = side_x * ny * nz
owned = scenes.cube_pile_slabs(n_slabs=size, side_x=side_x, ny=ny, nz=nz, seed=7, slab=rank, spacing=spacing, jitter=0.004)
halo_cap = max(1024, 4 * ny * nz)
# the squeezed pile reaches > 8 contacts per body: explicit capacities (the same on every rank), checked below
caps = dict(max_pairs=64 * (m + halo_cap), max_contacts=24 * (m + halo_cap))
sw = SlabWorld(owned, rank, size, dist, local, gid_base=rank * m, halo_cap=halo_cap, capacity=m + halo_cap, **caps)
sw.rebuild_vertices()
ref = None
if rank == 0:
    ref = World(scenes.cube_pile_slabs(n_slabs=size, side_x=side_x, ny=ny, nz=nz, seed=7, spacing=spacing, jitter=0.004), device=local,
                max_pairs=size * caps["max_pairs"], max_contacts=size * caps["max_contacts"])
    ref.rebuild_vertices()
dt = np.float32(1 / 60.)
bad, max_ghosts, max_cross, checked = 0, 0, 0, 0
t0 = time.time()
for k in range(settle + steps):
    sw.step(dt)
    if ref is not None:
        ref.step(dt)
    if k < settle and k % every and not os.environ.get("SLAB_DIAG"):
        continue
    if bad >= 2 and os.environ.get("SLAB_DIAG"):
        break
    try:
        st, err = sw.status(), None
    except Exception as ex:        # the exchange pattern violated: the squeezed pile has blown apart (chaotic step)
        st, err = {"ghosts": -1}, str(ex)
    st["stats"] = sw.world.stats(strict=False)
    o = sw.download_owned()
    cg = sw.contacts_global()
    parts = [None] * size
    dist.all_gather_object(parts, (rank * m, (rank + 1) * m, {f: getattr(o, f) for f in ("pos", "vel", "ang", "angvel", "verts")},
                                   cg, st["ghosts"], st["stats"], err))
    if any(p[6] for p in parts):
        if rank == 0:
            print(f"step {k}: the pile has blown apart and the exchange refuses to go on ({[p[6] for p in parts if p[6]][0][:90]}...): "
                  f"end of the check after {checked} exact steps", flush=True)
            if checked < 4:
                bad += 1
        break
    if rank == 0:
        full = ref.download()
        assert ref.stats()["overflow"] == 0 and all(p[5]["overflow"] == 0 for p in parts), "capacity overflow: enlarge the check's capacities"
        rc = ref.contacts()
        ok = True
        for lo, hi, fields, _, _, _, _ in parts:
            for f, a in fields.items():
                if not np.array_equal(a.view(np.uint32), getattr(full, f)[lo:hi].view(np.uint32)):
                    ok = False
                    rows = np.nonzero((a.view(np.uint32) != getattr(full, f)[lo:hi].view(np.uint32)).reshape(hi - lo, -1).any(1))[0]
                    print(f"step {k}: rank range [{lo},{hi}) field {f} differs from the single-GPU world on {len(rows)} rows, "
                          f"first {rows[:6] + lo}", flush=True)
        allc = np.concatenate([p[3] for p in parts])
        allc = allc[np.lexsort((allc["b"], allc["a"], allc["type"] != 0))]   # CC first, then CF; by (a, b)
        if allc.tobytes() != rc.tobytes():
            ok = False
            n_ = min(len(allc), len(rc))
            key = lambda c: np.stack([c["type"][:n_], c["a"][:n_], c["b"][:n_]], 1)
            d_ = np.nonzero((key(allc) != key(rc)).any(1))[0]
            print(f"step {k}: contact lists differ ({len(allc)} vs {len(rc)}); first differing key at {d_[:1]}: "
                  f"{allc[d_[0]] if len(d_) else None} vs {rc[d_[0]] if len(d_) else None}", flush=True)
        cross = int(((allc["type"] == 0) & (allc["a"] // m != allc["b"] // m)).sum())
        bad += (not ok)
        checked += int(ok)
        max_cross = max(max_cross, cross)
        max_ghosts = max(max_ghosts, max(p[4] for p in parts))
        if k % every == 0 or not ok:
            print(f"step {k:3d} ok={ok} contacts={len(rc)} cross-slab contacts={cross} ghosts/rank={[p[4] for p in parts]} "
                  f"levels={ref.stats()['solver_levels']} overflow/rank={[p[5]['overflow'] for p in parts]} "
                  f"contacts/rank={[p[5]['n_contacts'] for p in parts]}", flush=True)
if rank == 0:
    if max_cross == 0:
        bad += 1
        print("no contact ever joined two slabs: the cross-GPU solve was not exercised", flush=True)
    print("SLAB CHECK", "PASSED" if bad == 0 else f"FAILED ({bad} steps)",
          f"ranks={size} bodies={size * m} ({size} slabs of {side_x}x{ny}x{nz}) max ghosts/rank={max_ghosts} "
          f"max cross-slab contacts={max_cross} halo message={sw.status()['halo_message_bytes']} B  {time.time() - t0:.1f} s", flush=True)
flag = torch.tensor([bad], device="cuda")
dist.broadcast(flag, 0)
sw.close()
dist.destroy_process_group()
sys.exit(1 if int(flag.item()) else 0)
