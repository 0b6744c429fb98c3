"""Multi-GPU equivalence check (run under torchrun, one rank per GPU):
the slab-decomposed world must be BIT-IDENTICAL, step by step, to the same world on one GPU.

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/slab_check.py [side] [layers] [steps] [settle]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from nans_projekat_b200 import scenes
from nans_projekat_b200.slab import SlabWorld, CudaEngine
from nans_projekat_b200.world import World

side = int(sys.argv[1]) if len(sys.argv) > 1 else 48
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 8
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 30
settle = int(sys.argv[4]) if len(sys.argv) > 4 else 30
rank, size, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
scene = scenes.cube_pile(n_side=side, layers=layers, seed=7)
scene.pos[:, 1] -= 0.0
eng = CudaEngine(scene, rank, size, local)
eng.rebuild_vertices()
sw = SlabWorld(eng, rank, size, dist)
ref = None
if rank == 0:
    ref = World(scene, device=local)
    ref.rebuild_vertices()
dt = np.float32(1 / 60.)
bad = 0
for k in range(settle + steps):
    sw.step(dt)
    if ref is not None:
        ref.step(dt)
    if k < settle and k % 10:
        continue
    owned = eng.download_owned()
    cg = eng.contacts_global()
    parts = [None] * size
    dist.all_gather_object(parts, (eng.lo, eng.hi, {f: getattr(owned, f) for f in ("pos", "vel", "ang", "angvel", "verts")}, cg,
                                   sw.n_ghosts, sw.halo_bytes))
    if rank == 0:
        full = ref.download()
        rc = ref.contacts()
        ok = True
        for lo, hi, st, _, _, _ in parts:
            for f, a in st.items():
                if not np.array_equal(a.view(np.uint32), getattr(full, f)[lo:hi].view(np.uint32)):
                    ok = False
                    print(f"step {k}: rank range [{lo},{hi}) field {f} differs from the single-GPU world", flush=True)
        allc = np.concatenate([p[3] for p in parts])
        order = np.lexsort((allc["b"], allc["a"], allc["type"] != 0))   # CC first, then CF; by (a, b)
        allc = allc[order]
        key = lambda c: np.stack([c["type"], c["a"], c["b"]], 1)
        if len(allc) != len(rc) or not np.array_equal(key(allc), key(rc)) or allc.tobytes() != rc.tobytes():
            ok = False
            print(f"step {k}: contact lists differ ({len(allc)} vs {len(rc)})", flush=True)
        bad += (not ok)
        if k % 10 == 0 or not ok:
            print(f"step {k:3d} ok={ok} contacts={len(rc)} ghosts/rank={[p[4] for p in parts]} "
                  f"halo_bytes/rank={[p[5] for p in parts]} levels={ref.stats()['solver_levels']}", flush=True)
flag = torch.tensor([bad], device="cuda")
dist.broadcast(flag, 0)
if rank == 0:
    print("SLAB CHECK", "PASSED" if bad == 0 else f"FAILED ({bad} steps)", f"ranks={size} bodies={scene.n_cubes}", flush=True)
dist.destroy_process_group()
sys.exit(1 if int(flag.item()) else 0)
