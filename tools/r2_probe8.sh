#!/bin/bash
# round 2: 8-GPU run: exactness of one world in 8 x-slabs, then the default bench line at N=8 (incl. the slab sub-record)
cd "$GRAFT_REPO_ROOT"
nvidia-smi -L > gpurun_out/r2_probe8_gpus.txt
nvidia-smi topo -m > gpurun_out/r2_probe8_topo.txt 2>&1
NG=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29531 tools/slab_check.py 12 16 24 20 10 5 > gpurun_out/r2_probe8_slabcheck.txt 2>&1
grep -v "^\*\*\*\|OMP_NUM" gpurun_out/r2_probe8_slabcheck.txt | cut -c1-300 | tail -14
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $NG --steps 20 --warmup 5 > gpurun_out/r2_probe8_bench.json 2> gpurun_out/r2_probe8_bench.err
grep "^{" gpurun_out/r2_probe8_bench.json | tail -c 2500; tail -5 gpurun_out/r2_probe8_bench.err
