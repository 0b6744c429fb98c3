#!/bin/bash
# bench.py at N = the GPUs of this box (incl. the slab sub-record) + the slab exactness check
cd "$GRAFT_REPO_ROOT"
NG=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541 tools/slab_check.py 16 16 24 20 10 5 > gpurun_out/r3_probeN_slabcheck_$NG.txt 2>&1
grep "SLAB CHECK" gpurun_out/r3_probeN_slabcheck_$NG.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $NG --steps 20 --warmup 5 > gpurun_out/r3_probeN_bench_$NG.json 2> gpurun_out/r3_probeN_bench_$NG.err
grep "^{" gpurun_out/r3_probeN_bench_$NG.json | tail -c 1500; tail -3 gpurun_out/r3_probeN_bench_$NG.err
