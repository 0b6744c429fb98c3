#!/bin/bash
# round 2: after the solver rewrite (b128 rows, shuffled mode) and slab v2 (NCCL + peer stores): tests, slab check, benches
cd "$GRAFT_REPO_ROOT"
nvidia-smi -L > gpurun_out/r2_probe3_gpus.txt
timeout 420 python -m pytest tests -x -q -m gpu > gpurun_out/r2_probe3_pytest.txt 2>&1
tail -15 gpurun_out/r2_probe3_pytest.txt
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" -ge 2 ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/slab_check.py 24 16 24 20 10 5 > gpurun_out/r2_probe3_slabcheck.txt 2>&1
tail -12 gpurun_out/r2_probe3_slabcheck.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_probe3_bench2.json 2> gpurun_out/r2_probe3_bench2.err
tail -c 1800 gpurun_out/r2_probe3_bench2.json; tail -5 gpurun_out/r2_probe3_bench2.err
fi
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_probe3_bench1.json 2> gpurun_out/r2_probe3_bench1.err
tail -c 600 gpurun_out/r2_probe3_bench1.json; tail -5 gpurun_out/r2_probe3_bench1.err
