#!/bin/bash
# round 2, first probe: headline-size parity tests + 100^3 shape timing with the round-1 kernels
cd "$GRAFT_REPO_ROOT"
nproc > gpurun_out/r2_probe1_nproc.txt
timeout 900 python -m pytest tests/test_headline_gpu.py -x -q -m gpu > gpurun_out/r2_probe1_pytest.txt 2>&1
tail -5 gpurun_out/r2_probe1_pytest.txt
for settle in 40 80 100; do
timeout 600 python bench.py --steps 20 --warmup 3 --side 100 --settle $settle --no-cpu-baseline > gpurun_out/r2_probe1_bench100_s$settle.json 2> gpurun_out/r2_probe1_bench100_s$settle.err
done
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_probe1_bench250.json 2> gpurun_out/r2_probe1_bench250.err
tail -c 600 gpurun_out/r2_probe1_bench100_s80.json
