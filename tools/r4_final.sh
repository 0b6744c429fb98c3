#!/bin/bash
# end-of-round confirmation on ONE GPU (bounded: the round's GPU budget is nearly spent): the default bench line,
# the GPU tests, smoke()
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 150 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r4_bench.json 2> gpurun_out/r4_bench.err
tail -c 300 gpurun_out/r4_bench.json; tail -3 gpurun_out/r4_bench.err
timeout 120 python -m pytest tests -q -m gpu > gpurun_out/r4_pytest_gpu.txt 2>&1; tail -2 gpurun_out/r4_pytest_gpu.txt
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4_smoke.txt 2>&1; tail -2 gpurun_out/r4_smoke.txt
