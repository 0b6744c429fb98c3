#!/bin/bash
# end-of-round confirmation on ONE GPU (bounded: the round's GPU budget is nearly spent): the GPU tests, the default
# bench line, smoke(), and the ncu launch list of two steps of the headline workload (100^3 pile, steps 85-87)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 120 python -m pytest tests -q -m gpu > gpurun_out/r4_pytest_gpu.txt 2>&1; tail -2 gpurun_out/r4_pytest_gpu.txt
timeout 100 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r4_bench.json 2> gpurun_out/r4_bench.err
tail -c 300 gpurun_out/r4_bench.json; tail -3 gpurun_out/r4_bench.err
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4_smoke.txt 2>&1; tail -2 gpurun_out/r4_smoke.txt
timeout 45 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r4_launches.csv python tools/profile_step.py 1000000 100 85 2 > gpurun_out/r4_launches.log 2>&1
wc -l gpurun_out/r4_launches.csv
