#!/bin/bash
# round 2: full GPU test-suite + the default bench line
cd "$GRAFT_REPO_ROOT"
timeout 420 python -m pytest tests -x -q -m gpu > gpurun_out/r2_probe2_pytest.txt 2>&1
tail -5 gpurun_out/r2_probe2_pytest.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_probe2_bench.json 2> gpurun_out/r2_probe2_bench.err
tail -c 1500 gpurun_out/r2_probe2_bench.json; tail -5 gpurun_out/r2_probe2_bench.err
