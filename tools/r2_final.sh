#!/bin/bash
# end-of-round evidence on ONE GPU: smoke(), GPU tests, the default bench line, the reference arm
cd "$GRAFT_REPO_ROOT"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.txt 2>&1; tail -2 gpurun_out/r2_final_smoke.txt
timeout 420 python -m pytest tests -q -m gpu > gpurun_out/r2_final_pytest_gpu.txt 2>&1; tail -2 gpurun_out/r2_final_pytest_gpu.txt
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final_bench_reference.json 2> gpurun_out/r2_final_bench.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final_bench.json 2>> gpurun_out/r2_final_bench.err
tail -c 400 gpurun_out/r2_final_bench.json; tail -3 gpurun_out/r2_final_bench.err
