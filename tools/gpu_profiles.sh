#!/bin/bash
# End-of-session evidence run on the GPU box: GPU tests, bench lines of every configuration, the ncu launch
# list of two steps and one `ncu --set full` capture of every kernel of a step (+ the two C3 kernels).
# Everything lands in gpurun_out/<tag>_*; copy what is to be judged into profiles/.
cd "$(dirname "$0")/.."
TAG=${1:-r1_s2}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee $O/${TAG}_pytest_gpu.txt
timeout 600 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -c 300 $O/${TAG}_bench.err
timeout 300 python bench.py --impl reference > $O/${TAG}_bench_reference.json 2>> $O/${TAG}_bench.err
timeout 600 python bench.py --c3 16777216 --no-cpu-baseline --no-e2e --steps 40 > $O/${TAG}_bench_with_c3.json 2>> $O/${TAG}_bench.err
timeout 300 python bench.py --workload drop10k --no-cpu-baseline > $O/${TAG}_bench_c2_drop10k.json 2>> $O/${TAG}_bench.err
timeout 300 python bench.py --workload worlds4096 --no-cpu-baseline > $O/${TAG}_bench_c4_worlds4096.json 2>> $O/${TAG}_bench.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/${TAG}_launches.csv python tools/profile_step.py 1000000 250 45 2 > $O/${TAG}_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -c 40 -f -o $O/${TAG}_full \
    python tools/profile_step.py 1000000 250 45 1 > $O/${TAG}_full.log 2>&1
ncu -i $O/${TAG}_full.ncu-rep --page raw --csv > $O/${TAG}_full_raw.csv 2>/dev/null
python tools/ncu_summary.py $O/${TAG}_full_raw.csv > $O/${TAG}_ncu_full_summary.txt 2>&1
cat > /tmp/c3_small.py <<'PY'
import sys; sys.path.insert(0, ".")
import torch, bench
torch.cuda.cudart().cudaProfilerStart()
print(bench.run_c3(1 << 20, 0, reps=1))
torch.cuda.cudart().cudaProfilerStop()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gjk_split|epa_refill" -c 2 -f -o $O/${TAG}_c3_full \
    python /tmp/c3_small.py > $O/${TAG}_c3_full.log 2>&1
ncu -i $O/${TAG}_c3_full.ncu-rep --page raw --csv > $O/${TAG}_c3_full_raw.csv 2>/dev/null
python tools/ncu_summary.py $O/${TAG}_c3_full_raw.csv > $O/${TAG}_c3_ncu_full_summary.txt 2>&1
rm -f $O/${TAG}_c3_full.ncu-rep
ls -la $O | tail -25
