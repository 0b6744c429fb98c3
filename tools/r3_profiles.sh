#!/bin/bash
# Evidence run on ONE GPU: compute-sanitizer on every kernel of the step, the ncu launch list of two
# steps of the headline workload (100x100x100 pile, steps 85-87) and one `ncu --set full` capture of a step.
# Everything lands in gpurun_out/<tag>_*; what is to be judged is copied into profiles/.
cd "$(dirname "$0")/.."
TAG=${1:-r3}
O=gpurun_out; mkdir -p $O
for tool in memcheck initcheck synccheck racecheck; do
  echo "== compute-sanitizer --tool $tool" >> $O/${TAG}_sanitizer.txt
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_small.py 2>&1 | grep -v "^$" | tail -6 >> $O/${TAG}_sanitizer.txt
done
tail -30 $O/${TAG}_sanitizer.txt
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/${TAG}_launches.csv python tools/profile_step.py 1000000 100 85 2 > $O/${TAG}_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -c 45 -f -o $O/${TAG}_full \
    python tools/profile_step.py 1000000 100 85 1 > $O/${TAG}_full.log 2>&1
ncu -i $O/${TAG}_full.ncu-rep --page raw --csv > $O/${TAG}_full_raw.csv 2>/dev/null
python tools/ncu_summary.py $O/${TAG}_full_raw.csv > $O/${TAG}_ncu_full_summary.txt 2>&1
python tools/ncu_traffic.py $O/${TAG}_full_raw.csv "profiles/${TAG}_ncu_full_summary.txt (ncu --set full, 1M-cube pile 100x100x100, step 86; per kernel: the largest launch of the step)" cube_pile_1M_100x100x100 > $O/${TAG}_ncu_traffic.json
rm -f $O/${TAG}_full.ncu-rep
ls -la $O | tail -12
