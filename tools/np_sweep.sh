#!/bin/bash
# rebuild the narrowphase with different occupancy targets and report the stage time (run on the GPU box)
cd "$(dirname "$0")/.."
for cfg in "128 4" "128 5" "128 6" "128 8" "64 10" "64 12" "256 2" "256 3"; do
  set -- $cfg
  rm -f nans_projekat_b200/csrc/narrowphase.o
  make -C nans_projekat_b200/csrc EXTRA="-DNANS_NP_THREADS=$1 -DNANS_NP_MINBLOCKS=$2" >/dev/null 2>&1 || { echo "build failed $cfg"; continue; }
  regs=$(grep -A1 "narrowphase_world" nans_projekat_b200/csrc/narrowphase.ptxas.log | grep -o "Used [0-9]* registers" | head -1)
  echo -n "threads=$1 minblocks=$2 ($regs): "
  python bench.py --no-cpu-baseline --no-e2e --steps 40 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('narrowphase %.3f ms  step %.3f ms' % (d['stages_ms']['narrowphase'], d['ms_per_step']))"
done
rm -f nans_projekat_b200/csrc/narrowphase.o; make -C nans_projekat_b200/csrc >/dev/null 2>&1
