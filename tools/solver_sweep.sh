#!/bin/bash
# sweep the dataflow solver's knobs; prints solver-stage ms for each setting (1M-cube pile, steps 40-60)
for cfg in "1 0 32 3" "1 1 32 3" "2 0 32 3" "1 0 0 3" "1 0 100 3" "1 0 32 2" "1 0 32 1" "4 0 32 3" "1 1 0 2"; do
  set -- $cfg
  echo -n "hops=$1 atomics=$2 sleep=$3 blocks=$4 : "
  NANS_FLOW_HOPS=$1 NANS_FLOW_ATOMICS=$2 NANS_FLOW_SLEEP=$3 NANS_FLOW_BLOCKS=$4 python bench.py --no-cpu-baseline --no-e2e --steps 40 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('solver %.3f ms  step %.3f ms' % (d['stages_ms']['solver'], d['ms_per_step']))"
done
