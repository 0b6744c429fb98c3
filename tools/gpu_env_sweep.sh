#!/bin/bash
# bench.py (device-resident legs only) once per environment setting: tools/gpu_env_sweep.sh "A=1" "A=2 B=3" ...
cd "$(dirname "$0")/.."
for cfg in "$@"; do
  echo -n "$cfg: "
  env $cfg timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-subrecords --steps 60 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('step %.4f ms  solver %.3f  np %.3f  bp %.3f' % (d['ms_per_step'], d['stages_ms']['solver'], d['stages_ms']['narrowphase'], d['stages_ms']['broadphase']))"
done
