#!/bin/bash
# A/B on the GPU box: "name|library|ENV=.. ENV=.." per argument; bench the headline (short form) with each, then
# optionally the GPU tests under the LAST configuration.  usage: tools/r3_ab.sh [--tests] spec...
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TESTS=0
if [ "$1" = "--tests" ]; then TESTS=1; shift; fi
cp nans_projekat_b200/libnans_b200.so /tmp/libnans_default.so
for spec in "$@"; do
  IFS='|' read -r name lib envs <<< "$spec"
  [ -n "$lib" ] && cp "$lib" nans_projekat_b200/libnans_b200.so || cp /tmp/libnans_default.so nans_projekat_b200/libnans_b200.so
  env $envs timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-subrecords --steps 40 $BENCH_EXTRA > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/ab_{n}.json").read().strip().splitlines()[-1])
    st = d["stages_ms"]
    print(f"{n}: step {d['ms_per_step']:.4f} ms  " + "  ".join(f"{k} {v:.4f}" for k, v in st.items() if k != "step"), flush=True)
except Exception as e:
    print(n, "FAILED", e, open(f"gpurun_out/ab_{n}.err").read()[-600:], flush=True)
PY
done
if [ $TESTS = 1 ]; then
  env $envs timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
fi
cp /tmp/libnans_default.so nans_projekat_b200/libnans_b200.so
