#!/bin/bash
# A/B on the GPU box: bench every library in gpurun_variants/ (built offline by tools/build_variants.sh),
# then optionally run the GPU tests and one ncu --set full capture of a kernel with the default library.
# usage: tools/gpu_ab.sh [--tests] [--ncu KERNEL_REGEX NAME] [--workload W]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TESTS=0; NCU=""; NAME=cap; EXTRA=""
while [ $# -gt 0 ]; do
  case "$1" in
    --tests) TESTS=1;;
    --ncu) NCU="$2"; NAME="$3"; shift 2;;
    --extra) EXTRA="$2"; shift;;
  esac
  shift
done
cp nans_projekat_b200/libnans_b200.so /tmp/libnans_default.so
for so in gpurun_variants/*.so; do
  [ -f "$so" ] || continue
  cp "$so" nans_projekat_b200/libnans_b200.so
  n=$(basename "$so" .so)
  timeout 300 python bench.py --no-cpu-baseline --no-e2e --no-subrecords --steps 40 $EXTRA > gpurun_out/ab_$n.json 2> gpurun_out/ab_$n.err
  python - "$n" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/ab_{n}.json").read().strip().splitlines()[-1])
    st = d["stages_ms"]
    print(f"{n}: step {d['ms_per_step']:.3f} ms  " + "  ".join(f"{k} {v:.3f}" for k, v in st.items() if k != "step"), flush=True)
except Exception as e:
    print(n, "FAILED", e, open(f"gpurun_out/ab_{n}.err").read()[-600:], flush=True)
PY
done
cp /tmp/libnans_default.so nans_projekat_b200/libnans_b200.so
if [ $TESTS = 1 ]; then
  timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
fi
if [ -n "$NCU" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$NCU" -c 1 \
      -f -o gpurun_out/$NAME python tools/profile_step.py 1000000 100 85 1 > gpurun_out/$NAME.log 2>&1
  ls -la gpurun_out/$NAME.ncu-rep
fi
