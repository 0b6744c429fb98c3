#!/bin/bash
# Narrowphase residency sweep.  `build` (here, no GPU needed) compiles one library per configuration
# into gpurun_variants/; `run` (on the GPU box) benches each of them.
# cfg = THREADS MINBLOCKS [extra -D flags]
cd "$(dirname "$0")/.."
CFGS=(
  "128 5"
  "128 4"
  "128 6"
  "128 5 -DNANS_NP_BOX_EPA=1"
  "128 5 -DNANS_NP_V4=1"
  "128 5 -DNANS_NP_TREE_SUPPORT=1"
)
V=gpurun_variants
case "$1" in
build)
  mkdir -p $V; rm -f $V/*.so
  for cfg in "${CFGS[@]}"; do
    set -- $cfg
    t=$1; b=$2; shift 2; extra="$*"
    name="t${t}_b${b}$(echo "$extra" | sed 's/-DNANS_NP_/_/g; s/[ =]//g')"
    rm -f nans_projekat_b200/csrc/narrowphase.o
    make -C nans_projekat_b200/csrc EXTRA="-DNANS_NP_THREADS=$t -DNANS_NP_MINBLOCKS=$b $extra" >/dev/null 2>&1 || { echo "build failed $cfg"; continue; }
    regs=$(grep -A2 "narrowphase_world" nans_projekat_b200/csrc/narrowphase.ptxas.log | grep -o "Used [0-9]* registers" | head -1)
    echo "$name: $regs"
    cp nans_projekat_b200/libnans_b200.so $V/$name.so
  done
  rm -f nans_projekat_b200/csrc/narrowphase.o; make -C nans_projekat_b200/csrc >/dev/null 2>&1
  ;;
run)
  cp nans_projekat_b200/libnans_b200.so /tmp/libnans_default.so
  for so in $V/*.so; do
    cp $so nans_projekat_b200/libnans_b200.so
    echo -n "$(basename $so .so): "
    timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 40 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('narrowphase %.3f ms  step %.3f ms' % (d['stages_ms']['narrowphase'], d['ms_per_step']))"
  done
  cp /tmp/libnans_default.so nans_projekat_b200/libnans_b200.so
  ;;
esac
