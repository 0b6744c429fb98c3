#!/bin/bash
# Build one libnans_b200.so per "name:flags" argument into gpurun_variants/ (offline; nvcc cross-compiles),
# then restore the default build.  Files whose objects must be rebuilt: all (flags may touch any header).
# usage: tools/build_variants.sh "base:" "blk6:-DNANS_NP_MINBLOCKS=6" ...
cd "$(dirname "$0")/.."
V=gpurun_variants; mkdir -p $V; rm -f $V/*.so
C=nans_projekat_b200/csrc
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  rm -f $C/*.o
  make -C $C EXTRA="$flags" >/dev/null 2>&1 || { echo "build failed: $spec"; continue; }
  regs=$(grep -A3 "Compiling entry function '_ZN4nans24narrowphase_world" $C/narrowphase.ptxas.log | grep -o "Used [0-9]* registers.*smem" | head -1)
  echo "$name [$flags]: $regs"
  cp nans_projekat_b200/libnans_b200.so $V/$name.so
done
rm -f $C/*.o; make -C $C >/dev/null 2>&1
