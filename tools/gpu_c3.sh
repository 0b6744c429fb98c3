#!/bin/bash
# config C3 (GJK+EPA pairs/s on random pairs) for every library in gpurun_variants/
cd "$(dirname "$0")/.."
N=${1:-16777216}
cp nans_projekat_b200/libnans_b200.so /tmp/libnans_default.so
for so in gpurun_variants/*.so; do
  cp "$so" nans_projekat_b200/libnans_b200.so
  echo -n "$(basename $so .so): "
  timeout 600 python - $N <<'PY'
import sys, json
sys.path.insert(0, ".")
import bench
r = bench.run_c3(int(sys.argv[1]), 0)
print(json.dumps({k: r[k] for k in ("ms", "pairs_per_s", "hit_rate", "flags_bit_exact_vs_oracle")}))
PY
done
cp /tmp/libnans_default.so nans_projekat_b200/libnans_b200.so
