#!/bin/bash
# A/B on the GPU box: the whole default bench line (no CPU-baseline leg) with every library in gpurun_variants/
# (built offline by tools/build_variants.sh); prints the numbers that decide.  The in-tree library is restored.
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
cp nans_projekat_b200/libnans_b200.so /tmp/libnans_default.so
for so in gpurun_variants/*.so; do
  [ -f "$so" ] || continue
  cp "$so" nans_projekat_b200/libnans_b200.so
  n=$(basename "$so" .so)
  timeout 100 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r4_ab_$n.json 2> gpurun_out/r4_ab_$n.err
  python - "$n" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r4_ab_{n}.json").read().strip().splitlines()[-1])
    st = d["stages_ms"]; p = d["parity_in_run"]
    print(f"{n}: step {d['ms_per_step']:.4f}  np {st['narrowphase']:.4f}  bp {st['broadphase']:.4f}  sol {st['solver']:.4f}  "
          f"alt {d['alt_shapes'][0]['ms_per_step']:.4f} (np {d['alt_shapes'][0]['stages_ms']['narrowphase']:.4f})  "
          f"c2 {d['c2_drop10k']['ms_per_step']:.4f}  c4 {d['c4_worlds4096']['ms_per_step']:.4f} (np {d['c4_worlds4096']['stages_ms']['narrowphase']:.4f})  "
          f"c3 {d['c3_narrowphase']['pairs_per_s']:.4g} exact={d['c3_narrowphase']['flags_bit_exact']}  "
          f"parity contacts={p.get('contact_list_bit_exact')} state={p.get('state_bit_exact')}", flush=True)
except Exception as e:
    print(n, "FAILED", e, open(f"gpurun_out/r4_ab_{n}.err").read()[-600:], flush=True)
PY
done
cp /tmp/libnans_default.so nans_projekat_b200/libnans_b200.so
