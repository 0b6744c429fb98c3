#!/usr/bin/env python
"""profiles/ncu_traffic.json from an `ncu --set full ... --page raw --csv` dump of one step: per kernel the DRAM
bytes of its largest launch (bench.py's roofline.traffic) and, for the FP32-bound narrowphase, the pipe
utilisation figures north_star asks for.  usage: python tools/ncu_traffic.py raw.csv "source description" workload_name > json"""
import csv, json, sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
f = lambda r, k: float(r[ix[k]].replace(",", "") or 0) if k in ix else None
unit = {h: u for h, u in zip(rows[0], rows[1])}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
dram, pipes = {}, {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0]
    b = sum(f(r, k) * scale.get(unit[k], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    if b > dram.get(name, -1):
        dram[name] = b
        pipes[name] = {
            "fma_pipe_pct": f(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
            "alu_pipe_pct": f(r, "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
            "issue_active_pct": f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "lanes_per_instruction": f(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
            "warps_active_pct": f(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            "l1_hit_pct": f(r, "l1tex__t_sector_hit_rate.pct"),
            "duration_us": f(r, "gpu__time_duration.sum") * {"us": 1.0, "ms": 1e3, "ns": 1e-3}.get(unit["gpu__time_duration.sum"], 1.0),
        }
keep = ("narrowphase_world_kernel", "narrowphase_world_epa_kernel", "narrowphase_world_gjk_kernel", "solve_versioned_kernel",
        "pair_count_kernel")
print(json.dumps({"source": sys.argv[2] if len(sys.argv) > 2 else sys.argv[1],
                  "workload": sys.argv[3] if len(sys.argv) > 3 else None,   # bench.py only uses the capture for this workload
                  "dram_bytes_per_launch": dram,
                  "pipes": {k: pipes[k] for k in keep if k in pipes}}, indent=1))
