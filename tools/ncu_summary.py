"""Summarise an `ncu --page raw --csv` dump: one block per profiled kernel with the metrics the
roofline discussion uses.  usage: python tools/ncu_summary.py raw.csv [more-metric-substrings...]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
WANT = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed.sum', 'smsp__inst_executed.avg.per_cycle_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.sum',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct']
extra = sys.argv[2:]
for r in rows[2:]:
    print('=====', r[idx['Kernel Name']][:90])
    for w in WANT:
        if w in idx:
            print(f"  {w:72s} {r[idx[w]]:>18s} {units[idx[w]]}")
    stalls = [(h, r[i]) for h, i in idx.items() if 'issue_stalled' in h and 'pcsamp' in h and 'not_issued' not in h]
    tot = sum(float(v.replace(',', '') or 0) for _, v in stalls) or 1
    top = sorted(stalls, key=lambda kv: -float(kv[1].replace(',', '') or 0))[:7]
    print('  stall samples: ' + ', '.join(f"{h.split('issue_stalled_')[1]}={float(v.replace(',', ''))/tot*100:.0f}%" for h, v in top))
    for e in extra:
        for h, i in idx.items():
            if e in h:
                print(f"  {h:72s} {r[i]:>18s} {units[i]}")
