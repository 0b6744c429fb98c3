#!/usr/bin/env python
"""Region view of a kernel's SASS from an `ncu --page source --csv --print-source cuda,sass` dump:
consecutive instructions with the same execution count are one region (= one basic-block run of a loop nest).
Prints per region: instructions, executions per instruction, share of all warp instructions, active lanes,
stall samples and the CUDA source lines the instructions come from.
usage: python tools/ncu_regions.py dump.csv [min_share_pct]"""
import csv, io, sys
from collections import Counter

txt = open(sys.argv[1]).read()
minshare = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
ins = {}
for sec in txt.split('"File Path",')[1:]:
    lines = sec.split("\n")
    fname = lines[0].strip().strip('"').split("/")[-1]
    rdr = csv.reader(io.StringIO("\n".join(lines[2:])))
    hdr = next(rdr)
    iS, iI, iT = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    cur = None
    for r in rdr:
        if len(r) <= iT:
            continue
        if r[0].strip().isdigit():
            cur = f"{fname.split('.')[0][:6]}:{r[0]}"
        elif r[2].startswith("0x") and r[iI].isdigit():
            ins[int(r[2], 16)] = (r[3].strip(), int(r[iI]), int(r[iT]), int(r[iS] or 0), cur)
addrs = sorted(ins)
tot = sum(ins[a][1] for a in addrs)
base = addrs[0]
regs = []
for a in addrs:
    op, I, T, S, ln = ins[a]
    if regs and regs[-1]["I"] == I and a - regs[-1]["last"] <= 0x40:
        g = regs[-1]
    else:
        g = dict(start=a, I=I, n=0, T=0, S=0, lines=Counter(), ops=Counter())
        regs.append(g)
    g["n"] += 1; g["T"] += T; g["S"] += S; g["last"] = a
    g["lines"][ln] += 1; g["ops"][op.split()[1] if op.startswith("@") else op.split()[0]] += 1
print(f"{len(addrs)} instructions, {tot/1e6:.1f} M warp-instructions")
for g in regs:
    share = 100.0 * g["I"] * g["n"] / tot
    if share < minshare:
        continue
    lanes = g["T"] / max(g["I"] * g["n"], 1)
    ls = " ".join(f"{k}x{v}" for k, v in g["lines"].most_common(6))
    ops = " ".join(f"{k}x{v}" for k, v in g["ops"].most_common(4))
    print(f"+{(g['start']-base)//16:5d} n={g['n']:4d} execs={g['I']/1e3:7.1f}k share={share:5.2f}% lanes={lanes:4.1f} samp={g['S']:5d} | {ls} | {ops}")
