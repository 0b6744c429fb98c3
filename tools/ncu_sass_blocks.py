#!/usr/bin/env python
"""Execution profile of a kernel's SASS in blocks of N instructions: executions per instruction,
average active lanes, stall samples.  Shows where a warp runs diverged.
usage: ncu -i rep --page source --csv --print-source sass --kernel-name regex:K > dump.csv
       python tools/ncu_sass_blocks.py dump.csv [block=60]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
blk = int(sys.argv[2]) if len(sys.argv) > 2 else 60
for k, r in enumerate(rows):
    if r and r[0] == "Address":
        hdr, start = r, k + 1
        break
iI, iT, iS = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
data = [(r[1].strip(), int(r[iI]), int(r[iT]), int(r[iS])) for r in rows[start:] if len(r) > iT and r[iI].isdigit()]
print(len(data), "sass instructions")
for b in range(0, len(data), blk):
    seg = data[b:b + blk]
    I, T, S = sum(x[1] for x in seg), sum(x[2] for x in seg), sum(x[3] for x in seg)
    if I == 0:
        continue
    ops = {}
    for x in seg:
        t = x[0].split()
        op = t[1] if t[0].startswith("@") else t[0]
        ops[op] = ops.get(op, 0) + 1
    top = sorted(ops.items(), key=lambda kv: -kv[1])[:4]
    print(f"{b:5d} execs/instr {I / len(seg) / 1e3:8.1f}k lanes {T / max(I, 1):5.1f} samples {S:6d}  {top}")
