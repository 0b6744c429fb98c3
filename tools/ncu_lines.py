#!/usr/bin/env python
"""Per-CUDA-source-line summary of an `ncu --page source --csv --print-source cuda,sass` dump.

usage: ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:K > dump.csv
       python tools/ncu_lines.py dump.csv [top_n]
Prints, per source line: stall samples, share, warp instructions, average active lanes, long-scoreboard samples."""
import csv
import io
import sys


def main():
    txt = open(sys.argv[1]).read()
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rows = []
    for sec in txt.split('"File Path",')[1:]:
        lines = sec.split("\n")
        fname = lines[0].strip().strip('"').split("/")[-1]
        rdr = csv.reader(io.StringIO("\n".join(lines[2:])))
        hdr = next(rdr)
        iS, iI, iT = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
        iL = hdr.index("stall_long_sb")
        for r in rdr:
            if len(r) <= iL or not r[0].strip().isdigit():
                continue            # SASS rows have an empty line number
            try:
                rows.append((int(r[iS]), int(r[iI]), int(r[iT]), int(r[iL] or 0), fname, r[0], r[1]))
            except ValueError:
                pass
    tot = sum(x[0] for x in rows)
    toti = sum(x[1] for x in rows)
    tott = sum(x[2] for x in rows)
    print(f"samples {tot}  warp-instr {toti}  avg lanes {tott / max(toti, 1):.1f}")
    for s, i, t, l, f, ln, src in sorted(rows, reverse=True)[:top]:
        print(f"{s:7d} {100 * s / max(tot, 1):5.1f}%  instr {100 * i / max(toti, 1):5.1f}%  lanes {t / max(i, 1):5.1f}  "
              f"longsb {l:6d}  {f}:{ln}: {src.strip()[:100]}")


if __name__ == "__main__":
    main()
