#!/usr/bin/env python
"""Per-kernel count, mean duration and share of the step from an ncu launch list
(`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`, e.g. profiles/r4_launches.csv).
Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's stage split, not absolutes.

usage: python tools/launch_shares.py profiles/r4_launches.csv [steps_in_the_capture]"""
import collections
import csv
import sys


def main():
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
        a = agg.setdefault(r[ki].split("(")[0].replace("void ", ""), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"{sum(a[0] for a in agg.values())} launches, {tot / steps:.1f} us per step (serialised, cold cache), {steps} steps")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k:44s} n/step {a[0] / steps:4.1f}  mean {a[1] / a[0]:7.1f} us  share {a[1] / tot:.3f}")


if __name__ == "__main__":
    main()
