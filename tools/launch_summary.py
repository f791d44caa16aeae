#!/usr/bin/env python
"""Summarise an ncu --csv launch list (gpu__time_duration.sum [+ dram bytes]) per kernel.
usage: python tools/launch_summary.py <launches.csv> [iterations]"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    iters = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, mi, ui, vi, ii = (hdr.index(n) for n in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value", "ID"))
    per = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        if r[mi] == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        elif r[ui] in ("Kbyte", "Mbyte", "Gbyte"):
            v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[r[ui]]
        per.setdefault(r[ii], {"k": r[ki]})[r[mi]] = v
    agg = collections.OrderedDict()
    for v in per.values():
        e = agg.setdefault(v["k"], [0, 0.0, 0.0])
        e[0] += 1
        e[1] += v.get("gpu__time_duration.sum", 0.0)
        e[2] += v.get("dram__bytes_read.sum", 0.0) + v.get("dram__bytes_write.sum", 0.0)
    tot = sum(e[1] for e in agg.values())
    print("%-78s %8s %9s %9s %7s %9s" % ("kernel", "n/iter", "avg us", "us/iter", "share", "MB/launch"))
    for k, e in sorted(agg.items(), key=lambda x: -x[1][1]):
        name = k.replace("hg::<unnamed>::", "").replace("void ", "")[:78]
        print("%-78s %8.1f %9.1f %9.1f %7.3f %9.1f" % (name, e[0] / iters, e[1] / e[0], e[1] / iters, e[1] / tot, e[2] / e[0] / 1e6))
    print("total %.1f us per iteration" % (tot / iters))


if __name__ == "__main__":
    main()
