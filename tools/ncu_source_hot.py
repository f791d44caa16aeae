#!/usr/bin/env python
"""Hot source lines of one kernel of an .ncu-rep (needs -lineinfo + --import-source on).
usage: python tools/ncu_source_hot.py <report.ncu-rep> <kernel-regex> [top N] [launch id]"""
import csv
import io
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    cmd = ["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + rx]
    if len(sys.argv) > 4:
        cmd += ["--launch-skip", sys.argv[4], "--launch-count", "1"]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No" and len(r) > 4][0]
    hdr = rows[hi]
    si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
    lines = [r for r in rows[hi + 1:] if len(r) > si and r[0] != "" and r[si].isdigit()]
    tot_s = sum(int(r[si]) for r in lines) or 1
    tot_i = sum(int(r[ii]) for r in lines) or 1
    print("kernel %s: %d samples, %.2fM warp instructions" % (rx, tot_s, tot_i / 1e6))
    print("%6s %6s %7s  %s" % ("line", "samp%", "inst%", "source"))
    for r in sorted(lines, key=lambda r: -int(r[si]))[:top]:
        print("%6s %6.1f %7.1f  %s" % (r[0], 100.0 * int(r[si]) / tot_s, 100.0 * int(r[ii]) / tot_i, r[1].strip()[:130]))


if __name__ == "__main__":
    main()
