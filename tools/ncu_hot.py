#!/usr/bin/env python
"""Hot SASS lines of one kernel from an `ncu --set full --import-source on` report:
   python tools/ncu_hot.py report.ncu-rep kernel_regex [top]"""
import csv
import io
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
if not hi:
    sys.exit("no source page for " + rx)
hdr = rows[hi[0]]
end = hi[1] - 1 if len(hi) > 1 else len(rows)   # first matching launch only
data = [r for r in rows[hi[0] + 1:end] if len(r) == len(hdr)]
si, src, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
tot = sum(int(r[si]) for r in data)
print("kernel:", rows[hi[0] - 1][1][:100] if hi[0] else "?", "| samples", tot, "| SASS lines", len(data),
      "| warp instr", sum(int(r[ie]) for r in data))
top = sorted(((int(r[si]), i) for i, r in enumerate(data)), reverse=True)[:top_n]
for s, i in sorted(top, key=lambda x: x[1]):
    print("%5d %6d %5.1f%% %9s  %s" % (i, s, 100.0 * s / max(tot, 1), data[i][ie], data[i][src][:100]))
