#!/usr/bin/env python
"""Per-source-line instruction / stall-sample totals of an .ncu-rep
(needs --import-source on and -lineinfo).
usage: ncu_lines.py report.ncu-rep [topN]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source",
                      "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
fname, hdr = None, None
agg = collections.OrderedDict()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        iE = hdr.index("Instructions Executed")
        iN = hdr.index("# Samples")
        continue
    if r[0] == "Function Name" or hdr is None:
        continue
    if r[0] != "":  # a source line with its aggregated metrics
        try:
            e, n = int(r[iE]), int(r[iN])
        except ValueError:
            continue
        agg[(fname, int(r[0]))] = (e, n, r[1].strip())
tot = sum(v[0] for v in agg.values())
tots = sum(v[1] for v in agg.values())
print("total warp instructions", tot, "samples", tots)
for (f, l), (e, n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-22s %4d  inst %9d %5.1f%%  samp %5.1f%%  %s" % (f, l, e, 100.0 * e / tot,
                                                           100.0 * n / max(1, tots), s[:90]))
