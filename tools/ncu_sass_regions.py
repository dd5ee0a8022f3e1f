#!/usr/bin/env python
"""Groups the SASS page of an ncu report into regions of similar execution count (loop bodies) and prints
each region's share of executed warp instructions.  usage: ncu_sass_regions.py report.ncu-rep kernel_regex"""
import csv, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
H = rows[1]; data = rows[2:]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0   # n-th launch matching the regex
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
lo = starts[which]; hi = starts[which + 1] if which + 1 < len(starts) else len(rows)
H = rows[lo + 1]; data = rows[lo + 2:hi]
ie, so, ss = H.index("Instructions Executed"), H.index("Source"), H.index("# Samples")
data = [r for r in data if len(r) == len(H)]
tot = sum(int(r[ie] or 0) for r in data)
print("total warp instructions", tot, "sass lines", len(data))
reg = []
for k, r in enumerate(data):
    c = int(r[ie] or 0)
    if reg and abs(c - reg[-1][2]) <= 0.15 * max(c, reg[-1][2], 1):
        reg[-1][1] = k; reg[-1][3] += c; reg[-1][4] += int(r[ss] or 0)
    else:
        reg.append([k, k, c, c, int(r[ss] or 0)])
for a, b, c, s, sm in reg:
    if s > 0.01 * tot:
        print("sass %4d-%4d n=%3d per-inst=%9d share=%5.1f%% samples=%6d  %s | %s" % (a, b, b - a + 1, c, 100 * s / tot, sm, data[a][so][:38], data[b][so][:38]))
