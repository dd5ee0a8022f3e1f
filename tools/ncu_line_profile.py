#!/usr/bin/env python
"""Attributes the executed warp instructions of one kernel (ncu --set full --import-source on report) to CUDA source
lines, by aligning the SASS page of the report with `nvdisasm -g` of the cubin the kernel came from.
usage: ncu_line_profile.py report.ncu-rep kernel_regex lib.so mangled_substring [nth_launch]"""
import csv, os, re, subprocess, sys, tempfile
rep, kre, lib, mangled = sys.argv[1:5]
which = int(sys.argv[5]) if len(sys.argv) > 5 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
lo = starts[which]; hi = starts[which + 1] if which + 1 < len(starts) else len(rows)
H = rows[lo + 1]; data = [r for r in rows[lo + 2:hi] if len(r) == len(H)]
ie, so, ti = H.index("Instructions Executed"), H.index("Source"), H.index("Thread Instructions Executed")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
sass = []  # (opcode, file, line)
for f in sorted(os.listdir(tmp)):
    if "-" in f.split(".")[0]:
        continue
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout.splitlines()
    sec = [i for i, l in enumerate(dis) if l.startswith(".text.") and mangled in l]
    if not sec:
        continue
    cur = ("?", 0); inl = []
    for l in dis[sec[0] + 1:]:
        if l.startswith("//--------------------- "):
            break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)), m.group(3)); continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", l)
        if m:
            sass.append((m.group(1).strip(), cur))
    break
print("ncu sass lines", len(data), "nvdisasm instructions", len(sass))
n = min(len(data), len(sass))
agg = {}
tot = 0; tthr = 0
for k in range(n):
    c = int(data[k][ie] or 0); t = int(data[k][ti] or 0)
    key = sass[k][1][:2]
    a = agg.setdefault(key, [0, 0, 0]); a[0] += c; a[1] += t; a[2] += 1
    tot += c; tthr += t
print("total warp instr %d, avg active lanes %.1f" % (tot, tthr / max(tot, 1)))
for key in sorted(agg):
    a = agg[key]
    if a[0] >= 0.002 * tot:
        print("%-18s:%4d  sass=%3d  warp-instr=%10d  %5.1f%%  lanes=%4.1f" % (key[0], key[1], a[2], a[0], 100.0 * a[0] / tot, a[1] / max(a[0], 1)))
