#!/usr/bin/env python
"""Per-source-line hot spots of an ncu report: `python tools/ncu_lines.py report.ncu-rep [top]`.
Uses `ncu --page source --print-source cuda,sass --csv` (needs -lineinfo at compile time)."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
kernel = ["-k", "regex:" + sys.argv[3]] if len(sys.argv) > 3 else []  # optional third argument: kernel name regex
out = subprocess.run(["ncu", "-i", rep] + kernel + ["--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname = ""
hdr = None
lines = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        hdr = r; si = hdr.index("# Samples"); ii = hdr.index("Instructions Executed"); ti = hdr.index("Thread Instructions Executed")
        continue
    if hdr and r[0].isdigit():
        try:
            lines.append((int(r[si]), int(r[ii]), int(r[ti]), fname, int(r[0]), r[1].strip()))
        except ValueError:
            pass
ts = sum(l[0] for l in lines) or 1
tin = sum(l[1] for l in lines) or 1
print("total samples %d, warp instructions %d, thread instructions %d" % (ts, tin, sum(l[2] for l in lines)))
for s, i, t, f, n, src in sorted(lines, reverse=True)[:top]:
    print("%5.1f%% smp %5.1f%% inst  %-16s:%4d  %s" % (100.0 * s / ts, 100.0 * i / tin, f, n, src[:100]))
