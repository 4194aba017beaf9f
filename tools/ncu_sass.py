#!/usr/bin/env python
"""SASS of the profiled kernel in address order with per-instruction stall samples and executed counts:
`python tools/ncu_sass.py report.ncu-rep [min_inst_pct]` (ncu --page source --print-source cuda,sass --csv)."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur, fname, hdr, seen, seq = None, "", None, set(), []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r; continue
    if r[0].isdigit():
        cur = (fname, int(r[0])); continue
    if len(r) > 4 and r[2].startswith("0x") and hdr:
        d = dict(zip(hdr, r))
        a = int(r[2], 16)
        if a in seen:
            continue  # inlined code is listed once per source line that claims it
        seen.add(a)
        try:
            seq.append((a, r[3].strip(), int(d["# Samples"]), int(d["Instructions Executed"]), cur))
        except (ValueError, KeyError):
            pass
seq.sort()
ts = sum(s[2] for s in seq) or 1
ti = sum(s[3] for s in seq) or 1
print("instructions %d, executed %d, samples %d" % (len(seq), ti, ts))
for a, ins, smp, ie, ln in seq:
    if 100.0 * ie / ti >= thr:
        print("%5.2f%% smp %5.2f%% inst  %s:%-4d %s" % (100.0 * smp / ts, 100.0 * ie / ti, ln[0][:14], ln[1], ins))
