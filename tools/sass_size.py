#!/usr/bin/env python
"""SASS size per function and per source line: python tools/sass_size.py [kernel-substring]"""
import collections, os, re, subprocess, sys, tempfile
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "slam-gmapping-learn_b200", "libgm_b200.so")
pat = sys.argv[1] if len(sys.argv) > 1 else "k_scanmatchILb1"
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=td, capture_output=True)
    cub = [f for f in os.listdir(td) if f.startswith("gm_kernels.")][0]
    sass = subprocess.run(["nvdisasm", "-g", os.path.join(td, cub)], capture_output=True, text=True).stdout
cur = None; fn = None; cnt = collections.Counter(); sizes = collections.Counter()
for line in sass.splitlines():
    m = re.match(r'\s*\.text\.(\S+):', line) or re.match(r'^(\S+):\s*$', line)
    if m and not m.group(1).startswith('.L'): fn = m.group(1)
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', line) and fn and not fn.startswith('.'):
        sizes[fn] += 1
        if pat in fn and '$' not in fn: cnt[cur] += 1
for k, v in sizes.most_common(10): print("%7.1f KB %s" % (v * 16 / 1024, k))
for k, v in cnt.most_common(16): print(v, k)
