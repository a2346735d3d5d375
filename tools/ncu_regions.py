#!/usr/bin/env python
"""Instruction / stall-sample share per source line range.  Usage: ncu_regions.py <src.csv> <cubin> <kernel> <file> a-b:name ..."""
import csv, re, subprocess, sys
from collections import defaultdict
src_csv, cubin, kname, fname = sys.argv[1:5]
regs = []
for spec in sys.argv[5:]:
    rng, name = spec.split(":"); a, b = rng.split("-"); regs.append((int(a), int(b), name))
dis = subprocess.run(["nvdisasm", "-g", cubin], stdout=subprocess.PIPE).stdout.decode(errors="replace").splitlines()
lines = []; cur = None; inside = False
for l in dis:
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m: inside = kname in m.group(1); continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", l): lines.append(cur)
rows = list(csv.reader(open(src_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
reg = defaultdict(lambda: [0, 0]); tot = [0, 0]
for r, ln in zip(body, lines):
    ie = int(float(r[col["Instructions Executed"]] or 0)); ss = int(float(r[col["# Samples"]] or 0))
    name = ln[0]
    if ln[0] == fname:
        name = "other " + fname
        for a, b, nm in regs:
            if a <= ln[1] <= b: name = nm; break
    reg[name][0] += ie; reg[name][1] += ss; tot[0] += ie; tot[1] += ss
print("total warp instructions %d" % tot[0])
for k, v in sorted(reg.items(), key=lambda kv: -kv[1][0]):
    print("%-34s inst %5.1f%%  samples %5.1f%%" % (k, 100 * v[0] / tot[0], 100 * v[1] / tot[1]))
