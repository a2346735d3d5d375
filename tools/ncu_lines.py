#!/usr/bin/env python
"""Join an `ncu --page source --csv` export (per SASS instruction) with `nvdisasm -g` line info of the same cubin and
print the per-source-line totals: instructions executed and stall samples.  Usage:
  ncu_lines.py <src.csv> <cubin> <kernel-substring> [top]"""
import csv
import re
import subprocess
import sys
from collections import defaultdict


def main():
    src_csv, cubin, kname = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    dis = subprocess.run(["nvdisasm", "-g", cubin], stdout=subprocess.PIPE).stdout.decode(errors="replace").splitlines()
    lines = []  # (file:line) per instruction, in order, for the kernel
    cur, inside = None, False
    for l in dis:
        m = re.match(r"\s*\.text\.(\S+):", l)
        if m:
            inside = kname in m.group(1)
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", l):
            lines.append(cur)
    rows = list(csv.reader(open(src_csv)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    col = {h: i for i, h in enumerate(hdr)}
    if len(body) != len(lines):
        print("warning: %d SASS rows in ncu vs %d in nvdisasm" % (len(body), len(lines)), file=sys.stderr)
    agg = defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    for r, ln in zip(body, lines):
        ie = int(float(r[col["Instructions Executed"]] or 0))
        ss = int(float(r[col["# Samples"]] or 0))
        lsb = int(float(r[col["stall_long_sb"]] or 0))
        a = agg[ln]
        a[0] += ie; a[1] += ss; a[2] += lsb
        tot[0] += ie; tot[1] += ss; tot[2] += lsb
    print("total inst %d samples %d long_sb %d" % tuple(tot))
    for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%-14s:%-5d inst %5.1f%%  samples %5.1f%%  long_sb %5.1f%%" % (ln[0], ln[1], 100.0 * a[0] / tot[0], 100.0 * a[1] / tot[1], 100.0 * a[2] / max(1, tot[2])))


if __name__ == "__main__":
    main()
