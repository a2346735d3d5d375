#!/usr/bin/env python
"""Annotated SASS listing: ncu source page joined with nvdisasm -g line info.  Prints per instruction: index, source line,
warp instructions per `unit` (e.g. per chain step), avg threads, stall samples %, SASS text.
Usage: ncu_annot.py <src.csv> <cubin> <kernel-substring> <unit-count>"""
import csv, re, subprocess, sys
src_csv, cubin, kname, unit = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
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
tot_s = sum(int(float(r[col["# Samples"]] or 0)) for r in body)
for i, (r, ln) in enumerate(zip(body, lines)):
    ie = int(float(r[col["Instructions Executed"]] or 0)); ss = int(float(r[col["# Samples"]] or 0))
    at = float(r[col["Avg. Threads Executed"]] or 0)
    print("%4d %-22s %7.2f thr %4.1f smp %5.2f%% lsb %5.2f%%  %s" % (i, "%s:%d" % (ln[0][:14], ln[1]) if ln else "?", ie / unit, at,
          100.0 * ss / tot_s, 100.0 * int(float(r[col["stall_long_sb"]] or 0)) / tot_s, r[col["Source"]]))
