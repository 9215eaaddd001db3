#!/usr/bin/env python
"""Warp instructions executed per CUDA source line of one kernel of an ncu report (-lineinfo build, --import-source on).

    [NCU_KERNEL=regex] python tools/ncu_lines.py <report.ncu-rep> [top N]"""
import csv, os, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
flt = " -k regex:%s -c 1" % os.environ["NCU_KERNEL"] if os.environ.get("NCU_KERNEL") else ""
out = subprocess.run("ncu -i %s%s --page source --csv --print-source cuda,sass" % (rep, flt), shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
h = rows[hdr]; ie = h.index("Instructions Executed"); isamp = h.index("# Samples")
lines = []; cur_file = ""
for r in rows[:hdr]:
    if r and r[0] == "File Path": cur_file = os.path.basename(r[1])
tot = 0
for r in rows[hdr + 1:]:
    if r and r[0] == "File Path": cur_file = os.path.basename(r[1]); continue
    if len(r) <= ie: continue
    if r[0].isdigit():
        try: e = int(r[ie]); sm = int(r[isamp])
        except ValueError: continue
        lines.append((e, sm, cur_file, int(r[0]), r[1].strip())); tot += e
ts = max(1, sum(l[1] for l in lines))
print("total warp instructions attributed to source lines: %d" % tot)
for e, sm, f, n, src in sorted(lines, reverse=True)[:top]:
    print("%6.2f%% instr %6.2f%% samples  %s:%d  %s" % (100.0 * e / tot, 100.0 * sm / ts, f, n, src[:110]))
