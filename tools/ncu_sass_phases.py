#!/usr/bin/env python
"""Executed warp-instructions of a kernel per barrier-delimited SASS segment, from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > dump.csv`.

Inlined library code (sincos, log, tanh, atomics) has no useful line of its own in the per-line view; in address
order it sits inside the phase that calls it, so cutting the SASS at the BAR.SYNC instructions gives the true
per-phase instruction counts.  Usage: ncu_sass_phases.py dump.csv [launch_index]"""
import csv
import re
import sys
from collections import Counter

path = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = list(csv.reader(open(path)))
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
# launches: a new launch starts when the first file section (pve_mcc.cu, kernel entry) reappears
first_file = rows[hdr_idx[0] - 2][1]
starts = [h for h in hdr_idx if rows[h - 2][1] == first_file]
lo = starts[which]
hi = starts[which + 1] if which + 1 < len(starts) else len(rows)
col = {h: i for i, h in enumerate(rows[hdr_idx[0]])}
ci, ct = col["Instructions Executed"], col["Thread Instructions Executed"]
ce, cw, cwi = col["L1 Wavefronts Shared Excessive"], col["L1 Wavefronts Shared"], col["L1 Wavefronts Shared Ideal"]
sass = {}
cur_line = None
cur_file = None
for i in range(lo, hi):
    r = rows[i]
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] in ("Function Name", "Line No") or len(r) <= ct:
        continue
    if r[0]:
        cur_line = (cur_file, int(r[0]))
        continue
    try:
        addr = int(r[2], 16)
    except ValueError:
        continue
    def f(c):
        try:
            return int(float(r[c]))
        except ValueError:
            return 0
    sass[addr] = (r[3].strip(), f(ci), f(ct), cur_line, f(cw), f(ce))
addrs = sorted(sass)
tot = sum(sass[a][1] for a in addrs)
print("static SASS instructions %d, executed warp-instructions %d" % (len(addrs), tot))
seg = []
cur = {"n": 0, "inst": 0, "thr": 0, "ops": Counter(), "lines": Counter(), "first": None, "wf": 0, "ex": 0}
for a in addrs:
    txt, inst, thr, line, wf, ex = sass[a]
    op = re.sub(r"^@!?U?P\d+\s+", "", txt).split()[0].split(".")[0] if txt else "?"
    cur["n"] += 1; cur["inst"] += inst; cur["thr"] += thr; cur["ops"][op] += inst; cur["wf"] += wf; cur["ex"] += ex
    if line and line[0] and line[0].startswith("scene_step"):
        cur["lines"][line[1]] += inst
    if cur["first"] is None:
        cur["first"] = line
    if op in ("BAR", "EXIT"):
        cur["end"] = txt
        seg.append(cur)
        cur = {"n": 0, "inst": 0, "thr": 0, "ops": Counter(), "lines": Counter(), "first": None, "wf": 0, "ex": 0}
if cur["n"]:
    cur["end"] = "(end)"
    seg.append(cur)
print("%3s %6s %9s %6s %5s %7s %6s  %-22s %s" % ("seg", "static", "inst", "inst%", "thr", "smem_wf", "excess", "lines(min-max by inst)", "top opcodes"))
for k, s in enumerate(seg):
    if s["inst"] == 0:
        continue
    ls = [l for l, c in s["lines"].most_common(6)]
    rng = "%d-%d" % (min(ls), max(ls)) if ls else "-"
    ops = " ".join("%s:%.0f%%" % (o, 100.0 * c / s["inst"]) for o, c in s["ops"].most_common(6))
    print("%3d %6d %9d %5.1f%% %5.1f %7d %6d  %-22s %s   | %s" % (k, s["n"], s["inst"], 100.0 * s["inst"] / tot,
          s["thr"] / max(s["inst"], 1), s["wf"], s["ex"], rng, ops, s["end"][:28]))
