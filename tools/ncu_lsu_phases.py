#!/usr/bin/env python
"""Shared-memory wavefronts (and their excess over the conflict-free count) and global sectors per phase of the step
kernel, from an ncu source dump.  Runs here, no GPU:

    ncu -i gpurun_out/<tag>_prof.ncu-rep --page source --csv --print-source cuda,sass > dump.csv
    python tools/ncu_lsu_phases.py dump.csv        (uses the first kernel of the dump)
"""
import csv
import re
import sys

SRC = "pve_mcc_for_unsignalized_intersection_b200/csrc/scene_step.cuh"
lines = open(SRC).read().split("\n")
marks = []
for i, l in enumerate(lines, 1):
    m = re.search(r"/\* ---- (.*?) -*\s*\*/", l) or re.search(r"/\* ---- ([^-].*)$", l)
    if m:
        marks.append((i, m.group(1).strip()[:44]))
    elif l.startswith("template <int NT>") or l.startswith("PVE_DEV void pve_world_xy"):
        marks.append((i, "fn: " + lines[i][:40]))


def phase_of(ln):
    lab = "(top)"
    for s, l in marks:
        if ln >= s:
            lab = "%4d %s" % (s, l)
    return lab


rows = list(csv.reader(open(sys.argv[1])))
cur, col, hdr, agg, order, files_seen = None, None, None, {}, [], 0
for r in rows:
    if not r:
        continue
    if r[0] in ("File Path", "File Name"):
        cur = r[1].split("/")[-1]
        files_seen += cur == "pve_mcc.cu"
        if files_seen > 1:
            break                                   # the dump's second kernel
        continue
    if r[0] == "Line No":
        col, hdr = {}, r
        for i, h in enumerate(r):
            col.setdefault(h, i)
        continue
    if col is None or len(r) < len(hdr):
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue

    def f(n):
        try:
            return float(r[col[n]])
        except (ValueError, KeyError):
            return 0.0
    key = phase_of(ln) if cur == "scene_step.cuh" else "other: " + cur
    if key not in agg:
        agg[key] = [0.0] * 7
        order.append(key)
    a = agg[key]
    for q, n in enumerate(("Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal",
                           "L1 Wavefronts Shared Excessive", "L2 Theoretical Sectors Global",
                           "L2 Theoretical Sectors Global Excessive", "# Samples")):
        a[q] += f(n)
tot = [sum(a[q] for a in agg.values()) for q in range(7)]
print("instructions %.0f; shared wavefronts %.0f (conflict-free %.0f, excess %.0f = %.0f%%); global sectors %.0f (excess %.0f)"
      % (tot[0], tot[1], tot[2], tot[3], 100 * tot[3] / max(tot[1], 1), tot[4], tot[5]))
print("%-50s %6s %8s %9s %8s %7s" % ("phase", "inst%", "shared%", "x ideal", "global%", "samp%"))
for k in sorted(order):
    a = agg[k]
    print("%-50s %5.1f%% %7.1f%% %9.2f %7.1f%% %6.1f%%" % (k, 100 * a[0] / tot[0], 100 * a[1] / max(tot[1], 1),
                                                         a[1] / max(a[2], 1), 100 * a[4] / max(tot[4], 1),
                                                         100 * a[6] / max(tot[6], 1)))
