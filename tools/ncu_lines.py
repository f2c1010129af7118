#!/usr/bin/env python
"""Per-CUDA-source-line view of an `ncu --page source --csv --print-source cuda,sass` dump.
Usage: ncu_lines.py dump.csv [min_pct]   (lines in file order with >= min_pct of instructions or samples)"""
import csv
import sys

path = sys.argv[1]
min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
rows = list(csv.reader(open(path)))
cur_file, hdr, col = None, None, None
data = []
for r in rows:
    if not r:
        continue
    if r[0] in ("File Path", "File Name"):
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        col = {h: i for i, h in enumerate(hdr)}
        continue
    if hdr is None or "Instructions Executed" not in col or len(r) < len(hdr):
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    def f(name):
        try:
            return int(float(r[col[name]]))
        except ValueError:
            return 0
    data.append((cur_file, ln, f("Instructions Executed"), f("# Samples"), f("Thread Instructions Executed"),
                 f("stall_barrier"), f("stall_long_sb"), f("stall_short_sb"), f("stall_wait"), r[1]))
tot = sum(d[2] for d in data)
ts = sum(d[3] for d in data)
print("total warp-instructions %d, samples %d" % (tot, ts))
print("%-16s %5s %9s %6s %6s %6s %5s %5s %5s %5s %5s  %s" % (
    "file", "line", "inst", "inst%", "samp", "samp%", "thr", "bar", "lsb", "ssb", "wait", "source"))
for fl, ln, inst, samp, thr, sb, sl, ss, sw, src in data:
    if inst >= min_pct / 100 * tot or samp >= min_pct / 100 * ts:
        print("%-16s %5d %9d %5.1f%% %6d %5.1f%% %5.1f %5d %5d %5d %5d  %s" % (
            fl[:16], ln, inst, 100.0 * inst / max(tot, 1), samp, 100.0 * samp / max(ts, 1), thr / max(inst, 1),
            sb, sl, ss, sw, src.strip()[:90]))
