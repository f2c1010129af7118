#!/usr/bin/env python
"""Aggregate the per-line ncu source dump into the kernel's phases (comments `/* ---- X: ...`)."""
import csv, re, sys
path = sys.argv[1]
src = sys.argv[2] if len(sys.argv) > 2 else "pve_mcc_for_unsignalized_intersection_b200/csrc/scene_step.cuh"
lines = open(src).read().split("\n")
marks = []  # (line_no, label)
for i, l in enumerate(lines, 1):
    m = re.search(r"/\* ---- (.*?) -*\s*\*/", l)
    if m: marks.append((i, m.group(1).strip()[:60]))
    elif l.startswith("template <int NT>") or l.startswith("PVE_DEV void pve_world_xy") or l.startswith("PVE_HD size_t pve_smem_carve"):
        marks.append((i, "fn: " + lines[i][:50] if i < len(lines) else "fn"))
def phase_of(ln):
    lab = "(top)"
    for s, l in marks:
        if ln >= s: lab = "%4d %s" % (s, l)
    return lab
rows = list(csv.reader(open(path)))
cur=None; col=None; agg={}; order=[]
for r in rows:
    if not r: continue
    if r[0] in ("File Path","File Name"): cur=r[1].split("/")[-1]; continue
    if r[0]=="Line No": col={h:i for i,h in enumerate(r)}; hdr=r; continue
    if col is None or "Instructions Executed" not in col or len(r)<len(hdr): continue
    try: ln=int(r[0])
    except ValueError: continue
    def f(n):
        try: return int(float(r[col[n]]))
        except ValueError: return 0
    key = phase_of(ln) if cur=="scene_step.cuh" else "other: "+cur
    if key not in agg: agg[key]=[0,0,0,0,0,0,0]; order.append(key)
    a=agg[key]; a[0]+=f("Instructions Executed"); a[1]+=f("# Samples"); a[2]+=f("Thread Instructions Executed")
    a[3]+=f("stall_barrier"); a[4]+=f("stall_long_sb"); a[5]+=f("stall_short_sb"); a[6]+=f("stall_wait")
ti=sum(a[0] for a in agg.values()); ts=sum(a[1] for a in agg.values())
print("total inst %d samples %d"%(ti,ts))
print("%-66s %9s %6s %6s %6s %5s %5s %5s %5s %5s"%("phase","inst","inst%","samp","samp%","thr","bar","lsb","ssb","wait"))
for k in sorted(order):
    a=agg[k]
    print("%-66s %9d %5.1f%% %6d %5.1f%% %5.1f %5d %5d %5d %5d"%(k,a[0],100*a[0]/ti,a[1],100*a[1]/ts,a[2]/max(a[0],1),a[3],a[4],a[5],a[6]))
