#!/usr/bin/env python
"""Static SASS instructions of one kernel per source line (top N), from a cubin: which source constructs cost code size.
Usage: sass_lines.py lib.so [kernel-substring] [top]   (needs cuobjdump + nvdisasm)"""
import collections
import os
import re
import subprocess
import sys
import tempfile

lib = os.path.abspath(sys.argv[1])
pat = sys.argv[2] if len(sys.argv) > 2 else "pve_step_kernelILi128ELi128ELi96ELb0"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cub = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-g", "-c", cub], stdout=subprocess.PIPE, text=True).stdout.split("\n")
inside, cur, cnt, total = False, None, collections.Counter(), 0
for ln in txt:
    if ln.startswith("\t.section\t.text."):
        inside = pat in ln
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        cnt[cur] += 1
        total += 1
print("total static instructions", total)
src = {}
for (f, l), c in cnt.most_common(top):
    if f not in src:
        for root in ("pve_mcc_for_unsignalized_intersection_b200/csrc", "include"):
            pth = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), root, f)
            if os.path.exists(pth):
                src[f] = open(pth).read().split("\n")
    text = src.get(f, [""] * (l + 1))[l - 1].strip()[:100] if f in src and l - 1 < len(src[f]) else ""
    print("%5d  %-22s %5d  %s" % (c, f[:22], l, text))
