#!/usr/bin/env python
"""SASS opcode census of the in-tree library: which kernels use which instruction families (profiles/r02_sass_opcodes.md).
Usage: python tools/sass_census.py [lib.so] > profiles/<name>.md"""
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "pve_mcc_for_unsignalized_intersection_b200/csrc/libpve_mcc.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True).stdout.split("\n")
FAM = [("HMMA", r"(?<!UTC)HMMA"), ("UTC*MMA", r"UTC[A-Z]*MMA"), ("LDTM/STTM", r"LDTM|STTM"), ("UTMALDG/UTMASTG/UBLKCP", r"UTMALDG|UTMASTG|UBLKCP"),
       ("LDGSTS", r"LDGSTS"), ("DFMA/DADD/DMUL", r" D(FMA|ADD|MUL)"), ("BAR", r" BAR\."), ("ATOMS", r"ATOMS"), ("REDG/ATOMG", r"REDG|ATOMG"),
       ("SHFL", r"SHFL"), ("MUFU", r"MUFU")]
rows, cur = [], None
for ln in sass:
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
        cur = {"name": name, "n": 0, **{f: 0 for f, _ in FAM}}
        rows.append(cur)
        continue
    if cur is not None and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s", ln):
        cur["n"] += 1
        for f, pat in FAM:
            if re.search(pat, ln):
                cur[f] += 1
show = re.compile(r"<96, 96, 64, false>|big_kernel<128, 128, 96, false>|pve_step_kernel<128, 128, 96, false>|<512, 384, 320, false>|actor|critic|pvn_|order|total|classify|reset|stats|gsum|recount")
print("# SASS opcode census of the in-tree library (`cuobjdump -sass csrc/libpve_mcc.so`, tools/sass_census.py), round 2\n")
print("What proves (or disproves) Blackwell-native tensor / TMA paths: `UTC*MMA`, `LDTM`/`STTM` (tcgen05 + TMEM), `UTMALDG`/`UTMASTG`/`UBLKCP` "
      "(TMA); `HMMA` is the legacy mma.sync path.\n")
print("| kernel | instructions | " + " | ".join(f for f, _ in FAM) + " |")
print("|---|---|" + "---|" * len(FAM))
for r in sorted(rows, key=lambda r: -r["n"]):
    if show.search(r["name"]):
        print("| `%s` | %d | %s |" % (r["name"].split("(")[0][:64], r["n"], " | ".join(str(r[f]) for f, _ in FAM)))
tot = {f: sum(r[f] for r in rows) for f, _ in FAM}
print("\n%d kernels in the library (the step kernel's other instantiations -- capacity class x CTA size x nbr_src -- have the same opcode mix "
      "and are not listed).  Library totals: %s.\n" % (len(rows), ", ".join("%s %d" % (f, tot[f]) for f, _ in FAM)))
print("Reading: the environment step (`pve_step_kernel`, `pve_step_big_kernel`) is scalar float64 + integer work with shared-memory staging -- no "
      "tensor instructions by design (north_star: no dense contraction on this path; the per-row `cp.async.bulk` stores tried in round 1 were "
      "slower).  The policy / critic kernels of rows N1 / N2 exist three times: `pve_actor_tc_kernel` / `pve_critic_tc_kernel` (csrc/mlp_tc5.cuh, the "
      "default) issue `UTCHMMA` (tcgen05.mma, accumulators in tensor memory read back with `LDTM`) and fetch their weights with one `UBLKCP` bulk "
      "copy; `pve_actor_mma_kernel` / `pve_critic_mma_kernel` are the round-1 `HMMA.16816` (mma.sync) versions kept for A/B "
      "(PVE_ACTOR_IMPL / PVE_CRITIC_IMPL = mma), `pve_actor_kernel` / `pve_critic_kernel` the fp32 FFMA ones (= ffma).")
