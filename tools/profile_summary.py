#!/usr/bin/env python
"""Turn an `ncu --set full` report of the step kernel into the summary committed under profiles/.

    python tools/profile_summary.py gpurun_out/<tag>_prof.ncu-rep profiles/<name>   (runs here, no GPU)
"""
import csv
import json
import os
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, r = rows[0], rows[1], rows[2]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
summary = {}
for k in KEYS:
    if k in hdr:
        i = hdr.index(k)
        summary[k] = {"value": r[i], "unit": units[i]}
def num(k):
    return float(summary[k]["value"].replace(",", ""))
def to_bytes(k):
    u = summary[k]["unit"].lower()
    return num(k) * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
traffic = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
summary["derived"] = {"dram_traffic_bytes_per_launch": traffic,
                      "dram_GBps_during_capture": traffic / (num("gpu__time_duration.sum") * 1e-6) / 1e9
                      if summary["gpu__time_duration.sum"]["unit"] == "us" else None}
os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
json.dump(summary, open(out + ".json", "w"), indent=1)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     stdout=subprocess.PIPE, text=True).stdout
open("/tmp/_src.csv", "w").write(src)
here = os.path.dirname(os.path.abspath(__file__))
phases = subprocess.run([sys.executable, os.path.join(here, "ncu_phases.py"), "/tmp/_src.csv"], stdout=subprocess.PIPE, text=True).stdout
with open(out + ".md", "w") as f:
    f.write("# ncu --set full, kernel `%s`\n\nsource report: `%s` (scratch, not committed)\n\n" % (summary["Kernel Name"]["value"][:60], rep))
    f.write("| metric | value | unit |\n|---|---|---|\n")
    for k in KEYS[1:]:
        if k in summary:
            f.write("| %s | %s | %s |\n" % (k, summary[k]["value"], summary[k]["unit"]))
    f.write("| DRAM traffic per launch (read+write) | %.1f | MB |\n\n" % (traffic / 1e6))
    f.write("## instructions and stall samples per kernel phase (from the source page)\n\n```\n%s```\n" % phases)
print(json.dumps(summary["derived"]))
