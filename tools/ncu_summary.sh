#!/bin/bash
# headline metrics of a capture: tools/ncu_summary.sh X.ncu-rep
ncu -i $1 --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin))
h=r[0]; v=r[2] if len(r)>2 else r[1]
want=['gpu__time_duration.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','launch__registers_per_thread','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','launch__shared_mem_per_block_dynamic','dram__bytes_read.sum','dram__bytes_write.sum','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','launch__waves_per_multiprocessor','sm__maximum_warps_per_active_cycle_pct','launch__occupancy_limit_warps','smsp__warps_eligible.avg.per_cycle_active']
for w in want:
    if w in h: print('%-90s %s' % (w, v[h.index(w)]))
"
