#!/bin/bash
# On the GPU box: one ncu --set full capture of the 4-/8-lane step kernel.  Usage: tools/gpu_ncu_lane.sh tag lane4|lane8
TAG=$1; W=${2:-lane4}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pve4_step_kernel -s 420 -c 1 -f -o gpurun_out/${TAG}_${W}_prof \
    python bench.py --workload $W --steps 10 --warmup 5 --no-cpu-baseline --no-e2e --no-traffic > gpurun_out/${TAG}_${W}_run.log 2>&1; echo "ncu rc=$?"; ls -la gpurun_out/${TAG}_${W}_prof.ncu-rep
