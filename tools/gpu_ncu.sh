#!/bin/bash
# On the GPU box: one ncu --set full capture of the step kernel of an A/B library.  Usage: tools/gpu_ncu.sh tag lib
TAG=$1; LIB=${2:-}
mkdir -p gpurun_out
[ -n "$LIB" ] && export PVE_MCC_LIBRARY=$PWD/$LIB
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:pve_step_kernel -s 415 -c 1 -f -o gpurun_out/${TAG}_prof \
    python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-e2e --no-traffic > gpurun_out/${TAG}_full_run.log 2>&1; echo "ncu full rc=$?"; ls -la gpurun_out/${TAG}_prof.ncu-rep
