#!/bin/bash
# On the GPU box: phase timing of the current source + one ncu --set full capture of the step kernel.
# Usage: tools/gpu_prof.sh tag [library]
TAG=$1; LIB=${2:-}
mkdir -p gpurun_out
python tools/phase_timing.py 128 > gpurun_out/${TAG}_phase.txt 2>&1; head -22 gpurun_out/${TAG}_phase.txt
[ -n "$LIB" ] && export PVE_MCC_LIBRARY=$PWD/$LIB
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:pve_step_kernel -s 415 -c 1 -f -o gpurun_out/${TAG}_prof \
    python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_full_run.log 2>&1; echo "ncu full rc=$?"; ls -la gpurun_out/${TAG}_prof.ncu-rep
