#!/bin/bash
# On the GPU box: one ncu --set full capture of an actor kernel launch at the bench workload.  Usage: tools/gpu_ncu_actor.sh tag [kernel regex]
TAG=$1; K=${2:-pve_actor_tc_kernel}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 405 -c 1 -f -o gpurun_out/${TAG}_actor_prof \
    python tools/actor_timing.py > gpurun_out/${TAG}_actor_run.log 2>&1; echo "ncu rc=$?"; ls -la gpurun_out/${TAG}_actor_prof.ncu-rep; tail -3 gpurun_out/${TAG}_actor_run.log
