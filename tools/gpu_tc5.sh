#!/bin/bash
# the tcgen05 actor / critic on the GPU: parity tests (under a timeout), then device time of the actor kernel per implementation
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_actor.py tests/test_gpu_nstep.py -x -q -k "tc5 or closed_loop or recorded_reference or folder_against" 2>&1 | tail -6
for impl in tc5 mma; do echo "== timing $impl"; PVE_ACTOR_IMPL=$impl timeout 300 python tools/actor_timing.py 2>&1 | tail -5; done
