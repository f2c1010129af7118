#!/bin/bash
# Build the current csrc/ into .ab/lib<name>.so (A/B measurement builds; .ab/ is git-ignored but travels with gpurun).
# Usage: tools/ab_build.sh name [extra nvcc flags...]
set -e
NAME=$1; shift
mkdir -p .ab
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -shared -Xcompiler -fPIC "$@" \
    -I include -I pve_mcc_for_unsignalized_intersection_b200/csrc pve_mcc_for_unsignalized_intersection_b200/csrc/pve_mcc.cu -o .ab/lib$NAME.so
echo built .ab/lib$NAME.so
