#!/bin/bash
# On the GPU box: bench every .ab/lib*.so (or the names given) with the default workload; prints ms per launch.
# Usage: tools/ab_run.sh [-t] [name...]    -t: also run the GPU parity tests with each library
TESTS=0; if [ "$1" = "-t" ]; then TESTS=1; shift; fi
NAMES=${*:-$(ls .ab/lib*.so | sed 's#.ab/lib##; s#\.so##')}
mkdir -p gpurun_out
for n in $NAMES; do
  if [ $TESTS = 1 ]; then
    PVE_MCC_LIBRARY=$PWD/.ab/lib$n.so timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/ab_${n}_tests.log 2>&1; echo "$n tests rc=$? $(tail -1 gpurun_out/ab_${n}_tests.log)"
  fi
  for rep in 1 2; do
  PVE_MCC_LIBRARY=$PWD/.ab/lib$n.so timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e ${AB_BENCH_ARGS:-} > gpurun_out/ab_$n.json 2> gpurun_out/ab_$n.err
  python - <<P
import json
try:
    d = json.load(open("gpurun_out/ab_$n.json"))
    print("%-14s ms/step %.4f kernel %.4f frac %.3f clocks %s %s" % ("$n", d["ms_per_step"], d["roofline"]["kernel_ms_per_launch"], d["roofline"]["frac"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
except Exception as e:
    print("$n FAILED", e); print(open("gpurun_out/ab_$n.err").read()[-800:])
P
  done
done
