#!/bin/bash
# Runs on the GPU box (via gpurun): GPU tests, bench, ncu launch list, one full ncu capture.
# Usage: tools/gpu_check.sh [tag] [what...]   what = tests bench launches full
set -u
TAG=${1:-r01}; shift || true
WHAT=${*:-tests bench launches full}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > $OUT/${TAG}_build.log 2>&1 || { echo BUILD FAILED; tail -20 $OUT/${TAG}_build.log; exit 1; }
for w in $WHAT; do
case $w in
tests)
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/${TAG}_pytest_gpu.log ;;
smoke)
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -5 $OUT/${TAG}_smoke.log ;;
bench)
  timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; tail -3 $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench.json ;;
benchref)
  timeout 900 python bench.py --impl reference --steps 100 --warmup 5 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; cat $OUT/${TAG}_bench_ref.json ;;
sweep)
  for th in 64 128 256; do timeout 600 python bench.py --threads $th --steps 60 --warmup 5 --no-cpu-baseline --no-e2e > $OUT/${TAG}_bench_t$th.json 2>> $OUT/${TAG}_sweep.err; cat $OUT/${TAG}_bench_t$th.json; done ;;
launches)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 430 -c 40 --csv --log-file $OUT/${TAG}_launches.csv \
      python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > $OUT/${TAG}_launches_run.log 2>&1; echo "ncu launches rc=$?"; tail -8 $OUT/${TAG}_launches.csv ;;
full)
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:pve_step_kernel -s 415 -c 2 -f -o $OUT/${TAG}_prof \
      python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-e2e > $OUT/${TAG}_full_run.log 2>&1; echo "ncu full rc=$?"; ls -la $OUT/${TAG}_prof.ncu-rep ;;
esac
done
