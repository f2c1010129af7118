#!/bin/bash
# Round-2 evidence run on the GPU box: full bench, launch list, ncu --set full of the small step kernel, sanitizer.
# Usage: tools/gpu_round2.sh [what...]   what = bench launches full stress rollout lanes train actor tests smoke sanitizer sanitizer2
OUT=gpurun_out; mkdir -p $OUT
WHAT=${*:-bench launches full}
for w in $WHAT; do
case $w in
bench)
  timeout 900 python bench.py > $OUT/r02_bench.json 2> $OUT/r02_bench.err; echo "bench rc=$?"; tail -2 $OUT/r02_bench.err; cut -c1-400 $OUT/r02_bench.json
  timeout 600 python bench.py --impl reference --steps 100 --warmup 5 > $OUT/r02_bench_reference_arm.json 2> $OUT/r02_bench_ref.err; cut -c1-200 $OUT/r02_bench_reference_arm.json ;;
launches)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 830 -c 60 --csv --log-file $OUT/r02_launches.csv \
      python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-traffic > $OUT/r02_launches_run.log 2>&1; echo "ncu launches rc=$?"; tail -4 $OUT/r02_launches.csv ;;
full)
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:pve_step_kernel -s 415 -c 1 -f -o $OUT/r02_prof \
      python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-e2e --no-traffic > $OUT/r02_full_run.log 2>&1; echo "ncu full rc=$?"; ls -la $OUT/r02_prof.ncu-rep ;;
stress)
  timeout 900 python bench.py --workload stress --steps 50 --warmup 5 --no-cpu-baseline > $OUT/r02_bench_stress.json 2> $OUT/r02_bench_stress.err; cut -c1-300 $OUT/r02_bench_stress.json ;;
rollout)
  timeout 900 python bench.py --workload rollout --steps 50 --warmup 5 --no-cpu-baseline > $OUT/r02_bench_rollout.json 2> $OUT/r02_bench_rollout.err; cut -c1-300 $OUT/r02_bench_rollout.json ;;
lanes)
  for w in lane4 lane8; do
    timeout 900 python bench.py --workload $w --steps 100 --warmup 5 > $OUT/r02_bench_$w.json 2> $OUT/r02_bench_$w.err; cut -c1-260 $OUT/r02_bench_$w.json
  done ;;
train)
  timeout 900 python bench.py --workload train --steps 50 --warmup 5 --no-cpu-baseline --no-e2e > $OUT/r02_bench_train.json 2> $OUT/r02_bench_train.err; cut -c1-300 $OUT/r02_bench_train.json ;;
actor)
  for impl in tc5 mma; do echo "== actor timing $impl"; PVE_ACTOR_IMPL=$impl timeout 300 python tools/actor_timing.py 2>&1 | tail -5; done > $OUT/r02_actor_timing.txt; cat $OUT/r02_actor_timing.txt
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:pve_actor_tc_kernel -s 405 -c 1 -f -o $OUT/r02_actor_prof \
      python tools/actor_timing.py > $OUT/r02_actor_run.log 2>&1; echo "ncu actor rc=$?"; ls -la $OUT/r02_actor_prof.ncu-rep ;;
tests)
  timeout 1700 python -m pytest tests -x -q -m gpu > $OUT/r02_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 $OUT/r02_gpu_tests.log ;;
sanitizer2)   # the round's new kernels: tcgen05 actor / critic, frame-log fold
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_actor.py tests/test_gpu_nstep.py -x -q -k "(tc5 and (shapes or critic_kernel or recorded_rows)) or frame_log" > $OUT/r02_sanitizer_memcheck_tc5.log 2>&1; echo "memcheck tc5 rc=$?"; tail -5 $OUT/r02_sanitizer_memcheck_tc5.log ;;
smoke)
  timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 ;;
sanitizer)
  timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity_more.py tests/test_gpu_lane4.py -x -q -k "pipelined or two_handles or out_cap or dual or lane8_matches or rollout8_direct" > $OUT/r02_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 $OUT/r02_sanitizer_memcheck.log
  timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "teacher_forced_every_tick" > $OUT/r02_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -8 $OUT/r02_sanitizer_racecheck.log ;;
esac
done
