#!/bin/bash
# BASELINE config 5: env-count sweep of the full rollout (GPU actor + environment step per tick) on one GPU, 1 ... 1 048 576
# intersections.  The two largest sizes run in the 192/128 capacity class (no deferred arrival in 1e8 intersection-ticks) with
# the dense outputs sized for 80 rows per intersection.
# Usage (GPU box): tools/env_sweep.sh [tag] [workload]   -> gpurun_out/<tag>_env_sweep.jsonl (one bench line per size)
TAG=${1:-r01}; WL=${2:-rollout}
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
OUT=gpurun_out/${TAG}_${WL}_env_sweep.jsonl; : > $OUT
for n in ${SWEEP_SIZES:-1 16 256 1024 4096 16384 65536 262144 1048576}; do
  steps=100; extra=""
  [ $n -ge 65536 ] && steps=30
  [ $n -ge 262144 ] && extra="--veh-cap 192 --agent-cap 128 --out-rows-per-env 80" && steps=20
  timeout 1500 python bench.py --workload $WL --envs $n --steps $steps --warmup 5 --no-cpu-baseline --no-e2e $extra 2>> gpurun_out/${TAG}_${WL}_env_sweep.err >> $OUT
  tail -1 $OUT | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('envs', $n, 'value %.3e' % d['value'], 'ms_per_step %.4f' % d['ms_per_step'], 'step kernel ms %.4f' % d['roofline']['kernel_ms_per_launch'], 'frac %.3f' % d['roofline']['frac'], 'class', d['config']['veh_cap'], d['config']['agent_cap'], 'overflow', d['stats']['overflow'])"
done
