#!/bin/bash
# BASELINE config 5: env-count sweep of the full rollout (GPU actor + environment step per tick) on one GPU.
# Usage (GPU box): tools/env_sweep.sh [tag] [workload]   -> gpurun_out/<tag>_env_sweep.jsonl (one bench line per size)
TAG=${1:-r01}; WL=${2:-rollout}
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
OUT=gpurun_out/${TAG}_${WL}_env_sweep.jsonl; : > $OUT
for n in 1 16 256 1024 4096 16384 65536 262144; do
  steps=100; [ $n -ge 65536 ] && steps=30
  timeout 900 python bench.py --workload $WL --envs $n --steps $steps --warmup 5 --no-cpu-baseline --no-e2e 2>> gpurun_out/${TAG}_${WL}_env_sweep.err >> $OUT
  tail -1 $OUT | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('envs', $n, 'value %.3e' % d['value'], 'ms_per_step %.4f' % d['ms_per_step'], 'step kernel ms %.4f' % d['roofline']['kernel_ms_per_launch'], 'frac %.3f' % d['roofline']['frac'], 'overflow', d['stats']['overflow'])"
done
