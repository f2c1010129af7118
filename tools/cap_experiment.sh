#!/bin/bash
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
for cap in 96 80; do
  PVE_BENCH_AGENT_CAP=$cap python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('agent_cap', d['config']['agent_cap'], 'smem', d['config']['smem_per_cta'], 'kernel_ms', d['roofline']['kernel_ms_per_launch'], 'frac', d['roofline']['frac'], 'ms_per_step', d['ms_per_step'], 'overflow', d['stats']['overflow'])"
done
