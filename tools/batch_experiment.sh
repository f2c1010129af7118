#!/bin/bash
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
for envs in 1184 2368 4096 4736 8192 16384; do
  r=$(python bench.py --threads ${1:-128} --envs $envs --steps 30 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'])")
  echo "threads ${1:-128} envs $envs kernel_ms frac $r"
done
