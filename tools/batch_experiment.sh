#!/bin/bash
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
for th in 64 128; do for envs in 2048 4096 8192 16384 32768; do
  r=$(python bench.py --threads $th --envs $envs --steps 30 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'])")
  echo "threads $th envs $envs kernel_ms frac $r"
done; done
