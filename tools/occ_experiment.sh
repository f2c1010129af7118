#!/bin/bash
# how does the step kernel scale with resident CTAs per SM?  (pads dynamic shared memory; class 128/80 = 26.7 KB)
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
for pad in 0 2200 5600 10400 17600 29600 53000 120000; do
  r=$(PVE_SMEM_PAD=$pad python bench.py --threads ${1:-128} --steps 40 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['roofline']['kernel_ms_per_launch'], d['config']['smem_per_cta'])")
  echo "threads ${1:-128} pad $pad kernel_ms smem $r"
done
