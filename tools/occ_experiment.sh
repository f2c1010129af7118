#!/bin/bash
# how does the step kernel scale with resident CTAs per SM?  (pads dynamic shared memory)
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
for th in 64 128; do for pad in 0 4000 12000 21000 32000 50000; do
  r=$(PVE_SMEM_PAD=$pad python bench.py --threads $th --steps 40 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['roofline']['kernel_ms_per_launch'])")
  echo "threads $th pad $pad kernel_ms $r"
done; done
