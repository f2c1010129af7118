#!/bin/bash
# A/B of an environment knob in one GPU session: tools/ab_experiment.sh VAR val_a val_b [bench args]
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
var=$1; a=$2; b=$3; shift 3
for rep in 1 2; do for val in $a $b; do
  env $var=$val python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$var=$val', 'kernel_ms', round(d['roofline']['kernel_ms_per_launch'],5), 'frac', round(d['roofline']['frac'],4), 'ms_per_step', round(d['ms_per_step'],5), 'overflow', d['stats']['overflow'])"
done; done
