python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
for rep in 1 2; do python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('default kernel_ms', round(d['roofline']['kernel_ms_per_launch'],5), 'frac', round(d['roofline']['frac'],4))"; done
timeout 900 python -m pytest tests/test_gpu_nstep.py tests/test_gpu_parity.py -x -q -k "distinct or neighbour" 2>&1 | tail -3
timeout 600 python bench.py --workload train --steps 100 --warmup 5 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('train value %.3e ms_per_step %.4f push %.4f step %.4f' % (d['value'], d['ms_per_step'], d['nstep']['ms_per_push'], d['roofline']['kernel_ms_per_launch']))"
