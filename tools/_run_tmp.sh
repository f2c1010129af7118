python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
run() { env PVE_PREFETCH_DIST=$1 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('dist=$1', 'kernel_ms', round(d['roofline']['kernel_ms_per_launch'],5), 'frac', round(d['roofline']['frac'],4), 'ms_per_step', round(d['ms_per_step'],5))"; }
for rep in 1 2; do for d in 0 1036 518 2072; do run $d; done; done
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
