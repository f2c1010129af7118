python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
timeout 900 python -m pytest tests/test_gpu_nstep.py -x -q 2>&1 | tail -4
timeout 600 python bench.py --workload train 2>/dev/null > gpurun_out/r06_bench_train.json; python -c "
import json; d=json.load(open('gpurun_out/r06_bench_train.json')); print('train value %.3e ms_per_step %.4f push %.4f step %.4f' % (d['value'], d['ms_per_step'], d['nstep']['ms_per_push'], d['roofline']['kernel_ms_per_launch']))"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4100 -c 30 --csv --log-file gpurun_out/r06_train_launches.csv python bench.py --workload train --steps 10 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r06_train_launches_run.log 2>&1; echo "ncu train rc=$?"; grep -c pve_ gpurun_out/r06_train_launches.csv
