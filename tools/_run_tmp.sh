python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --workload stress --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r03e_bench_stress_${N}gpu.json 2> gpurun_out/r03e_bench_stress_${N}gpu.err; echo "rc=$?"; tail -2 gpurun_out/r03e_bench_stress_${N}gpu.err; cut -c1-200 gpurun_out/r03e_bench_stress_${N}gpu.json
