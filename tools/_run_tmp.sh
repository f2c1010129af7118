nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/micro/pcie_write_bench.cu -o /tmp/pcie_write_bench && /tmp/pcie_write_bench
