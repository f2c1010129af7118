// Microbenchmark behind the host-buffer path (pve_step_host): how fast can SM stores deliver one tick's per-agent outputs
// to pinned host memory, as five separate arrays (4 + 16 + 4 + 1 + 4 bytes per agent: today's layout) or as one 32-byte
// record per agent?  4096 CTAs ("intersections") x 45 agents.  nvcc -arch=sm_100a -O3 pcie_write_bench.cu -o pcie_write_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void separate(float *reward, int4 *ids, int *cpv, uint8_t *status, float *jerk, int A) {
    const int b = blockIdx.x, g = threadIdx.x;
    if (g < A) {
        const size_t r = (size_t)b * A + g;
        reward[r] = (float)g; ids[r] = make_int4(b, g, g, b + g); cpv[r] = 0; status[r] = 1; jerk[r] = 2.f;
    }
}
__global__ void packed(int4 *rec, int A) {          // 2 x 16 B per agent, lanes follow the contiguous block
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < 2 * A; i += blockDim.x) rec[(size_t)b * 2 * A + i] = make_int4(b, i, i, b + i);
}
__global__ void separate_dev_then(float *reward, int A) { (void)reward; (void)A; }

int main() {
    const int B = 4096, A = 45, reps = 50;
    const size_t n = (size_t)B * A;
    float *reward, *jerk; int4 *ids, *rec; int *cpv; uint8_t *status;
    cudaHostAlloc(&reward, n * 4, 0); cudaHostAlloc(&jerk, n * 4, 0); cudaHostAlloc(&ids, n * 16, 0);
    cudaHostAlloc(&cpv, n * 4, 0); cudaHostAlloc(&status, n, 0); cudaHostAlloc(&rec, n * 32, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int pass = 0; pass < 2; ++pass) {
        cudaEventRecord(e0);
        for (int i = 0; i < reps; ++i) separate<<<B, 64>>>(reward, ids, cpv, status, jerk, A);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        if (pass) printf("separate arrays : %.4f ms per tick, %.1f GB/s (29 B/agent)\n", ms / reps, n * 29.0 / (ms / reps * 1e-3) / 1e9);
        cudaEventRecord(e0);
        for (int i = 0; i < reps; ++i) packed<<<B, 64>>>(rec, A);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        if (pass) printf("packed records  : %.4f ms per tick, %.1f GB/s (32 B/agent)\n", ms / reps, n * 32.0 / (ms / reps * 1e-3) / 1e9);
    }
    // the same bytes by one DMA copy from device memory
    int4 *drec; cudaMalloc(&drec, n * 32);
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) cudaMemcpyAsync(rec, drec, n * 32, cudaMemcpyDeviceToHost);
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("DMA copy        : %.4f ms per tick, %.1f GB/s (32 B/agent)\n", ms / reps, n * 32.0 / (ms / reps * 1e-3) / 1e9);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
