// Same-address global atomicAdd (with return value) under the contention pattern of the collision kernels: W warps, lane 0
// of each issues N dependent atomics on ONE counter (or on one of K counters).  Prints ns per atomic (whole grid) and the
// latency one warp sees.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/atomic_bench.cu -o scripts/atomic_bench
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* ctr, int n, int n_ctr, long long* sink, int work) {
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int acc = 0;
    float x = threadIdx.x;
    for (int i = 0; i < n; ++i) {
        int c = 0;
        if (lane == 0) c = atomicAdd(ctr + (w % n_ctr) * 32, 1);
        c = __shfl_sync(0xffffffffu, c, 0);
        acc += c;
        for (int j = 0; j < work; ++j) x = fmaf(x, 1.0001f, 0.5f);       // independent work between atomics
    }
    if (acc == -1 || x == 123.f) sink[0] = acc;
}
int main() {
    int* ctr; long long* sink;
    cudaMalloc(&ctr, 4096 * 4); cudaMalloc(&sink, 8);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int blocks = 592, threads = 128, warps = blocks * threads / 32;
    for (int work : {0, 2000, 20000}) for (int n_ctr : {1, 2, 16}) for (int n : {1, 4, 32}) {
        cudaMemset(ctr, 0, 4096 * 4);
        k<<<blocks, threads>>>(ctr, n, n_ctr, sink, work); cudaDeviceSynchronize();
        cudaMemset(ctr, 0, 4096 * 4);
        cudaEventRecord(a); k<<<blocks, threads>>>(ctr, n, n_ctr, sink, work); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("work %5d fma  counters %2d  atomics/warp %2d: %8.2f us total, %6.2f ns per atomic (grid), %6.2f us per warp-iteration\n",
               work, n_ctr, n, ms * 1e3, ms * 1e6 / ((double)warps * n), ms * 1e3 / n);
    }
    return 0;
}
