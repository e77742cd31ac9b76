// Scratch micro-benchmark (not part of the library): what write bandwidth does each store path reach on
// this GPU for the FK kernel's traffic shape (a 28 B/cfg read stream + 64 B/cfg and 168 B/cfg write streams)?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <algorithm>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ void bulk_store(void* g, const void* s, uint32_t bytes) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(s);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// TMA: each CTA iterates tiles; writes T tile (8 KB) and J tile (21 KB); optional read of q tile (3.5 KB)
template <bool READQ, bool REFILL>
__global__ void __launch_bounds__(128) k_tma(const float* q, float* T, float* J, int64_t ntiles) {
    extern __shared__ __align__(128) unsigned char sm[];
    float* sJ = (float*)sm; float* sT = sJ + 128 * 42; float* sq = sT + 128 * 16;
    for (int i = threadIdx.x; i < 128 * 58; i += 128) sJ[i] = (float)i;
    bool inflight = false;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        if (READQ) {
            const float4* s = (const float4*)(q + t * 128 * 7);
            for (int i = threadIdx.x; i < 224; i += 128) ((float4*)sq)[i] = __ldg(s + i);
        }
        if (inflight && threadIdx.x == 0) bulk_wait_read();
        __syncthreads();
        if (REFILL) {
            float add = READQ ? sq[threadIdx.x] : 1.0f;
            for (int k = 0; k < 21; ++k) ((float2*)(sJ + threadIdx.x * 42))[k] = make_float2(add + k, add);
            for (int k = 0; k < 4; ++k) ((float4*)(sT + threadIdx.x * 16))[k] = make_float4(add, add, add, add);
        }
        fence_async();
        __syncthreads();
        if (threadIdx.x == 0) {
            bulk_store(T + t * 128 * 16, sT, 128 * 16 * 4);
            bulk_store(J + t * 128 * 42, sJ, 128 * 42 * 4);
            bulk_commit();
        }
        inflight = true;
    }
    if (inflight && threadIdx.x == 0) bulk_wait_read();
}

// mixed: J through TMA, T straight from registers (4 x st.global.v4 per thread, 64 B rows), q via plain loads
template <bool READQ>
__global__ void __launch_bounds__(128) k_mixed(const float* q, float* T, float* J, int64_t ntiles) {
    extern __shared__ __align__(128) unsigned char sm[];
    float* sJ = (float*)sm; float* sq = sJ + 128 * 42;
    bool inflight = false;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        float add = 1.0f;
        if (READQ) {
            const float4* s = (const float4*)(q + t * 128 * 7);
            for (int i = threadIdx.x; i < 224; i += 128) ((float4*)sq)[i] = __ldg(s + i);
            __syncthreads();
            add = sq[threadIdx.x * 7];
        }
        float4* pT = (float4*)(T + (t * 128 + threadIdx.x) * 16);
        for (int k = 0; k < 4; ++k) pT[k] = make_float4(add, add + k, add, add);
        if (inflight && threadIdx.x == 0) bulk_wait_read();
        __syncthreads();
        for (int k = 0; k < 21; ++k) ((float2*)(sJ + threadIdx.x * 42))[k] = make_float2(add + k, add);
        fence_async();
        __syncthreads();
        if (threadIdx.x == 0) { bulk_store(J + t * 128 * 42, sJ, 128 * 42 * 4); bulk_commit(); }
        inflight = true;
    }
    if (inflight && threadIdx.x == 0) bulk_wait_read();
}

// staged: both tiles staged in smem, then copied out with coalesced st.global.v4 (no TMA)
__global__ void __launch_bounds__(128) k_staged(float* T, float* J, int64_t ntiles) {
    extern __shared__ __align__(128) unsigned char sm[];
    float* sJ = (float*)sm;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        __syncthreads();
        for (int k = 0; k < 29; ++k) ((float2*)(sJ + threadIdx.x * 58))[k] = make_float2(1.f + k, 2.f);
        __syncthreads();
        float4* pT = (float4*)(T + t * 128 * 16);
        float4* pJ = (float4*)(J + t * 128 * 42);
        const float4* s4 = (const float4*)sJ;
        for (int i = threadIdx.x; i < 512; i += 128) pT[i] = s4[i];
        for (int i = threadIdx.x; i < 1344; i += 128) pJ[i] = s4[512 + i];
    }
}

// TMA, one CTA per tile (non persistent)
__global__ void __launch_bounds__(128) k_tma_np(float* T, float* J, int64_t ntiles) {
    extern __shared__ __align__(128) unsigned char sm[];
    float* sJ = (float*)sm; float* sT = sJ + 128 * 42;
    const int64_t t = blockIdx.x;
    for (int k = 0; k < 21; ++k) ((float2*)(sJ + threadIdx.x * 42))[k] = make_float2(1.f + k, 2.f);
    for (int k = 0; k < 4; ++k) ((float4*)(sT + threadIdx.x * 16))[k] = make_float4(1.f, 2.f, 3.f, 4.f);
    fence_async();
    __syncthreads();
    if (threadIdx.x == 0) {
        bulk_store(T + t * 128 * 16, sT, 128 * 16 * 4);
        bulk_store(J + t * 128 * 42, sJ, 128 * 42 * 4);
        bulk_commit();
        bulk_wait_read();
    }
}

// persistent with a dynamic (atomic) tile counter: keeps the write frontier compact like the HW scheduler does
__device__ unsigned long long g_next;
template <bool TMA>
__global__ void __launch_bounds__(128) k_dyn(float* T, float* J, int64_t ntiles) {
    extern __shared__ __align__(128) unsigned char sm[];
    float* sJ = (float*)sm; float* sT = sJ + 128 * 42;
    __shared__ long long s_t;
    bool inflight = false;
    for (;;) {
        if (threadIdx.x == 0) { s_t = (long long)atomicAdd(&g_next, 1ull); if (TMA && inflight) bulk_wait_read(); }
        __syncthreads();
        const int64_t t = s_t;
        if (t >= ntiles) break;
        if (TMA) {
            for (int k = 0; k < 21; ++k) ((float2*)(sJ + threadIdx.x * 42))[k] = make_float2(1.f + k, 2.f);
            for (int k = 0; k < 4; ++k) ((float4*)(sT + threadIdx.x * 16))[k] = make_float4(1.f, 2.f, 3.f, 4.f);
            fence_async();
            __syncthreads();
            if (threadIdx.x == 0) {
                bulk_store(T + t * 128 * 16, sT, 128 * 16 * 4);
                bulk_store(J + t * 128 * 42, sJ, 128 * 42 * 4);
                bulk_commit();
            }
            inflight = true;
        } else {
            float4 v = make_float4(1.f, 2.f, 3.f, (float)t);
            float4* pT = (float4*)(T + t * 128 * 16);
            float4* pJ = (float4*)(J + t * 128 * 42);
            for (int i = threadIdx.x; i < 512; i += 128) pT[i] = v;
            for (int i = threadIdx.x; i < 1344; i += 128) pJ[i] = v;
            __syncthreads();
        }
    }
    if (TMA && inflight && threadIdx.x == 0) bulk_wait_read();
}

// LSU: same tiles written with coalesced st.global.v4 straight from registers
__global__ void __launch_bounds__(128) k_lsu(float* T, float* J, int64_t ntiles) {
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        float4 v = make_float4(1.f, 2.f, 3.f, (float)t);
        float4* pT = (float4*)(T + t * 128 * 16);
        float4* pJ = (float4*)(J + t * 128 * 42);
        for (int i = threadIdx.x; i < 512; i += 128) pT[i] = v;
        for (int i = threadIdx.x; i < 1344; i += 128) pJ[i] = v;
    }
}

int main() {
    const int64_t B = 1 << 24, ntiles = B / 128;
    float *q, *T, *J;
    CK(cudaMalloc(&q, B * 7 * 4)); CK(cudaMalloc(&T, B * 16 * 4)); CK(cudaMalloc(&J, B * 42 * 4));
    CK(cudaMemset(q, 0, B * 7 * 4));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const size_t smem = 4 * 128 * (42 + 16 + 7);
    auto run = [&](const char* name, auto launch, double bytes) {
        float best = 1e9;
        for (int r = 0; r < 8; ++r) {
            cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b); if (r >= 2) best = std::min(best, ms);
        }
        printf("%-44s %.3f ms  %.0f GB/s  (%s)\n", name, best, bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    };
    const double wb = (double)B * 232, rb = (double)B * 28;
    for (int per_sm : {2, 4, 6}) {
        const int grid = 148 * per_sm;
        printf("--- %d CTAs/SM\n", per_sm);
        cudaFuncSetAttribute(k_tma<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_tma<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_tma<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        run("TMA bulk store only (no smem refill)", [&] { k_tma<false, false><<<grid, 128, smem>>>(q, T, J, ntiles); }, wb);
        run("TMA bulk store + smem refill", [&] { k_tma<false, true><<<grid, 128, smem>>>(q, T, J, ntiles); }, wb);
        run("TMA bulk store + refill + q read", [&] { k_tma<true, true><<<grid, 128, smem>>>(q, T, J, ntiles); }, wb + rb);
        run("LSU st.global.v4 coalesced", [&] { k_lsu<<<grid, 128>>>(T, J, ntiles); }, wb);
        cudaFuncSetAttribute(k_mixed<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_mixed<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_staged, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        run("mixed: J TMA + T direct STG.128", [&] { k_mixed<false><<<grid, 128, smem>>>(q, T, J, ntiles); }, wb);
        run("mixed + q read", [&] { k_mixed<true><<<grid, 128, smem>>>(q, T, J, ntiles); }, wb + rb);
        run("staged smem -> coalesced STG.128", [&] { k_staged<<<grid, 128, smem>>>(T, J, ntiles); }, wb);
    }
    for (int per_sm : {8, 12, 16}) {
        const int grid = 148 * per_sm;
        printf("--- %d CTAs/SM (small smem variants)\n", per_sm);
        run("mixed: J TMA + T direct STG.128", [&] { k_mixed<false><<<grid, 128, 4 * 128 * 49>>>(q, T, J, ntiles); }, wb);
        run("mixed + q read", [&] { k_mixed<true><<<grid, 128, 4 * 128 * 49>>>(q, T, J, ntiles); }, wb + rb);
        run("LSU st.global.v4 coalesced", [&] { k_lsu<<<grid, 128>>>(T, J, ntiles); }, wb);
    }
    cudaFuncSetAttribute(k_tma_np, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    run("TMA, one CTA per tile", [&] { k_tma_np<<<(unsigned)ntiles, 128, 4 * 128 * 58>>>(T, J, ntiles); }, wb);
    for (int per_sm : {4, 7}) {
        const int grid = 148 * per_sm;
        unsigned long long zero = 0;
        cudaFuncSetAttribute(k_dyn<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        run("dynamic-counter persistent, TMA", [&] { cudaMemcpyToSymbol(g_next, &zero, 8); k_dyn<true><<<grid, 128, 4 * 128 * 58>>>(T, J, ntiles); }, wb);
        run("dynamic-counter persistent, LSU", [&] { cudaMemcpyToSymbol(g_next, &zero, 8); k_dyn<false><<<grid, 128, 4 * 128 * 58>>>(T, J, ntiles); }, wb);
    }
    run("LSU st.global.v4, one CTA per tile", [&] { k_lsu<<<(unsigned)ntiles, 128>>>(T, J, ntiles); }, wb);
    return 0;
}
