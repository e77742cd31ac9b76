// nbx_knn.cu -- exact brute-force k nearest neighbours in L2 for sm_100a.
//
// Replaces, for the batched PRM construction, the faiss `IndexFlatL2.search` behind reference
// numbotics/math/geometry/nearest_neighbors.py:6-85 (`ApproximateNearestNeighborsIndex.query / k_nearest`) as it is used by
// numbotics/planning/sampling_based/planners/prm.py:35-40: every new vertex asks for its k nearest EARLIER vertices.
//
// One thread per query keeps its k best (distance, index) pairs sorted in registers; the reference points stream through
// shared memory in tiles that all threads of the CTA read with the same address (broadcast), so a pair costs two 16-byte
// shared-memory loads per 8 dimensions plus 2 flops per dimension.  A candidate enters the list only if it beats the current
// worst entry (after the first few hundred candidates that is rare: ~k ln(N/k) insertions per query), by a branch-free
// bubble pass over the statically indexed list.  `causal` restricts query i to references j < i (the incremental roadmap);
// the triangular work is balanced by pairing heavy and light query blocks in one CTA.  FP32-issue bound: N^2/2 * ~20 instructions for the roadmap of N vertices.
#include "nbx_common.cuh"
#include <float.h>

namespace {

constexpr int kKnnThreads = 128;
constexpr int kKnnTileFloats = 8192; // shared-memory tile: 32 KB of reference points (1024 points at D = 8, 128 at D = 64)

template <int D, int K>
__global__ void __launch_bounds__(kKnnThreads)
knn_kernel(const float* __restrict__ refs, int64_t n_refs, const float* __restrict__ queries, int64_t n_q, int dof, int k,
           int causal, int64_t* __restrict__ idx_out, uint8_t* __restrict__ valid_out, float* __restrict__ d2_out) {
    constexpr int kKnnTile = kKnnTileFloats / D;
    __shared__ __align__(16) float tile[kKnnTile * D];
    // Causal mode is triangular (query i scans i references): a CTA takes the query blocks b and nblocks - 1 - b, one after
    // the other, so that every CTA scans the same number of references (one wave of CTAs, all of the same length).
    const int64_t nblocks = (n_q + kKnnThreads - 1) / kKnnThreads;
    const int passes = causal ? 2 : 1;
    for (int pass = 0; pass < passes; ++pass) {
    int64_t blk = (int64_t)blockIdx.x;
    if (causal) {
        blk = pass == 0 ? nblocks - 1 - (int64_t)blockIdx.x : (int64_t)blockIdx.x;
        if (pass == 1 && blk >= nblocks - 1 - (int64_t)blockIdx.x) break;     // odd count: the middle block was pass 0
    }
    const int64_t qb = blk * kKnnThreads;
    const int64_t qi = qb + threadIdx.x;
    const bool have = qi < n_q;
    float qv[D];
#pragma unroll
    for (int d = 0; d < D; ++d) qv[d] = (have && d < dof) ? queries[qi * dof + d] : 0.f;
    float bd[K];
    int bi[K];
#pragma unroll
    for (int s = 0; s < K; ++s) { bd[s] = FLT_MAX; bi[s] = -1; }
    float worst = FLT_MAX;           // the k-th best distance so far (entries beyond k are never reported)
    // references this CTA has to look at: all of them, or those before its last query
    const int64_t r_end = causal ? min(n_refs, qb + kKnnThreads) : n_refs;
    const int64_t my_end = causal ? min(n_refs, qi) : n_refs;
    for (int64_t r0 = 0; r0 < r_end; r0 += kKnnTile) {
        const int nt = (int)min((int64_t)kKnnTile, r_end - r0);
        __syncthreads();
        for (int i = threadIdx.x; i < nt * D; i += kKnnThreads) {
            const int r = i / D, d = i - r * D;
            tile[i] = d < dof ? refs[(r0 + r) * dof + d] : 0.f;
        }
        __syncthreads();
        if (!have) continue;
        const int lim = (int)max((int64_t)0, min((int64_t)nt, my_end - r0));
#pragma unroll 4
        for (int r = 0; r < lim; ++r) {
            const float4* t4 = reinterpret_cast<const float4*>(tile + r * D);
            float da = 0.f, db = 0.f;          // two partial sums: half the dependent FFMA chain
#pragma unroll
            for (int d = 0; d < D / 4; ++d) {
                const float4 c = t4[d];
                const float e0 = qv[4 * d] - c.x, e1 = qv[4 * d + 1] - c.y, e2 = qv[4 * d + 2] - c.z, e3 = qv[4 * d + 3] - c.w;
                da = fmaf(e0, e0, da); db = fmaf(e1, e1, db); da = fmaf(e2, e2, da); db = fmaf(e3, e3, db);
            }
            const float d2 = da + db;
            if (d2 < worst) {
                float cd = d2;
                int ci = (int)(r0 + r);
#pragma unroll
                for (int s = 0; s < K; ++s) {
                    if (cd < bd[s]) {             // strict: of equal distances the earlier reference stays in front
                        const float td = bd[s]; bd[s] = cd; cd = td;
                        const int ti = bi[s]; bi[s] = ci; ci = ti;
                    }
                }
                // k-th entry.  k <= K is a runtime value and the list must stay statically indexed (a dynamic bd[k - 1]
                // would move it to local memory): the list ascends, so the k-th entry is the minimum of the entries from k - 1 on
                float w = FLT_MAX;
#pragma unroll
                for (int s = 0; s < K; ++s) w = fminf(w, s >= k - 1 ? bd[s] : FLT_MAX);
                worst = w;
            }
        }
    }
    if (have) {
#pragma unroll
        for (int s = 0; s < K; ++s) {
            if (s < k) {
                const bool ok = bi[s] >= 0;
                idx_out[qi * k + s] = ok ? (int64_t)bi[s] : 0;
                if (valid_out != nullptr) valid_out[qi * k + s] = ok ? 1 : 0;
                if (d2_out != nullptr) d2_out[qi * k + s] = ok ? bd[s] : INFINITY;
            }
        }
    }
    }   // pass
}

template <int D>
int launch_knn_d(const float* refs, int64_t n_refs, const float* queries, int64_t n_q, int dof, int k, int causal,
                 int64_t* idx, uint8_t* valid, float* d2, cudaStream_t stream) {
    const int64_t nblocks = (n_q + kKnnThreads - 1) / kKnnThreads;
    const int64_t grid = causal ? (nblocks + 1) / 2 : nblocks;      // causal: two query blocks per CTA (see the kernel)
    NBX_REQUIRE(grid <= 0x7fffffffLL, "too many queries for one launch");
    if (k <= 16) knn_kernel<D, 16><<<(unsigned)grid, kKnnThreads, 0, stream>>>(refs, n_refs, queries, n_q, dof, k, causal, idx, valid, d2);
    else knn_kernel<D, 32><<<(unsigned)grid, kKnnThreads, 0, stream>>>(refs, n_refs, queries, n_q, dof, k, causal, idx, valid, d2);
    nbx_count_launch();
    NBX_CUDA_OK(cudaGetLastError());
    return NBX_OK;
}

}  // namespace

extern "C" int nbx_knn_f32(const float* refs, int64_t n_refs, const float* queries, int64_t n_queries, int dof, int k, int causal,
                           int64_t* idx, uint8_t* valid, float* dist2, int device, void* stream) {
    NBX_REQUIRE(n_refs >= 0 && n_queries >= 0, "negative size");
    NBX_REQUIRE(dof >= 1 && dof <= 64, "dimension %d out of range [1, 64]", dof);
    NBX_REQUIRE(k >= 1 && k <= 32, "k = %d out of range [1, 32]", k);
    if (n_queries == 0) return NBX_OK;
    NBX_REQUIRE(queries != nullptr && idx != nullptr && (refs != nullptr || n_refs == 0), "NULL buffer");
    NBX_REQUIRE(n_refs <= 0x7fffffffLL, "too many reference points");
    NBX_CUDA_OK(cudaSetDevice(device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dof <= 8) return launch_knn_d<8>(refs, n_refs, queries, n_queries, dof, k, causal, idx, valid, dist2, st);
    if (dof <= 16) return launch_knn_d<16>(refs, n_refs, queries, n_queries, dof, k, causal, idx, valid, dist2, st);
    if (dof <= 32) return launch_knn_d<32>(refs, n_refs, queries, n_queries, dof, k, causal, idx, valid, dist2, st);
    return launch_knn_d<64>(refs, n_refs, queries, n_queries, dof, k, causal, idx, valid, dist2, st);
}
