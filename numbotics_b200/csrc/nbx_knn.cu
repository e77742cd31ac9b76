// nbx_knn.cu -- exact brute-force k nearest neighbours in L2 for sm_100a.
//
// Replaces, for the batched PRM construction, the faiss `IndexFlatL2.search` behind reference
// numbotics/math/geometry/nearest_neighbors.py:6-85 (`ApproximateNearestNeighborsIndex.query / k_nearest`) as it is used by
// numbotics/planning/sampling_based/planners/prm.py:35-40: every new vertex asks for its k nearest EARLIER vertices.
//
// A CTA of four warps serves 32 queries: lane = query, warp = SLICE -- the four warps scan interleaved quarters of every
// reference tile for the same queries, each keeping its own k best (distance, index) pairs sorted in registers, and every
// few tiles the slices agree on the exact k-th best distance of the union of their lists (a four-way merge of the sorted
// lists through shared memory: ~12 k instructions against ~5 000 for the pairs in between), so a slice prunes with the
// bound a single scanner would have had at that point.  (The first version gave a query one thread for the whole scan:
// the roadmap's 3 125 warps each walked 10^5 references in sequence -- 10 warps per SM on average, every instruction
// waiting on its predecessor: ncu issue 0.60, `wait` 1.26 stalls per issue.  Slices cut that critical path by four and
// double the resident warps.)  The reference points are packed once (knn_pack_kernel) and stream through shared memory in
// tiles fetched by 16-byte cp.async, the next tile in flight while the current one is scanned; all lanes read a reference at
// the same address (broadcast).  A reference that passes the test against the bound is only noted in a per-lane pending
// list; the lists are drained -- exact distance, exact test, insertion -- by all lanes together (see the scan loop).
//
// Per pair, when the points have a spare component (dof < D, e.g. the 7-DoF arm in 8 floats): the tile stores
// -(1 - 1.6e-5) |b|^2 / 2 in the spare slot and the query 1, so ONE 8-term FFMA chain gives  a.b - |b|^2 / 2 (+ 8e-6 |b|^2)
// and the test  d^2 < worst  becomes a comparison with a per-query constant -- a conservative FILTER (its tolerance,
// 8e-6 (|a|^2 / 2 + |b|^2 + 1), is several times the worst-case rounding of the chain, of the stored norms and of the
// difference form it stands in for: <= ~1.5e-6 |a|^2 + 1.3e-6 |b|^2) in 2 shared-memory loads + D FFMA + a
// compare instead of 2 loads + D subtractions + D FFMA + an add + a compare.  A pair that passes the filter is
// recomputed as sum (a - b)^2 (no |a|^2 + |b|^2 - 2 a.b cancellation in what is compared or reported) and enters the list
// only if it beats the current k-th best, by a branch-free bubble pass over the statically indexed list: the results are
// those of the plain scan, bit for bit.  `causal` restricts query i to references j < i (the incremental roadmap); the
// heaviest query blocks are launched first.
#include "nbx_common.cuh"
#include <float.h>
#include <algorithm>
#include <cstdlib>
#include <type_traits>

namespace {

constexpr int kKnnThreads = 128;
constexpr int kKnnSlices = kKnnThreads / 32;
constexpr int kKnnPend = 8;           // pending candidates per lane between two drains
constexpr int kKnnShare = 4;          // tiles between two agreements on the k-th best distance

__device__ __forceinline__ void knn_cp_async16(void* sdst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void knn_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void knn_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// refs (n, dof) -> packed (n, D): zero padding, and with FILTER the norm term of the filter chain in slot D - 1.  The scan
// kernel then streams tiles with 16-byte cp.async, two tiles in flight (3.2 MB for 10^5 points: it stays in L2).
template <int D, bool FILTER>
__global__ void knn_pack_kernel(const float* __restrict__ refs, int64_t n, int dof, float* __restrict__ packed) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const float* src = refs + r * dof;
    float v[D];
    float n2 = 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        v[d] = d < dof ? src[d] : 0.f;
        n2 = fmaf(v[d], v[d], n2);
    }
    if (FILTER) v[D - 1] = -0.5f * (1.f - 1.6e-5f) * n2;
    float4* dst = reinterpret_cast<float4*>(packed + r * D);
#pragma unroll
    for (int d = 0; d < D / 4; ++d) dst[d] = make_float4(v[4 * d], v[4 * d + 1], v[4 * d + 2], v[4 * d + 3]);
}

template <int D, int K, bool FILTER>
__global__ void __launch_bounds__(kKnnThreads)
knn_kernel(const float* __restrict__ packed, int64_t n_refs, const float* __restrict__ queries, int64_t n_q, int dof, int k,
           int causal, int64_t* __restrict__ idx_out, uint8_t* __restrict__ valid_out, float* __restrict__ d2_out, int share) {
    constexpr int kTileFloats = K <= 16 ? 3072 : 1536;       // two tiles + the lists + the pending lists fit the 48 KB of static shared memory
    constexpr int kKnnTile = kTileFloats / D;                // 384 points at D = 8
    __shared__ __align__(16) float tiles[2][kTileFloats];
    __shared__ float ld[kKnnSlices][K][32];      // the slices' lists, [slice][rank][query]: conflict-free for lane = query
    __shared__ int li[kKnnSlices][K][32];
    __shared__ int pend[kKnnPend][kKnnThreads];  // per lane: tile-local indices of the references that passed the test
    int cnt = 0;
    const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const int64_t nblocks = (n_q + 31) / 32;
    // causal mode is triangular (query i scans i references): heaviest blocks first, the light ones fill the tail
    const int64_t blk = causal ? nblocks - 1 - (int64_t)blockIdx.x : (int64_t)blockIdx.x;
    const int64_t qb = blk * 32;
    const int64_t qi = qb + lane;
    const bool have = qi < n_q;
    float qv[D];
#pragma unroll
    for (int d = 0; d < D; ++d) qv[d] = (have && d < dof) ? queries[qi * dof + d] : 0.f;
    float half_a2 = 0.f;
    if (FILTER) {
#pragma unroll
        for (int d = 0; d < D; ++d) half_a2 = fmaf(qv[d], qv[d], half_a2);
        half_a2 *= 0.5f;
    }
    float fv[D];                               // what the filter chain multiplies: the query with 1 in the spare slot
#pragma unroll
    for (int d = 0; d < D; ++d) fv[d] = qv[d];
    if (FILTER) fv[D - 1] = 1.f;               // dof < D: slot D - 1 of a point is padding
    float bd[K];
    int bi[K];
#pragma unroll
    for (int s = 0; s < K; ++s) { bd[s] = FLT_MAX; bi[s] = -1; }
    float worst = FLT_MAX;           // bound on the k-th best distance of the query (entries beyond k are never reported)
    // pass the filter iff  a.b - |b|^2 / 2 + 8e-6 |b|^2  >  |a|^2 / 2 (1 - 8e-6) - worst / 2 - 8e-6
    auto filter_thr = [&](float w) { return w == FLT_MAX ? -FLT_MAX : fmaf(half_a2, 1.f - 8e-6f, -0.5f * w) - 8e-6f; };
    float thr = -FLT_MAX;
    // references this CTA has to look at: all of them, or those before its last query
    const int64_t r_end = causal ? min(n_refs, qb + 32) : n_refs;
    const int64_t my_end = causal ? min(n_refs, qi) : n_refs;
    // request the tile that starts at reference r0 into buffer b (asynchronous, 16 bytes per copy, coalesced)
    auto request = [&](int64_t r0, int b) {
        const int nt = (int)min((int64_t)kKnnTile, r_end - r0);
        const float4* src = reinterpret_cast<const float4*>(packed + r0 * D);
        float4* dst = reinterpret_cast<float4*>(tiles[b]);
        for (int i = threadIdx.x; i < nt * (D / 4); i += kKnnThreads) knn_cp_async16(dst + i, src + i);
        knn_cp_async_commit();
    };
    if (r_end > 0) request(0, 0);
    int it = 0;
    for (int64_t r0 = 0; r0 < r_end; r0 += kKnnTile, ++it) {
        const int nt = (int)min((int64_t)kKnnTile, r_end - r0);
        const float* tile = tiles[it & 1];
        knn_cp_async_wait<0>();
        __syncthreads();                       // tile `it` has landed for everybody; everybody is done with tile it - 1 and the lists
        if (r0 + kKnnTile < r_end) request(r0 + kKnnTile, (it + 1) & 1);      // lands while this tile is scanned
        // Scan: the warp's quarter of the tile, four references per trip.  A reference that passes the test against the
        // (possibly stale) bound is only NOTED in the lane's pending list; the lists are drained -- exact distance, exact
        // test against the current bound, insertion -- by all lanes together when some lane's list is full and at the end of
        // the tile.  (Inserting on the spot ran ~250 instructions with 2.9 of 32 lanes active every time ANY lane had a
        // candidate: 42 % of the kernel's instructions, ncu source page.)
        const int lim = have ? (int)max((int64_t)0, min((int64_t)nt, my_end - r0)) : 0;
        auto drain = [&]() {
            const int n_max = __reduce_max_sync(0xffffffffu, cnt);
            for (int j = 0; j < n_max; ++j) {
                if (j < cnt) {
                    const int r = pend[j][threadIdx.x];
                    const float4* t4 = reinterpret_cast<const float4*>(tile + r * D);
                    float da = 0.f, db = 0.f;
#pragma unroll
                    for (int d = 0; d < D / 4; ++d) {
                        const float4 c = t4[d];
                        const float e0 = qv[4 * d] - c.x, e1 = qv[4 * d + 1] - c.y, e2 = qv[4 * d + 2] - c.z;
                        const float e3 = (FILTER && d == D / 4 - 1) ? 0.f : qv[4 * d + 3] - c.w;     // the spare slot is not a coordinate
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
                        // k-th entry of the slice's own list.  k <= K is a runtime value and the list must stay statically
                        // indexed (a dynamic bd[k - 1] would move it to local memory): the list ascends, so the k-th entry
                        // is the minimum of the entries from k - 1 on
                        float w = FLT_MAX;
#pragma unroll
                        for (int s = 0; s < K; ++s) w = fminf(w, s >= k - 1 ? bd[s] : FLT_MAX);
                        worst = fminf(worst, w);
                    }
                }
            }
            cnt = 0;
            thr = filter_thr(worst);
        };
        // `check`: some lane's limit is below the tile's end, or the tile ends inside a trip (the diagonal tiles of the causal
        // mode, the last tile, the last, ragged query block); everywhere else the limit test is compiled out of the loop
        auto scan = [&](auto check) {
            constexpr bool CHECK = decltype(check)::value;
            constexpr bool CLAMP = (kKnnTile % (4 * kKnnSlices)) != 0;     // a trip may run past the END of the tile buffer
            int* const pp0 = &pend[0][threadIdx.x];
            int* pp = pp0;                                                  // next free pending slot (cnt is 0 here)
            for (int rb = 0; rb < nt; rb += 4 * kKnnSlices) {               // warp-uniform trip count (the votes need every lane)
                const float* tp = tile + (rb + slice) * D;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int r = rb + u * kKnnSlices + slice;              // beyond nt (CHECK only): stale data, never noted (r < lim <= nt)
                    const float4* t4 = reinterpret_cast<const float4*>(CLAMP ? tile + min(r, kKnnTile - 1) * D : tp + u * kKnnSlices * D);
                    bool pass;
                    if (FILTER) {
                        float sa = 0.f, sb = 0.f;      // two partial sums: half the dependent FFMA chain
#pragma unroll
                        for (int d = 0; d < D / 4; ++d) {
                            const float4 c = t4[d];
                            sa = fmaf(fv[4 * d], c.x, sa); sb = fmaf(fv[4 * d + 1], c.y, sb);
                            sa = fmaf(fv[4 * d + 2], c.z, sa); sb = fmaf(fv[4 * d + 3], c.w, sb);
                        }
                        pass = sa + sb > thr;
                    } else {
                        float da = 0.f, db = 0.f;
#pragma unroll
                        for (int d = 0; d < D / 4; ++d) {
                            const float4 c = t4[d];
                            const float e0 = qv[4 * d] - c.x, e1 = qv[4 * d + 1] - c.y, e2 = qv[4 * d + 2] - c.z, e3 = qv[4 * d + 3] - c.w;
                            da = fmaf(e0, e0, da); db = fmaf(e1, e1, db); da = fmaf(e2, e2, da); db = fmaf(e3, e3, db);
                        }
                        pass = da + db < worst;
                    }
                    if (pass && (!CHECK || r < lim)) { *pp = r; pp += kKnnThreads; }
                }
                if (__any_sync(0xffffffffu, pp > pp0 + (kKnnPend - 4) * kKnnThreads)) {
                    cnt = (int)(pp - pp0) / kKnnThreads;
                    drain();
                    pp = pp0;
                }
            }
            cnt = (int)(pp - pp0) / kKnnThreads;
        };
        if (__all_sync(0xffffffffu, lim == nt) && nt % (4 * kKnnSlices) == 0) scan(std::false_type{}); else scan(std::true_type{});
        if (__any_sync(0xffffffffu, cnt > 0)) drain();
        // every few tiles the slices agree on the k-th best distance of the union of their lists; after the last tile the
        // lists go to shared memory for the final merge
        const bool last = r0 + kKnnTile >= r_end;
        if (!last && (it % share) != share - 1) continue;
#pragma unroll
        for (int s = 0; s < K; ++s) { ld[slice][s][lane] = bd[s]; li[slice][s][lane] = bi[s]; }
        __syncthreads();
        if (!last) {
            int pos[kKnnSlices];
            float head[kKnnSlices];
#pragma unroll
            for (int t = 0; t < kKnnSlices; ++t) { pos[t] = 0; head[t] = ld[t][0][lane]; }
            float kth = FLT_MAX;
            for (int s = 0; s < k; ++s) {
                int m = 0;
#pragma unroll
                for (int t = 1; t < kKnnSlices; ++t) m = head[t] < head[m] ? t : m;
                kth = head[m];
#pragma unroll
                for (int t = 0; t < kKnnSlices; ++t)
                    if (t == m) { ++pos[t]; head[t] = pos[t] < K ? ld[t][pos[t]][lane] : FLT_MAX; }
            }
            worst = fminf(worst, kth);
            thr = filter_thr(worst);
        }
    }
    // the lists of the last tile are in shared memory: slice 0 merges them by (distance, index) and reports the first k
    if (r_end <= 0) {                              // no tile at all (no references): nothing was written to the lists
#pragma unroll
        for (int s = 0; s < K; ++s) { ld[slice][s][lane] = FLT_MAX; li[slice][s][lane] = -1; }
        __syncthreads();
    }
    if (slice == 0 && have) {
        int pos[kKnnSlices];
#pragma unroll
        for (int t = 0; t < kKnnSlices; ++t) pos[t] = 0;
        for (int s = 0; s < k; ++s) {
            int m = -1;
            float md = FLT_MAX;
            int mi = -1;
#pragma unroll
            for (int t = 0; t < kKnnSlices; ++t) {
                if (pos[t] >= K) continue;
                const float d = ld[t][pos[t]][lane];
                const int i = li[t][pos[t]][lane];
                if (i < 0) continue;
                if (m < 0 || d < md || (d == md && i < mi)) { m = t; md = d; mi = i; }
            }
#pragma unroll
            for (int t = 0; t < kKnnSlices; ++t)
                if (t == m) ++pos[t];
            const bool ok = m >= 0;
            idx_out[qi * k + s] = ok ? (int64_t)mi : 0;
            if (valid_out != nullptr) valid_out[qi * k + s] = ok ? 1 : 0;
            if (d2_out != nullptr) d2_out[qi * k + s] = ok ? md : INFINITY;
        }
    }
}

template <int D>
int launch_knn_d(const float* refs, int64_t n_refs, const float* queries, int64_t n_q, int dof, int k, int causal,
                 int64_t* idx, uint8_t* valid, float* d2, int device, cudaStream_t stream) {
    const int64_t grid = (n_q + 31) / 32;
    NBX_REQUIRE(grid <= 0x7fffffffLL, "too many queries for one launch");
    static const bool no_filter = std::getenv("NBX_KNN_NO_FILTER") != nullptr;      // A/B measurements and tests only
    const bool filter = dof < D && !no_filter;
    static const int share_env = std::getenv("NBX_KNN_SHARE") ? std::atoi(std::getenv("NBX_KNN_SHARE")) : 0;   // tuning only
    const int share = share_env > 0 ? share_env : kKnnShare;
    // packed copy of the reference points (stream-ordered workspace, freed after the scan)
    cudaMemPool_t pool = nullptr;
    int rc = nbx_workspace_pool(device, &pool);
    if (rc != NBX_OK) return rc;
    struct Tmp { void* p = nullptr; cudaStream_t st = nullptr; ~Tmp() { if (p) cudaFreeAsync(p, st); } } tmp;
    tmp.st = stream;
    NBX_CUDA_OK(cudaMallocFromPoolAsync(&tmp.p, sizeof(float) * (size_t)D * (size_t)std::max<int64_t>(n_refs, 1), pool, stream));
    float* packed = static_cast<float*>(tmp.p);
    if (n_refs > 0) {
        const unsigned pg = (unsigned)((n_refs + 255) / 256);
        if (filter) knn_pack_kernel<D, true><<<pg, 256, 0, stream>>>(refs, n_refs, dof, packed);
        else knn_pack_kernel<D, false><<<pg, 256, 0, stream>>>(refs, n_refs, dof, packed);
        nbx_count_launch();
    }
#define NBX_KNN_LAUNCH(KK, FF) knn_kernel<D, KK, FF><<<(unsigned)grid, kKnnThreads, 0, stream>>>(packed, n_refs, queries, n_q, dof, k, causal, idx, valid, d2, share)
    if (k <= 16) { if (filter) NBX_KNN_LAUNCH(16, true); else NBX_KNN_LAUNCH(16, false); }
    else         { if (filter) NBX_KNN_LAUNCH(32, true); else NBX_KNN_LAUNCH(32, false); }
#undef NBX_KNN_LAUNCH
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
    if (dof <= 8) return launch_knn_d<8>(refs, n_refs, queries, n_queries, dof, k, causal, idx, valid, dist2, device, st);
    if (dof <= 16) return launch_knn_d<16>(refs, n_refs, queries, n_queries, dof, k, causal, idx, valid, dist2, device, st);
    if (dof <= 32) return launch_knn_d<32>(refs, n_refs, queries, n_queries, dof, k, causal, idx, valid, dist2, device, st);
    return launch_knn_d<64>(refs, n_refs, queries, n_queries, dof, k, causal, idx, valid, dist2, device, st);
}
