// nbx_fkjac_dense.cuh -- the fully unrolled, register-resident FK + Jacobian kernel and the device / host helpers
// it shares with the other FK kernels.  Included by nbx_fkjac.cu (paths of 1..8 joints, float and double) and by
// nbx_fkjac_wide_*.cu (float paths of 9..32 joints, split over two translation units to build in parallel).
//
// Replaces reference numbotics/robots/helpers.py:91-187 (see nbx_fkjac.cu for the design notes).
#pragma once
#include "nbx_common.cuh"
#include "nbx_joint.cuh"
#include <algorithm>
#include <cstdlib>

#ifndef NBX_DENSE_MINB
#define NBX_DENSE_MINB 5   // float kernels: <= 96 registers, 5 CTAs of 128 threads per SM
#endif
#ifndef NBX_DENSE_MINB_F64
#define NBX_DENSE_MINB_F64 3   // double kernels: <= 168 registers, 3 CTAs per SM (what 74 KB of shared memory per CTA allows at 7 joints).
                               // Unbounded they took 188 registers = 2 CTAs = 8 warps per SM: ncu issue 0.31, FP64 pipe 27 %, DRAM 63 % --
                               // latency-bound, not FP64-bound; 2^24 x 7-DoF 1.68 -> 1.37 ms (0.79 -> 0.97 of the measured HBM copy rate)
#endif

namespace nbxd {
using namespace nbxj;

constexpr int kSmallMax = 8;   // dense fully-unrolled kernels cover paths of 1..8 joints

template <typename T>
__device__ __forceinline__ void pose_mul_global(Pose<T>& X, const T* __restrict__ g16) {
    // X <- X * G, G a row-major 4x4 in global memory (per-configuration local_pose)
    T c12[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) c12[k] = g16[k];
    pose_mul_const<T>(X, c12);
}

// ---- TMA bulk copy shared -> global --------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
    uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(ssrc));
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// 16-byte streaming load (q is read exactly once)
__device__ __forceinline__ int4 ld_stream16(const void* p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

template <typename T> struct Vec2;
template <> struct Vec2<float> { using type = float2; };
template <> struct Vec2<double> { using type = double2; };

template <typename T>
__device__ __forceinline__ void store_pose16(T* dst, const Pose<T>& X) {
    // row-major 4x4, vectorised as 4 x 16 bytes (float) / 8 x 16 bytes (double)
    using V2 = typename Vec2<T>::type;
    V2* d = reinterpret_cast<V2*>(dst);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        V2 a, b;
        a.x = X.R[r * 3 + 0]; a.y = X.R[r * 3 + 1];
        b.x = X.R[r * 3 + 2]; b.y = X.p[r];
        d[r * 2 + 0] = a;
        d[r * 2 + 1] = b;
    }
    V2 z0, z1;
    z0.x = T(0); z0.y = T(0); z1.x = T(0); z1.y = T(1);
    d[6] = z0;
    d[7] = z1;
}

// Stage a q tile in shared memory.  Full, 16-byte aligned tiles go through cp.async (LDGSTS: no
// registers, completion tracked by the async group, so the NEXT tile can be in flight while this
// one is being computed); ragged / unaligned tiles use plain loads.
__device__ __forceinline__ void cp_async16(void* sdst, const void* gsrc) {
    const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(sdst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <typename T, int TILE>
__device__ __forceinline__ void load_q_tile(T* sq, const T* __restrict__ q, int64_t tile0, int nvalid, int dof, bool vec_ok) {
    const int n = nvalid * dof;
    const T* src = q + tile0 * dof;
    if (vec_ok && nvalid == TILE) {
        const int n16 = (TILE * dof * (int)sizeof(T)) / 16;   // TILE*dof*sizeof(T) is a multiple of 16 for TILE % 4 == 0
        const int4* s16 = reinterpret_cast<const int4*>(src);
        int4* d16 = reinterpret_cast<int4*>(sq);
        for (int i = threadIdx.x; i < n16; i += TILE) cp_async16(d16 + i, s16 + i);
    } else {
        for (int i = threadIdx.x; i < n; i += TILE) sq[i] = src[i];
    }
    cp_async_commit();
}

// =============================================================================================
// dense kernel: path covers q columns 0..NJ-1 exactly (serial arms: Kinova tool_frame), NJ <= 8.
// Fully unrolled; screws live in registers.  Persistent CTAs: each loops over tiles with the next
// tile's q prefetched (cp.async) while the current one is computed, and the previous tile's output
// still draining through the TMA engine.
// =============================================================================================
template <typename T, int NJ, int TILE, int ALLZ, int KT, int CAP>
__global__ void __launch_bounds__(TILE, (sizeof(T) == 4 && TILE == 128) ? NBX_DENSE_MINB : ((sizeof(T) == 8 && TILE == 128) ? NBX_DENSE_MINB_F64 : 1))   // float: <= 96 registers, 5 CTAs per SM
fkjac_dense_kernel(const __grid_constant__ PathK<T, CAP> P, const T* __restrict__ q, int64_t B,
                   const T* __restrict__ lp, int64_t lps, const T* __restrict__ gp, int64_t gps,
                   T* __restrict__ T_out, T* __restrict__ J_out, int vec_ok, int64_t ntiles) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // layout: [J tile][T tile][q tile x KT]; J and T tiles are 128-byte aligned for the bulk copies
    T* sJ = reinterpret_cast<T*>(smem_raw);
    T* sT = sJ + TILE * 6 * NJ;
    T* sq0 = sT + TILE * 16;
    const int tid = threadIdx.x;
    bool store_in_flight = false;

    // CTA c owns the KT CONSECUTIVE tiles [c*KT, c*KT + KT): the hardware block scheduler then keeps the
    // chip-wide write frontier compact (measured: 7.5 TB/s vs 6.3 TB/s for a grid-stride persistent loop,
    // profiles/membench_r01.txt) while one q-load latency and one CTA launch are shared by KT tiles.
    const int64_t first = (int64_t)blockIdx.x * KT;
#pragma unroll
    for (int k = 0; k < KT; ++k) {
        const int64_t t = first + k;
        if (t < ntiles) load_q_tile<T, TILE>(sq0 + k * (TILE * NJ), q, t * TILE, (int)min((int64_t)TILE, B - t * TILE), NJ, vec_ok != 0);
    }
    cp_async_wait_all();
    __syncthreads();

#pragma unroll 1
    for (int k = 0; k < KT; ++k) {
        const int64_t tile = first + k;
        if (tile >= ntiles) break;
        const int64_t tile0 = tile * TILE;
        const int nvalid = (int)min((int64_t)TILE, B - tile0);
        const T* sq = sq0 + k * (TILE * NJ);

        Pose<T> X;
        T W[NJ][3], Mo[NJ][3];
        T pe[3];
        const bool active = tid < nvalid;
        if (active) {
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                X.R[r * 3 + 0] = P.base[r * 4 + 0]; X.R[r * 3 + 1] = P.base[r * 4 + 1]; X.R[r * 3 + 2] = P.base[r * 4 + 2];
                X.p[r] = P.base[r * 4 + 3];
            }
#pragma unroll
            for (int i = 0; i < NJ; ++i) {
                const T qv = sq[tid * NJ + i];
                T sn, cs;
                sincos_t<T>(qv, &sn, &cs);
                if (ALLZ) {
                    // every joint is a revolute about its own +-z: branch-free straight-line code
                    // (ALLZ == 2: and every offset rotation is a roll about x -- DH-style links)
                    if (ALLZ == 2) joint_step_zx<T, JointK<T>>(X, P.j[i], sn, cs);
                    else joint_step_z<T, JointK<T>>(X, P.j[i], sn, cs);
                    joint_screw_z<T, JointK<T>>(X, P.j[i], W[i], Mo[i]);
                } else {
                    joint_step_sc<T, JointK<T>>(X, P.j[i], qv, sn, cs);
                    joint_screw<T, JointK<T>>(X, P.j[i], W[i], Mo[i]);
                }
            }
            pose_mul_const<T>(X, P.tail);
            const int64_t b = tile0 + tid;
            if (lp != nullptr) pose_mul_global<T>(X, lp + b * lps);
            pe[0] = X.p[0]; pe[1] = X.p[1]; pe[2] = X.p[2];
            if (gp != nullptr) {
                const T* g = gp + b * gps;
                pe[0] = g[3]; pe[1] = g[7]; pe[2] = g[11];
            }
        }

        // the previous tile's bulk stores must have finished READING the staging tiles
        if (store_in_flight && tid == 0) bulk_wait_read_all();
        __syncthreads();

        if (active) {
            if (T_out != nullptr) store_pose16<T>(sT + tid * 16, X);
            if (J_out != nullptr) {
                T row[6 * NJ];
#pragma unroll
                for (int i = 0; i < NJ; ++i) {
                    row[0 * NJ + i] = fma_t<T>(W[i][1], pe[2], fma_t<T>(-W[i][2], pe[1], -Mo[i][0]));
                    row[1 * NJ + i] = fma_t<T>(W[i][2], pe[0], fma_t<T>(-W[i][0], pe[2], -Mo[i][1]));
                    row[2 * NJ + i] = fma_t<T>(W[i][0], pe[1], fma_t<T>(-W[i][1], pe[0], -Mo[i][2]));
                    row[3 * NJ + i] = W[i][0];
                    row[4 * NJ + i] = W[i][1];
                    row[5 * NJ + i] = W[i][2];
                }
                using V2 = typename Vec2<T>::type;
                V2* d = reinterpret_cast<V2*>(sJ + tid * 6 * NJ);
#pragma unroll
                for (int k = 0; k < 3 * NJ; ++k) {
                    V2 v;
                    v.x = row[2 * k]; v.y = row[2 * k + 1];
                    d[k] = v;
                }
            }
        }

        if (nvalid == TILE && vec_ok) {
            // full tile: hand both output tiles to the TMA engine and move on
            fence_proxy_async_smem();
            __syncthreads();
            if (tid == 0) {
                if (T_out != nullptr) bulk_store(T_out + tile0 * 16, sT, TILE * 16 * sizeof(T));
                if (J_out != nullptr) bulk_store(J_out + tile0 * 6 * NJ, sJ, TILE * 6 * NJ * sizeof(T));
                bulk_commit();
            }
            store_in_flight = true;
        } else {
            // ragged last tile (or unaligned user pointers): cooperative coalesced copy-out
            __syncthreads();
            if (T_out != nullptr)
                for (int i = tid; i < nvalid * 16; i += TILE) T_out[tile0 * 16 + i] = sT[i];
            if (J_out != nullptr)
                for (int i = tid; i < nvalid * 6 * NJ; i += TILE) J_out[tile0 * 6 * NJ + i] = sJ[i];
            store_in_flight = false;
        }
    }
    if (store_in_flight && tid == 0) bulk_wait_read_all();   // shared memory must outlive the copies
}

// ---- host side ------------------------------------------------------------------------------
template <typename T, int MAXJ>
int build_path(const nbx_chain* ch, int frame, const double* base_pose16, int use_com, PathK<T, MAXJ>& P) {
    const std::vector<int>& path = ch->frame_path[frame];
    NBX_REQUIRE((int)path.size() <= MAXJ, "path of frame %d has %d movable joints (max %d)", frame, (int)path.size(), MAXJ);
    P.n = (int)path.size();
    P.dof = ch->dof;
    P.use_com = use_com;
    bool dense = (P.n == ch->dof);
    for (int i = 0; i < P.n; ++i) {
        const int j = path[i];
        nbx_fill_joint<T>(P.j[i], &ch->mov_offset[12 * j], &ch->mov_axis[3 * j], ch->mov_type[j], ch->mov_qidx[j]);
        dense = dense && (ch->mov_qidx[j] == i);
    }
    P.dense = dense ? 1 : 0;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c) P.base[r * 4 + c] = (T)base_pose16[r * 4 + c];
    double tail[12];
    for (int k = 0; k < 12; ++k) tail[k] = ch->frame_tail[12 * frame + k];
    if (use_com) pose12_mul(tail, &ch->frame_com[12 * frame], tail);
    for (int k = 0; k < 12; ++k) P.tail[k] = (T)tail[k];
    return NBX_OK;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename T, int NJ, int TILE, int ALLZ, int KT, int CAP>
int launch_dense_kt(const PathK<T, CAP>& P, const T* q, int64_t B, const T* lp, int64_t lps, const T* gp,
                    int64_t gps, T* T_out, T* J_out, int vec_ok, cudaStream_t stream) {
    const size_t smem = sizeof(T) * (size_t)TILE * (6 * NJ + 16 + KT * NJ);
    auto kern = fkjac_dense_kernel<T, NJ, TILE, ALLZ, KT, CAP>;
    NBX_SMEM_LIMIT(kern, smem);
    const int64_t ntiles = (B + TILE - 1) / TILE;
    const int64_t grid = (ntiles + KT - 1) / KT;
    NBX_REQUIRE(grid <= 0x7fffffffLL, "batch too large for one launch");
    kern<<<(unsigned)grid, TILE, smem, stream>>>(P, q, B, lp, lps, gp, gps, T_out, J_out, vec_ok, ntiles);
    nbx_count_launch();
    NBX_CUDA_OK(cudaGetLastError());
    return NBX_OK;
}

}  // namespace nbxd
