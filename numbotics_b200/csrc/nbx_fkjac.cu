// nbx_fkjac.cu -- fused forward kinematics + geometric Jacobian for sm_100a.
//
// Replaces reference numbotics/robots/helpers.py:33-187 (nb_joint_transform,
// nb_compute_transformation, nb_compute_jacobian) and the arithmetic around them in
// numbotics/robots/arm.py:394-406, 451-457.
//
// Design:
//   * one thread per configuration, one CTA per tile of TILE consecutive configurations;
//   * the path descriptor (pre-folded Rodrigues constants A, Bm, Cm, t per joint) is a
//     __grid_constant__ kernel parameter: it sits in constant bank 0, so every joint constant
//     is a direct FFMA operand -- no loads on the arithmetic path;
//   * q is read with 128-bit coalesced loads into shared memory and re-read per thread;
//   * the transform is accumulated in registers as a 3x4 rigid product; the Jacobian column of
//     joint i needs the end point p_e that is only known after the last joint, so the walk keeps
//     w_i = R_i a_i and the moment m_i = w_i x p_i and finishes with  v_i = w_i x p_e - m_i.
//     No (B, I, 4, 4) `T_mats` intermediate (reference helpers.py:138-158) ever exists;
//   * outputs (B,4,4) and (B,6,dof) are AoS with 64 / 24*dof byte rows: each thread writes its row
//     into a shared-memory tile laid out exactly like global memory and one elected thread hands
//     the whole tile to the TMA engine (cp.async.bulk shared -> global), so HBM sees full-line
//     writes and the SM spends no issue slots on global stores.
//
// HBM-bound: 4*dof + 64 + 24*dof bytes per configuration (fp32), see DESIGN.md.
#include "nbx_common.cuh"
#include "nbx_joint.cuh"
#include <algorithm>
#include <cstdlib>
#include <mutex>

using namespace nbxj;

namespace {

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
template <typename T, int NJ, int TILE, bool ALLZ, int KT>
__global__ void __launch_bounds__(TILE)
fkjac_dense_kernel(const __grid_constant__ PathK<T, kSmallMax> P, const T* __restrict__ q, int64_t B,
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
                    joint_step_z<T, JointK<T>>(X, P.j[i], sn, cs);
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

// =============================================================================================
// generic kernel: any path length <= NBX_MAX_PATH, any dof, sparse columns (BASELINE config 5, 30-DoF).
//
// Registers stay O(1) in the chain depth.  Two phases per tile:
//   A  one thread per configuration walks the path and parks each joint's screw (w_i, m_i) in a
//      shared-memory scratch laid out [component][joint][configuration] with an ODD configuration
//      stride: the phase-A writes (lanes = configurations) and the phase-B reads (lanes = joints)
//      are both bank-conflict free;
//   B  one warp per configuration, lanes over the dof COLUMNS: v = w x p_e - m, and the six rows of
//      that configuration's Jacobian go straight to global memory as six coalesced runs of 4*dof
//      bytes (the 24*dof-byte block of one configuration is contiguous, so L2 sees full lines).
//      Columns of joints that are not on the path are written as zeros in the same pass.
// The 64-byte pose rows still leave through a TMA bulk copy of the whole tile.
// =============================================================================================
template <typename T, int TILE, bool ALLZ>
__global__ void __launch_bounds__(TILE)
fkjac_generic_kernel(const __grid_constant__ PathK<T, NBX_MAX_PATH> P, const T* __restrict__ q, int64_t B,
                     const T* __restrict__ lp, int64_t lps, const T* __restrict__ gp, int64_t gps,
                     T* __restrict__ T_out, T* __restrict__ J_out, int vec_ok) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int SW = TILE + 1;                  // odd stride between components
    constexpr int SJ = 6 * SW + 1;                // odd stride between joints
    const int dof = P.dof;
    const int n = P.n;
    T* sT = reinterpret_cast<T*>(smem_raw);       // [TILE][16]   (first: 128-byte aligned for the bulk copy)
    T* sq = sT + TILE * 16;                       // [TILE][dof]  q, overwritten in place by sin(q)
    T* sc = sq + TILE * dof;                      // [TILE][dof]  cos(q)
    T* spe = sc + TILE * dof;                     // [3][TILE]
    T* sW = spe + 3 * TILE;                       // [n][SJ]: joint-major, then component (stride SW), then configuration
    short* colmap = reinterpret_cast<short*>(sW + (size_t)n * SJ);   // [dof] column -> path index or -1

    const int64_t tile0 = (int64_t)blockIdx.x * TILE;
    const int nvalid = (int)min((int64_t)TILE, B - tile0);
    const int tid = threadIdx.x;

    load_q_tile<T, TILE>(sq, q, tile0, nvalid, dof, vec_ok != 0);
    for (int c = tid; c < dof; c += TILE) colmap[c] = -1;
    cp_async_wait_all();
    __syncthreads();
    for (int i = tid; i < n; i += TILE) colmap[P.j[i].qidx] = (short)i;

    if (tid < nvalid) {
        T* myq = sq + tid * dof;
        T* myc = sc + tid * dof;
        // ---- phase 0: every sine / cosine of this configuration.  No loop-carried dependency, so the
        // transcendental latency overlaps across joints instead of sitting on the serial walk below.
#pragma unroll 4
        for (int i = 0; i < n; ++i) {
            const int col = P.dense ? i : P.j[i].qidx;
            if (ALLZ || P.j[i].kind != NBX_KIND_PRISMATIC) {
                T sn, cs;
                sincos_t<T>(myq[col], &sn, &cs);
                myq[col] = sn;
                myc[col] = cs;
            }
        }
        // ---- phase A: the walk ------------------------------------------------------------------
        Pose<T> X;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            X.R[r * 3 + 0] = P.base[r * 4 + 0]; X.R[r * 3 + 1] = P.base[r * 4 + 1]; X.R[r * 3 + 2] = P.base[r * 4 + 2];
            X.p[r] = P.base[r * 4 + 3];
        }
        T* pw = sW + tid;
        const bool want_j = J_out != nullptr;
#pragma unroll 2
        for (int i = 0; i < n; ++i, pw += SJ) {
            const JointK<T>& jk = P.j[i];
            const int col = P.dense ? i : jk.qidx;
            T w[3], m[3];
            if (ALLZ) {
                joint_step_z<T, JointK<T>>(X, jk, myq[col], myc[col]);
                joint_screw_z<T, JointK<T>>(X, jk, w, m);
            } else {
                // prismatic joints kept q itself in myq[col]
                joint_step_sc<T, JointK<T>>(X, jk, myq[col], myq[col], myc[col]);
                joint_screw<T, JointK<T>>(X, jk, w, m);
            }
            if (want_j) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    pw[k * SW] = m[k];
                    pw[(3 + k) * SW] = w[k];
                }
            }
        }
        pose_mul_const<T>(X, P.tail);
        const int64_t b = tile0 + tid;
        if (lp != nullptr) pose_mul_global<T>(X, lp + b * lps);
        T pe[3] = {X.p[0], X.p[1], X.p[2]};
        if (gp != nullptr) {
            const T* g = gp + b * gps;
            pe[0] = g[3]; pe[1] = g[7]; pe[2] = g[11];
        }
        spe[0 * TILE + tid] = pe[0]; spe[1 * TILE + tid] = pe[1]; spe[2 * TILE + tid] = pe[2];
        if (T_out != nullptr) store_pose16<T>(sT + tid * 16, X);
    }

    const bool bulk = (nvalid == TILE) && vec_ok;
    if (bulk) fence_proxy_async_smem();
    __syncthreads();
    if (T_out != nullptr) {
        if (bulk) {
            if (tid == 0) { bulk_store(T_out + tile0 * 16, sT, TILE * 16 * sizeof(T)); bulk_commit(); }
        } else {
            for (int i = tid; i < nvalid * 16; i += TILE) T_out[tile0 * 16 + i] = sT[i];
        }
    }

    // ---- phase B: one warp per configuration, lanes over the Jacobian's columns -------------------
    if (J_out != nullptr) {
        const int lane = tid & 31, wid = tid >> 5;
        constexpr int NW = TILE / 32;
        for (int col = lane; col < dof; col += 32) {
            const int i = P.dense ? col : (int)colmap[col];
            const T* Wi = sW + (i < 0 ? 0 : i) * SJ;
            T* Jcol = J_out + tile0 * 6 * (int64_t)dof + col;
#pragma unroll 2
            for (int c = wid; c < nvalid; c += NW) {
                T v0 = T(0), v1 = T(0), v2 = T(0), w0 = T(0), w1 = T(0), w2 = T(0);
                if (i >= 0) {
                    const T pe0 = spe[c], pe1 = spe[TILE + c], pe2 = spe[2 * TILE + c];
                    const T m0 = Wi[0 * SW + c], m1 = Wi[1 * SW + c], m2 = Wi[2 * SW + c];
                    w0 = Wi[3 * SW + c]; w1 = Wi[4 * SW + c]; w2 = Wi[5 * SW + c];
                    v0 = fma_t<T>(w1, pe2, fma_t<T>(-w2, pe1, -m0));
                    v1 = fma_t<T>(w2, pe0, fma_t<T>(-w0, pe2, -m1));
                    v2 = fma_t<T>(w0, pe1, fma_t<T>(-w1, pe0, -m2));
                }
                T* Jc = Jcol + (int64_t)c * (6 * dof);
                Jc[0] = v0; Jc[dof] = v1; Jc[2 * dof] = v2;
                Jc[3 * dof] = w0; Jc[4 * dof] = w1; Jc[5 * dof] = w2;
            }
        }
    }
    if (T_out != nullptr && bulk && tid == 0) bulk_wait_read_all();
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
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

template <typename T, int NJ, int TILE, bool ALLZ, int KT>
int launch_dense_kt(const PathK<T, kSmallMax>& P, const T* q, int64_t B, const T* lp, int64_t lps, const T* gp,
                    int64_t gps, T* T_out, T* J_out, int vec_ok, cudaStream_t stream) {
    const size_t smem = sizeof(T) * (size_t)TILE * (6 * NJ + 16 + KT * NJ);
    auto kern = fkjac_dense_kernel<T, NJ, TILE, ALLZ, KT>;
    NBX_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t ntiles = (B + TILE - 1) / TILE;
    const int64_t grid = (ntiles + KT - 1) / KT;
    NBX_REQUIRE(grid <= 0x7fffffffLL, "batch too large for one launch");
    kern<<<(unsigned)grid, TILE, smem, stream>>>(P, q, B, lp, lps, gp, gps, T_out, J_out, vec_ok, ntiles);
    nbx_count_launch();
    NBX_CUDA_OK(cudaGetLastError());
    return NBX_OK;
}

template <typename T, int NJ, int TILE, bool ALLZ>
int launch_dense(const nbx_chain* ch, const PathK<T, kSmallMax>& P, const T* q, int64_t B, const T* lp, int64_t lps, const T* gp,
                 int64_t gps, T* T_out, T* J_out, int vec_ok, cudaStream_t stream) {
    (void)ch;
    // consecutive tiles per CTA; small batches keep one tile per CTA so the grid still fills the machine
    static const int kt_env = std::getenv("NBX_DENSE_KT") ? std::atoi(std::getenv("NBX_DENSE_KT")) : 0;   // tuning only
    const int kt = kt_env > 0 ? kt_env : (B >= (int64_t)TILE * 148 * 32 ? 2 : 1);
    if (kt >= 4) return launch_dense_kt<T, NJ, TILE, ALLZ, 4>(P, q, B, lp, lps, gp, gps, T_out, J_out, vec_ok, stream);
    if (kt >= 2) return launch_dense_kt<T, NJ, TILE, ALLZ, 2>(P, q, B, lp, lps, gp, gps, T_out, J_out, vec_ok, stream);
    return launch_dense_kt<T, NJ, TILE, ALLZ, 1>(P, q, B, lp, lps, gp, gps, T_out, J_out, vec_ok, stream);
}

template <typename T, int TILE, bool ALLZ>
int launch_generic_t(const PathK<T, NBX_MAX_PATH>& P, const T* q, int64_t B, const T* lp, int64_t lps, const T* gp, int64_t gps,
                     T* T_out, T* J_out, int vec_ok, size_t smem, cudaStream_t stream) {
    auto kern = fkjac_generic_kernel<T, TILE, ALLZ>;
    NBX_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t grid = (B + TILE - 1) / TILE;
    NBX_REQUIRE(grid <= 0x7fffffffLL, "batch too large for one launch");
    kern<<<(unsigned)grid, TILE, smem, stream>>>(P, q, B, lp, lps, gp, gps, T_out, J_out, vec_ok);
    nbx_count_launch();
    NBX_CUDA_OK(cudaGetLastError());
    return NBX_OK;
}

template <typename T>
size_t generic_smem(int tile, int dof, int n) {
    return sizeof(T) * ((size_t)tile * 16 + 2 * (size_t)tile * dof + 3 * (size_t)tile + (size_t)n * (6 * (tile + 1) + 1)) + 2 * (size_t)dof + 16;
}

template <typename T, int TILE>
int launch_generic(const PathK<T, NBX_MAX_PATH>& P, const T* q, int64_t B, const T* lp, int64_t lps, const T* gp, int64_t gps,
                   T* T_out, T* J_out, int vec_ok, cudaStream_t stream) {
    const size_t smem = generic_smem<T>(TILE, P.dof, P.n);
    bool allz = P.n > 0;
    for (int i = 0; i < P.n; ++i) allz = allz && P.j[i].kind == NBX_KIND_Z;
    if (allz) return launch_generic_t<T, TILE, true>(P, q, B, lp, lps, gp, gps, T_out, J_out, vec_ok, smem, stream);
    return launch_generic_t<T, TILE, false>(P, q, B, lp, lps, gp, gps, T_out, J_out, vec_ok, smem, stream);
}

template <typename T>
int launch_fkjac(const nbx_chain* ch, int frame, const T* q, int64_t B, const double* base_pose, int use_com,
                 const T* lp, int64_t lps, const T* gp, int64_t gps, T* T_out, T* J_out, cudaStream_t stream) {
    NBX_REQUIRE(ch != nullptr, "chain handle is NULL");
    NBX_REQUIRE(frame >= 0 && frame < ch->n_frames, "frame id %d out of range [0, %d)", frame, ch->n_frames);
    NBX_REQUIRE(B >= 0, "negative batch size");
    NBX_REQUIRE(base_pose != nullptr, "base_pose is NULL");
    NBX_REQUIRE(lp == nullptr || lps == 0 || lps == 16, "local_pose_stride must be 0 or 16");
    NBX_REQUIRE(gp == nullptr || gps == 0 || gps == 16, "global_pose_stride must be 0 or 16");
    NBX_REQUIRE(!(lp != nullptr && gp != nullptr), "local_pose and global_pose cannot both be provided");
    if (B == 0 || (T_out == nullptr && J_out == nullptr)) return NBX_OK;
    NBX_REQUIRE(q != nullptr, "q is NULL");
    NBX_CUDA_OK(cudaSetDevice(ch->device));
    const int vec_ok = aligned16(q) && (T_out == nullptr || aligned16(T_out)) && (J_out == nullptr || aligned16(J_out));
    const int n = (int)ch->frame_path[frame].size();

    constexpr int TILE = 128;
    if (n >= 1 && n <= kSmallMax) {
        PathK<T, kSmallMax> P;
        int rc = build_path<T, kSmallMax>(ch, frame, base_pose, use_com, P);
        if (rc != NBX_OK) return rc;
        if (P.dense) {
            bool allz = true;
            for (int i = 0; i < P.n; ++i) allz = allz && P.j[i].kind == NBX_KIND_Z;
            switch (n) {
#define NBX_CASE(N) case N: return allz ? launch_dense<T, N, TILE, true>(ch, P, q, B, lp, lps, gp, gps, T_out, J_out, vec_ok, stream) \
                                        : launch_dense<T, N, TILE, false>(ch, P, q, B, lp, lps, gp, gps, T_out, J_out, vec_ok, stream);
                NBX_CASE(1) NBX_CASE(2) NBX_CASE(3) NBX_CASE(4) NBX_CASE(5) NBX_CASE(6) NBX_CASE(7) NBX_CASE(8)
#undef NBX_CASE
            }
        }
    }
    PathK<T, NBX_MAX_PATH> P;
    int rc = build_path<T, NBX_MAX_PATH>(ch, frame, base_pose, use_com, P);
    if (rc != NBX_OK) return rc;
    // tile size: the largest tile that still leaves room for >= 3 CTAs per SM (tuning knob NBX_GENERIC_TILE)
    static const int tile_env = std::getenv("NBX_GENERIC_TILE") ? std::atoi(std::getenv("NBX_GENERIC_TILE")) : 0;
    const size_t budget = 72 * 1024;
    int tile = 128;
    if (generic_smem<T>(128, ch->dof, n) > budget) tile = 64;
    if (generic_smem<T>(64, ch->dof, n) > budget) tile = 32;
    if (tile_env == 32 || tile_env == 64 || tile_env == 128) tile = tile_env;
    NBX_REQUIRE(generic_smem<T>(tile, ch->dof, n) <= 220 * 1024, "dof %d / path %d too large for the shared-memory staged Jacobian", ch->dof, n);
    if (tile == 128) return launch_generic<T, 128>(P, q, B, lp, lps, gp, gps, T_out, J_out, vec_ok, stream);
    if (tile == 64) return launch_generic<T, 64>(P, q, B, lp, lps, gp, gps, T_out, J_out, vec_ok, stream);
    return launch_generic<T, 32>(P, q, B, lp, lps, gp, gps, T_out, J_out, vec_ok, stream);
}

// ---- host-buffer pipeline ------------------------------------------------------------------
struct HostPipe {
    static constexpr int kSlots = 3;
    int device = -1;
    size_t cap[kSlots] = {0, 0, 0};
    void* buf[kSlots] = {nullptr, nullptr, nullptr};
    cudaStream_t st[kSlots] = {nullptr, nullptr, nullptr};
    std::mutex mu;
};
HostPipe g_pipe[16];

int pipe_prepare(HostPipe& hp, int device, size_t bytes_per_slot) {
    NBX_CUDA_OK(cudaSetDevice(device));
    for (int s = 0; s < HostPipe::kSlots; ++s) {
        if (hp.st[s] == nullptr) NBX_CUDA_OK(cudaStreamCreateWithFlags(&hp.st[s], cudaStreamNonBlocking));
        if (hp.cap[s] < bytes_per_slot) {
            if (hp.buf[s]) NBX_CUDA_OK(cudaFree(hp.buf[s]));
            hp.buf[s] = nullptr;
            hp.cap[s] = 0;
            NBX_CUDA_OK(cudaMalloc(&hp.buf[s], bytes_per_slot));
            hp.cap[s] = bytes_per_slot;
        }
    }
    hp.device = device;
    return NBX_OK;
}

template <typename T>
int fkjac_host(const nbx_chain* ch, int frame, const T* q, int64_t B, const double* base_pose, int use_com, T* T_out, T* J_out) {
    NBX_REQUIRE(ch != nullptr, "chain handle is NULL");
    NBX_REQUIRE(B >= 0, "negative batch size");
    if (B == 0 || (T_out == nullptr && J_out == nullptr)) return NBX_OK;
    NBX_REQUIRE(ch->device >= 0 && ch->device < 16, "device index out of range");
    const int dof = ch->dof;
    // chunk so that three slots stay well inside HBM and the copies are long enough to saturate PCIe
    const int64_t chunk = 1 << 19;
    const size_t q_bytes = sizeof(T) * (size_t)chunk * dof;
    const size_t t_bytes = T_out ? sizeof(T) * (size_t)chunk * 16 : 0;
    const size_t j_bytes = J_out ? sizeof(T) * (size_t)chunk * 6 * dof : 0;
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t slot_bytes = up(q_bytes) + up(t_bytes) + up(j_bytes);
    HostPipe& hp = g_pipe[ch->device];
    std::lock_guard<std::mutex> lock(hp.mu);
    int rc = pipe_prepare(hp, ch->device, slot_bytes);
    if (rc != NBX_OK) return rc;
    int slot = 0;
    for (int64_t off = 0; off < B; off += chunk, slot = (slot + 1) % HostPipe::kSlots) {
        const int64_t n = std::min(chunk, B - off);
        cudaStream_t st = hp.st[slot];
        unsigned char* base = static_cast<unsigned char*>(hp.buf[slot]);
        T* dq = reinterpret_cast<T*>(base);
        T* dT = T_out ? reinterpret_cast<T*>(base + up(q_bytes)) : nullptr;
        T* dJ = J_out ? reinterpret_cast<T*>(base + up(q_bytes) + up(t_bytes)) : nullptr;
        // stream order guarantees the slot's previous D2H finished before this H2D overwrites it
        NBX_CUDA_OK(cudaMemcpyAsync(dq, q + off * dof, sizeof(T) * (size_t)n * dof, cudaMemcpyHostToDevice, st));
        rc = launch_fkjac<T>(ch, frame, dq, n, base_pose, use_com, nullptr, 0, nullptr, 0, dT, dJ, st);
        if (rc != NBX_OK) return rc;
        if (T_out) NBX_CUDA_OK(cudaMemcpyAsync(T_out + off * 16, dT, sizeof(T) * (size_t)n * 16, cudaMemcpyDeviceToHost, st));
        if (J_out) NBX_CUDA_OK(cudaMemcpyAsync(J_out + off * 6 * dof, dJ, sizeof(T) * (size_t)n * 6 * dof, cudaMemcpyDeviceToHost, st));
    }
    for (int s = 0; s < HostPipe::kSlots; ++s) NBX_CUDA_OK(cudaStreamSynchronize(hp.st[s]));
    return NBX_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
template <typename T>
void nbx_fill_joint(JointK<T>& jk, const double* off12, const double* axis, int type, int qidx) {
    // R, t of the folded offset
    double R[9], t[3];
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) R[r * 3 + c] = off12[r * 4 + c];
        t[r] = off12[r * 4 + 3];
    }
    const double a[3] = {axis[0], axis[1], axis[2]};
    double K[9], Ke[9], Ks[9] = {0, -a[2], a[1], a[2], 0, -a[0], -a[1], a[0], 0};
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            K[r * 3 + c] = a[r] * a[c];
            Ke[r * 3 + c] = K[r * 3 + c] - (r == c ? 1.0 : 0.0);
        }
    auto mm = [&](const double* X, const double* Y, T* out) {
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
                out[r * 3 + c] = (T)(X[r * 3 + 0] * Y[0 * 3 + c] + X[r * 3 + 1] * Y[1 * 3 + c] + X[r * 3 + 2] * Y[2 * 3 + c]);
    };
    if (type == NBX_JOINT_REVOLUTE) {
        mm(R, K, jk.A);
        mm(R, Ke, jk.Bm);
        mm(R, Ks, jk.Cm);
        for (int k = 0; k < 3; ++k) jk.d[k] = (T)0;
    } else {
        for (int k = 0; k < 9; ++k) { jk.A[k] = (T)R[k]; jk.Bm[k] = (T)0; jk.Cm[k] = (T)0; }
        for (int r = 0; r < 3; ++r) jk.d[r] = (T)(R[r * 3 + 0] * a[0] + R[r * 3 + 1] * a[1] + R[r * 3 + 2] * a[2]);
    }
    for (int k = 0; k < 3; ++k) { jk.t[k] = (T)t[k]; jk.ax[k] = (T)a[k]; }
    jk.type = type;
    jk.qidx = qidx;
    jk.kind = type == NBX_JOINT_REVOLUTE ? NBX_KIND_GENERIC : NBX_KIND_PRISMATIC;
    jk.sgn = (T)1;
    static const bool force_generic = std::getenv("NBX_FORCE_GENERIC_AXIS") != nullptr;   // A/B measurements only
    if (type == NBX_JOINT_REVOLUTE && !force_generic) {
        // exactly +-e_k ?  (anything else, including non-unit axes, keeps the literal Rodrigues form)
        for (int k = 0; k < 3; ++k) {
            const int k1 = (k + 1) % 3, k2 = (k + 2) % 3;
            if ((a[k] == 1.0 || a[k] == -1.0) && a[k1] == 0.0 && a[k2] == 0.0) {
                jk.kind = k == 0 ? NBX_KIND_X : (k == 1 ? NBX_KIND_Y : NBX_KIND_Z);
                jk.sgn = (T)a[k];
                for (int i = 0; i < 9; ++i) { jk.A[i] = (T)R[i]; jk.Bm[i] = (T)0; jk.Cm[i] = (T)0; }
            }
        }
    }
}
template void nbx_fill_joint<float>(JointK<float>&, const double*, const double*, int, int);
template void nbx_fill_joint<double>(JointK<double>&, const double*, const double*, int, int);

int nbx_launch_fkjac_f32(const nbx_chain* chain, int frame, const float* q, int64_t B, const double* base_pose, int use_com,
                         const float* lp, int64_t lps, const float* gp, int64_t gps, float* T_out, float* J_out, cudaStream_t stream) {
    return launch_fkjac<float>(chain, frame, q, B, base_pose, use_com, lp, lps, gp, gps, T_out, J_out, stream);
}
int nbx_launch_fkjac_f64(const nbx_chain* chain, int frame, const double* q, int64_t B, const double* base_pose, int use_com,
                         const double* lp, int64_t lps, const double* gp, int64_t gps, double* T_out, double* J_out, cudaStream_t stream) {
    return launch_fkjac<double>(chain, frame, q, B, base_pose, use_com, lp, lps, gp, gps, T_out, J_out, stream);
}

extern "C" {

int nbx_fk_jac_f32(const nbx_chain* chain, int frame, const float* q, int64_t B, const double* base_pose, int use_com,
                   const float* local_pose, int64_t local_pose_stride, const float* global_pose, int64_t global_pose_stride,
                   float* T_out, float* J_out, void* stream) {
    return launch_fkjac<float>(chain, frame, q, B, base_pose, use_com, local_pose, local_pose_stride, global_pose,
                               global_pose_stride, T_out, J_out, static_cast<cudaStream_t>(stream));
}

int nbx_fk_jac_f64(const nbx_chain* chain, int frame, const double* q, int64_t B, const double* base_pose, int use_com,
                   const double* local_pose, int64_t local_pose_stride, const double* global_pose, int64_t global_pose_stride,
                   double* T_out, double* J_out, void* stream) {
    return launch_fkjac<double>(chain, frame, q, B, base_pose, use_com, local_pose, local_pose_stride, global_pose,
                                global_pose_stride, T_out, J_out, static_cast<cudaStream_t>(stream));
}

int nbx_fk_jac_host_f32(const nbx_chain* chain, int frame, const float* q, int64_t B, const double* base_pose, int use_com,
                        float* T_out, float* J_out) {
    return fkjac_host<float>(chain, frame, q, B, base_pose, use_com, T_out, J_out);
}

int nbx_fk_jac_host_f64(const nbx_chain* chain, int frame, const double* q, int64_t B, const double* base_pose, int use_com,
                        double* T_out, double* J_out) {
    return fkjac_host<double>(chain, frame, q, B, base_pose, use_com, T_out, J_out);
}

}  // extern "C"
