// nbx_fkjac.cu -- fused forward kinematics + geometric Jacobian for sm_100a.
//
// Replaces reference numbotics/robots/helpers.py:33-187 (nb_joint_transform,
// nb_compute_transformation, nb_compute_jacobian) and the arithmetic around them in
// numbotics/robots/arm.py:394-406, 451-457.
//
// This file: the dispatcher (dense <= 8 joints -> wide 9..32 joints -> deep -> legacy generic), the deep and legacy generic
// kernels, the host-buffer pipeline and the C-ABI entry points.  The dense kernel lives in nbx_fkjac_dense.cuh, the wide
// kernels in nbx_fkjac_wide.inc.
//
// Design of the dense kernel:
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
#include "nbx_fkjac_dense.cuh"
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <type_traits>

using namespace nbxj;
using namespace nbxd;

namespace {

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

// =============================================================================================
// deep kernel: long paths whose joints are all canonical-axis revolutes or prismatic (BASELINE
// config 5, the 30-DoF chain).  The generic kernel above parks one screw (24 bytes) per joint and
// configuration in shared memory until the end point p_e is known; at 30 joints that is ~1 KB per
// configuration and bounds the occupancy at ~6 warps per SM.  This kernel needs no per-joint storage:
//
//   backward pass   e <- X_i(q_i) e  for i = n..1 starting from the trailing offset gives e_0, the
//                   end point expressed in the base frame (vector recursion only: 13 FMA per joint),
//                   hence p_e = base * e_0 BEFORE the forward walk starts;
//   forward walk    carries the rotation R_i and the end point in the CURRENT joint frame,
//                   e_i = X_i(q_i)^-1 e_{i-1}  (rigid inverse, 16 FMA), so the Jacobian column of joint i
//                   is complete the moment the joint is reached:  [R_i (a x e_i) ; R_i a]
//                   (= w x (p_e - p_i) of reference helpers.py:182-183, with p_e - p_i = R_i e_i).
//
// Columns are produced in q order (movable joints on a root->frame path have increasing q indices;
// the host checks it), off-path columns are zeros, and K consecutive columns at a time go through a
// small per-warp staging slab [row][column][configuration] and leave as coalesced runs.  Warps are
// autonomous (their own 32 configurations, their own slab, __syncwarp only).  Shared memory per
// configuration: 8 n bytes of sines / cosines + 24 K bytes of staging (K = 8: ~430 B at 30-DoF).
// =============================================================================================
template <typename T, bool ALLZ>
__global__ void __launch_bounds__(128)
fkjac_deep_kernel(const __grid_constant__ PathK<T, NBX_MAX_PATH> P, const T* __restrict__ q, int64_t B,
                  const T* __restrict__ lp, int64_t lps, const T* __restrict__ gp, int64_t gps,
                  T* __restrict__ T_out, T* __restrict__ J_out, int vec_ok, int K, int slab_elems) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int LS = 33;                         // configuration stride inside the staging slab (odd)
    const int dof = P.dof;
    const int n = P.n;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // per-warp slab: [sin n x 32][cos n x 32][union: q tile 32 x dof | staging 6 x K x 33]
    T* slab = reinterpret_cast<T*>(smem_raw) + (size_t)wid * slab_elems;
    T* ssin = slab;
    T* scos = slab + n * 32;
    T* stg = slab + 2 * n * 32;
    short* colmap = reinterpret_cast<short*>(reinterpret_cast<T*>(smem_raw) + (size_t)4 * slab_elems);   // [dof]

    const int64_t tile0 = ((int64_t)blockIdx.x * 4 + wid) * 32;
    const int nvalid = (int)max((int64_t)0, min((int64_t)32, B - tile0));

    for (int c = threadIdx.x; c < dof; c += 128) colmap[c] = -1;
    // q tile of this warp -> staging area (row-major [configuration][dof], as in global memory)
    if (nvalid > 0) {
        const T* src = q + tile0 * dof;
        if (vec_ok && nvalid == 32) {
            const int n16 = (32 * dof * (int)sizeof(T)) / 16;
            const int4* s16 = reinterpret_cast<const int4*>(src);
            int4* d16 = reinterpret_cast<int4*>(stg);
            for (int i = lane; i < n16; i += 32) cp_async16(d16 + i, s16 + i);
        } else {
            for (int i = lane; i < nvalid * dof; i += 32) stg[i] = src[i];
        }
    }
    cp_async_commit();
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += 128) colmap[P.j[i].qidx] = (short)i;
    cp_async_wait_all();
    __syncthreads();
    if (nvalid == 0) return;
    const bool active = lane < nvalid;
    const bool want_j = J_out != nullptr;

    // ---- phase 0: every sine / cosine (prismatic joints keep q itself in the sine slot) ------------
    {
        const T* myq = stg + lane * dof;
#pragma unroll 4
        for (int i = 0; i < n; ++i) {
            const T qv = active ? myq[P.j[i].qidx] : T(0);
            T sn = qv, cs = T(1);
            if (ALLZ || P.j[i].kind != NBX_KIND_PRISMATIC) {
                sincos_t<T>(qv, &sn, &cs);
                sn *= (T)P.j[i].sgn;               // rotation by sgn * q
            }
            ssin[i * 32 + lane] = sn;
            scos[i * 32 + lane] = cs;
        }
    }
    __syncwarp();                                   // the q tile is dead: its area becomes the staging slab

    const int64_t b = tile0 + (active ? lane : 0);
    // ---- backward pass: e_0 = end point in the base frame ---------------------------------------
    T e[3] = {P.tail[3], P.tail[7], P.tail[11]};
    if (lp != nullptr) {
        const T* g = lp + b * lps;
        const T l0 = g[3], l1 = g[7], l2 = g[11];
#pragma unroll
        for (int r = 0; r < 3; ++r)
            e[r] = fma_t<T>(P.tail[r * 4 + 0], l0, fma_t<T>(P.tail[r * 4 + 1], l1, fma_t<T>(P.tail[r * 4 + 2], l2, e[r])));
    }
#pragma unroll 2
    for (int i = n - 1; i >= 0; --i) {
        const JointK<T>& jk = P.j[i];
        const T s = ssin[i * 32 + lane], c = scos[i * 32 + lane];
        T x[3] = {e[0], e[1], e[2]};
        T tt[3] = {(T)jk.t[0], (T)jk.t[1], (T)jk.t[2]};
        if (ALLZ || jk.kind == NBX_KIND_Z) {
            x[0] = fma_t<T>(c, e[0], -(s * e[1])); x[1] = fma_t<T>(s, e[0], c * e[1]);
        } else if (jk.kind == NBX_KIND_X) {
            x[1] = fma_t<T>(c, e[1], -(s * e[2])); x[2] = fma_t<T>(s, e[1], c * e[2]);
        } else if (jk.kind == NBX_KIND_Y) {
            x[2] = fma_t<T>(c, e[2], -(s * e[0])); x[0] = fma_t<T>(s, e[2], c * e[0]);
        } else {   // prismatic: s holds q
#pragma unroll
            for (int k = 0; k < 3; ++k) tt[k] = fma_t<T>(s, (T)jk.d[k], tt[k]);
        }
#pragma unroll
        for (int r = 0; r < 3; ++r)
            e[r] = fma_t<T>((T)jk.A[r * 3 + 0], x[0], fma_t<T>((T)jk.A[r * 3 + 1], x[1], fma_t<T>((T)jk.A[r * 3 + 2], x[2], tt[r])));
    }
    // true end point (the pose's translation) and the point the Jacobian refers to
    T pe[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
        pe[r] = fma_t<T>(P.base[r * 4 + 0], e[0], fma_t<T>(P.base[r * 4 + 1], e[1], fma_t<T>(P.base[r * 4 + 2], e[2], P.base[r * 4 + 3])));
    if (gp != nullptr) {
        // Jacobian of a given world point: e_0 = base^-1 * g
        const T* g = gp + b * gps;
        const T d0 = g[3] - P.base[3], d1 = g[7] - P.base[7], d2 = g[11] - P.base[11];
#pragma unroll
        for (int k = 0; k < 3; ++k) e[k] = fma_t<T>(P.base[0 * 4 + k], d0, fma_t<T>(P.base[1 * 4 + k], d1, P.base[2 * 4 + k] * d2));
    }

    // ---- forward walk over the q columns -----------------------------------------------------
    T R[9];
#pragma unroll
    for (int r = 0; r < 3; ++r) { R[r * 3 + 0] = P.base[r * 4 + 0]; R[r * 3 + 1] = P.base[r * 4 + 1]; R[r * 3 + 2] = P.base[r * 4 + 2]; }
    int k = 0, c0 = 0;
    const int ncols = want_j ? dof : 0;
    int next = 0;                                   // next path joint (FK-only launches walk the path, not the columns)
    const int steps = want_j ? dof : n;
#pragma unroll 1
    for (int c = 0; c < steps; ++c) {
        const int i = want_j ? (int)colmap[c] : next++;
        T v[3] = {T(0), T(0), T(0)}, w[3] = {T(0), T(0), T(0)};
        if (i >= 0) {
            const JointK<T>& jk = P.j[i];
            const T s = ssin[i * 32 + lane], cs = scos[i * 32 + lane];
            // R <- R * R_off
            T Rn[9];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int cc = 0; cc < 3; ++cc)
                    Rn[r * 3 + cc] = fma_t<T>(R[r * 3 + 0], (T)jk.A[cc], fma_t<T>(R[r * 3 + 1], (T)jk.A[3 + cc], R[r * 3 + 2] * (T)jk.A[6 + cc]));
            // y = R_off^T (e - t [- q d])  (only the Jacobian needs the end point in the joint frame)
            T y[3] = {T(0), T(0), T(0)};
            const bool prism = !ALLZ && jk.kind == NBX_KIND_PRISMATIC;
            if (want_j) {
                T u[3] = {e[0] - (T)jk.t[0], e[1] - (T)jk.t[1], e[2] - (T)jk.t[2]};
                if (prism) {
#pragma unroll
                    for (int kk = 0; kk < 3; ++kk) u[kk] = fma_t<T>(-s, (T)jk.d[kk], u[kk]);
                }
#pragma unroll
                for (int kk = 0; kk < 3; ++kk) y[kk] = fma_t<T>((T)jk.A[0 + kk], u[0], fma_t<T>((T)jk.A[3 + kk], u[1], (T)jk.A[6 + kk] * u[2]));
            }
            // plane rotation of two columns of R and the inverse rotation of y; column of the Jacobian
#define NBX_DEEP_ROT(A_, B_, C_)                                                                   \
            {                                                                                      \
                _Pragma("unroll") for (int r = 0; r < 3; ++r) {                                    \
                    const T a_ = Rn[r * 3 + A_], b_ = Rn[r * 3 + B_];                              \
                    Rn[r * 3 + A_] = fma_t<T>(cs, a_, s * b_);                                     \
                    Rn[r * 3 + B_] = fma_t<T>(cs, b_, -(s * a_));                                  \
                }                                                                                  \
                if (want_j) {                                                                      \
                    const T ya = y[A_], yb = y[B_];                                                \
                    y[A_] = fma_t<T>(cs, ya, s * yb);                                              \
                    y[B_] = fma_t<T>(cs, yb, -(s * ya));                                           \
                    const T sg = (T)jk.sgn;                                                        \
                    const T ea = sg * y[A_], eb = sg * y[B_];                                      \
                    _Pragma("unroll") for (int r = 0; r < 3; ++r) {                                \
                        v[r] = fma_t<T>(ea, Rn[r * 3 + B_], -(eb * Rn[r * 3 + A_]));               \
                        w[r] = sg * Rn[r * 3 + C_];                                                \
                    }                                                                              \
                }                                                                                  \
            }
            if (ALLZ || jk.kind == NBX_KIND_Z) NBX_DEEP_ROT(0, 1, 2)
            else if (jk.kind == NBX_KIND_X) NBX_DEEP_ROT(1, 2, 0)
            else if (jk.kind == NBX_KIND_Y) NBX_DEEP_ROT(2, 0, 1)
            else if (want_j) {
                // prismatic: column = [R a ; 0]  (reference helpers.py:184-185)
#pragma unroll
                for (int r = 0; r < 3; ++r)
                    v[r] = fma_t<T>(Rn[r * 3 + 0], (T)jk.ax[0], fma_t<T>(Rn[r * 3 + 1], (T)jk.ax[1], Rn[r * 3 + 2] * (T)jk.ax[2]));
            }
#undef NBX_DEEP_ROT
#pragma unroll
            for (int kk = 0; kk < 9; ++kk) R[kk] = Rn[kk];
            e[0] = y[0]; e[1] = y[1]; e[2] = y[2];
        }
        if (c < ncols) {
            T* sp = stg + k * LS + lane;
            const int rs = K * LS;
            sp[0 * rs] = v[0]; sp[1 * rs] = v[1]; sp[2 * rs] = v[2];
            sp[3 * rs] = w[0]; sp[4 * rs] = w[1]; sp[5 * rs] = w[2];
            if (++k == K || c == dof - 1) {
                // flush K columns of 32 configurations: lanes over (row, column), runs of 4*k bytes per row
                __syncwarp();
                const int tot = 6 * k;
                T* Jt = J_out + tile0 * 6 * (int64_t)dof + c0;
                for (int idx = lane; idx < tot; idx += 32) {
                    const int r = idx / k, kk = idx - r * k;
                    const T* sr = stg + (r * K + kk) * LS;
                    T* Jc = Jt + r * dof + kk;
                    int cf = 0;
                    for (; cf + 4 <= nvalid; cf += 4) {
                        const T a0 = sr[cf], a1 = sr[cf + 1], a2 = sr[cf + 2], a3 = sr[cf + 3];
                        Jc[(int64_t)(cf + 0) * (6 * dof)] = a0;
                        Jc[(int64_t)(cf + 1) * (6 * dof)] = a1;
                        Jc[(int64_t)(cf + 2) * (6 * dof)] = a2;
                        Jc[(int64_t)(cf + 3) * (6 * dof)] = a3;
                    }
                    for (; cf < nvalid; ++cf) Jc[(int64_t)cf * (6 * dof)] = sr[cf];
                }
                __syncwarp();
                c0 = c + 1;
                k = 0;
            }
        }
    }

    if (T_out != nullptr && active) {
        // pose: R_n * tail.R (* local_pose.R), translation p_e
        Pose<T> X;
#pragma unroll
        for (int kk = 0; kk < 9; ++kk) X.R[kk] = R[kk];
        X.p[0] = X.p[1] = X.p[2] = T(0);
        pose_mul_const<T>(X, P.tail);
        if (lp != nullptr) pose_mul_global<T>(X, lp + b * lps);
        X.p[0] = pe[0]; X.p[1] = pe[1]; X.p[2] = pe[2];
        store_pose16<T>(T_out + b * 16, X);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <typename T, int NJ, int TILE, int ALLZ>
int launch_dense(const nbx_chain* ch, const PathK<T, kSmallMax>& P, const T* q, int64_t B, const T* lp, int64_t lps, const T* gp,
                 int64_t gps, T* T_out, T* J_out, int vec_ok, cudaStream_t stream) {
    (void)ch;
    // consecutive tiles per CTA; small batches keep one tile per CTA so the grid still fills the machine
    static const int kt_env = std::getenv("NBX_DENSE_KT") ? std::atoi(std::getenv("NBX_DENSE_KT")) : 0;   // tuning only
    const int kt = kt_env > 0 ? kt_env : (B >= (int64_t)TILE * 148 * 32 ? 2 : 1);
    if (kt >= 4) return launch_dense_kt<T, NJ, TILE, ALLZ, 4, kSmallMax>(P, q, B, lp, lps, gp, gps, T_out, J_out, vec_ok, stream);
    if (kt >= 2) return launch_dense_kt<T, NJ, TILE, ALLZ, 2, kSmallMax>(P, q, B, lp, lps, gp, gps, T_out, J_out, vec_ok, stream);
    return launch_dense_kt<T, NJ, TILE, ALLZ, 1, kSmallMax>(P, q, B, lp, lps, gp, gps, T_out, J_out, vec_ok, stream);
}

template <typename T, int TILE, bool ALLZ>
int launch_generic_t(const PathK<T, NBX_MAX_PATH>& P, const T* q, int64_t B, const T* lp, int64_t lps, const T* gp, int64_t gps,
                     T* T_out, T* J_out, int vec_ok, size_t smem, cudaStream_t stream) {
    auto kern = fkjac_generic_kernel<T, TILE, ALLZ>;
    NBX_SMEM_LIMIT(kern, smem);
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

// Paths of joints about their own +-z with roll-only offsets (mode 2).  Rz(-q) = Rx(pi) Rz(q) Rx(pi): a joint about its
// own -z is one about +z whose offset roll is turned by pi; the trailing half turn goes into the next joint's offset
// ([Rx(pi) | 0] [Rx(a) | t] = [Rx(a + pi) | (t_x, -t_y, -t_z)]) or, after the last joint, into the tail.  Only signs
// change, so the result is the same to the bit; the kernel then never multiplies by sgn (joint_step_zxp, joint_screw_zp).
template <typename T, int MAXJ>
void fold_axis_signs(PathK<T, MAXJ>& P) {
    bool half_turn = false;                  // a pending Rx(pi) on the left of the next constant pose
    for (int i = 0; i < P.n; ++i) {
        JointK<T>& jk = P.j[i];
        const bool neg = jk.sgn < (T)0;
        if (half_turn) { jk.t[1] = -jk.t[1]; jk.t[2] = -jk.t[2]; }
        if (half_turn != neg) { jk.oc = -jk.oc; jk.os = -jk.os; }
        jk.A[0] = (T)1; jk.A[1] = jk.A[2] = jk.A[3] = jk.A[6] = (T)0;
        jk.A[4] = jk.oc; jk.A[5] = -jk.os; jk.A[7] = jk.os; jk.A[8] = jk.oc;
        jk.ax[0] = jk.ax[1] = (T)0; jk.ax[2] = (T)1;
        jk.sgn = (T)1;
        if (jk.okind == 2 && (jk.oc != (T)1 || jk.os != (T)0)) jk.okind = 1;
        half_turn = neg;
    }
    if (half_turn)
        for (int k = 4; k < 12; ++k) P.tail[k] = -P.tail[k];
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
            bool allz = true, roll = true;
            for (int i = 0; i < P.n; ++i) { allz = allz && P.j[i].kind == NBX_KIND_Z; roll = roll && P.j[i].okind != 0; }
            static const bool no_roll = std::getenv("NBX_NO_ROLL_FASTPATH") != nullptr;   // A/B measurements and tests only
            const int mode = allz ? ((roll && !no_roll) ? 2 : 1) : 0;
            switch (n) {
#define NBX_CASE(N) case N: return mode == 2 ? launch_dense<T, N, TILE, 2>(ch, P, q, B, lp, lps, gp, gps, T_out, J_out, vec_ok, stream) \
                                 : mode == 1 ? launch_dense<T, N, TILE, 1>(ch, P, q, B, lp, lps, gp, gps, T_out, J_out, vec_ok, stream) \
                                             : launch_dense<T, N, TILE, 0>(ch, P, q, B, lp, lps, gp, gps, T_out, J_out, vec_ok, stream);
                NBX_CASE(1) NBX_CASE(2) NBX_CASE(3) NBX_CASE(4) NBX_CASE(5) NBX_CASE(6) NBX_CASE(7) NBX_CASE(8)
#undef NBX_CASE
            }
        }
    }
    if constexpr (std::is_same<T, float>::value) {
        // float paths of 9..32 joints covering q columns 0..n-1: register-resident dense kernels, one warp per tile
        static const bool no_wide = std::getenv("NBX_NO_WIDE") != nullptr;   // A/B measurements and tests only
        if (n > kSmallMax && n <= NBX_WIDE_CAP && n == ch->dof && !no_wide) {
            PathK<float, NBX_WIDE_CAP> P;
            int rc = build_path<float, NBX_WIDE_CAP>(ch, frame, base_pose, use_com, P);
            if (rc != NBX_OK) return rc;
            if (P.dense) {
                bool allz_b = true, roll = true;
                for (int i = 0; i < P.n; ++i) { allz_b = allz_b && P.j[i].kind == NBX_KIND_Z; roll = roll && P.j[i].okind != 0; }
                static const bool no_roll = std::getenv("NBX_NO_ROLL_FASTPATH") != nullptr;   // A/B measurements and tests only
                const int allz = allz_b ? ((roll && !no_roll) ? 2 : 1) : 0;
                if (allz == 2) fold_axis_signs(P);
                return n <= NBX_WIDE_SPLIT ? nbx_launch_wide_a_f32(P, allz, q, B, lp, lps, gp, gps, T_out, J_out, vec_ok, ch->sm_count, stream)
                                           : nbx_launch_wide_b_f32(P, allz, q, B, lp, lps, gp, gps, T_out, J_out, vec_ok, ch->sm_count, stream);
            }
        }
    }
    PathK<T, NBX_MAX_PATH> P;
    int rc = build_path<T, NBX_MAX_PATH>(ch, frame, base_pose, use_com, P);
    if (rc != NBX_OK) return rc;
    // deep kernel: canonical-axis / prismatic joints only, q columns increasing along the path
    static const bool legacy = std::getenv("NBX_GENERIC_LEGACY") != nullptr;   // A/B measurements and tests only
    bool deep_ok = n >= 1 && !legacy, allz = true;
    for (int i = 0; i < n; ++i) {
        deep_ok = deep_ok && P.j[i].kind != NBX_KIND_GENERIC && (i == 0 || P.j[i].qidx > P.j[i - 1].qidx);
        allz = allz && P.j[i].kind == NBX_KIND_Z;
    }
    if (deep_ok) {
        static const int k_env = std::getenv("NBX_DEEP_K") ? std::atoi(std::getenv("NBX_DEEP_K")) : 0;
        const int dof = ch->dof;
        int K = k_env > 0 ? k_env : 8;
        K = std::min(K, dof);
        auto slab = [&](int kk) { return ((2 * n * 32 + std::max(std::max(32 * dof, 6 * kk * 33), 64)) + 31) & ~31; };
        auto bytes = [&](int kk) { return sizeof(T) * 4 * (size_t)slab(kk) + 2 * (size_t)dof + 16; };
        while (K > 1 && bytes(K) > 200 * 1024) K /= 2;
        NBX_REQUIRE(bytes(K) <= 220 * 1024, "dof %d / path %d too large for the shared-memory staged Jacobian", dof, n);
        const size_t smem = bytes(K);
        const int64_t grid = (B + 127) / 128;
        NBX_REQUIRE(grid <= 0x7fffffffLL, "batch too large for one launch");
        auto kern = allz ? fkjac_deep_kernel<T, true> : fkjac_deep_kernel<T, false>;
        NBX_SMEM_LIMIT(kern, smem);
        kern<<<(unsigned)grid, 128, smem, stream>>>(P, q, B, lp, lps, gp, gps, T_out, J_out, vec_ok, K, slab(K));
        nbx_count_launch();
        NBX_CUDA_OK(cudaGetLastError());
        return NBX_OK;
    }
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
// A pipe = kSlots device staging slots, each with its own stream (H2D -> kernel -> D2H in stream order; consecutive
// chunks go to consecutive slots, so chunk k+1's upload and kernel overlap chunk k's download).  Pipes are pooled per
// device: a call borrows one (creating it on first use) and returns it, so concurrent host threads -- the IRIS
// thread-pool caller of numbotics/utils/threading.py:33-37 -- run side by side instead of queueing on one mutex.
struct HostPipe {
    static constexpr int kSlots = 3;
    int device = -1;
    size_t cap[kSlots] = {0, 0, 0};
    void* buf[kSlots] = {nullptr, nullptr, nullptr};
    cudaStream_t st[kSlots] = {nullptr, nullptr, nullptr};
};
std::mutex g_pipe_mu;
std::vector<HostPipe*> g_pipe_free[64];

struct PipeLease {
    HostPipe* hp = nullptr;
    int device;
    explicit PipeLease(int dev) : device(dev) {
        std::lock_guard<std::mutex> lock(g_pipe_mu);
        std::vector<HostPipe*>& f = g_pipe_free[dev];
        if (!f.empty()) { hp = f.back(); f.pop_back(); }
        else hp = new HostPipe();
    }
    ~PipeLease() {
        std::lock_guard<std::mutex> lock(g_pipe_mu);
        g_pipe_free[device].push_back(hp);
    }
};

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
    NBX_REQUIRE(ch->device >= 0 && ch->device < 64, "device index out of range");
    const int dof = ch->dof;
    // chunk so that three slots stay well inside HBM and the copies are long enough to saturate PCIe; small batches
    // are cut into kSlots pieces so that their upload, kernel and download still overlap
    static const int64_t chunk_env = std::getenv("NBX_HOST_CHUNK") ? std::atoll(std::getenv("NBX_HOST_CHUNK")) : 0;   // tuning only
    int64_t chunk = chunk_env > 0 ? chunk_env : (int64_t)1 << 19;
    if (B < chunk * HostPipe::kSlots) chunk = std::max<int64_t>(4096, ((B + HostPipe::kSlots - 1) / HostPipe::kSlots + 127) & ~(int64_t)127);
    const size_t q_bytes = sizeof(T) * (size_t)chunk * dof;
    const size_t t_bytes = T_out ? sizeof(T) * (size_t)chunk * 16 : 0;
    const size_t j_bytes = J_out ? sizeof(T) * (size_t)chunk * 6 * dof : 0;
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t slot_bytes = up(q_bytes) + up(t_bytes) + up(j_bytes);
    PipeLease lease(ch->device);
    HostPipe& hp = *lease.hp;
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
        if (rc != NBX_OK) break;
        if (T_out) NBX_CUDA_OK(cudaMemcpyAsync(T_out + off * 16, dT, sizeof(T) * (size_t)n * 16, cudaMemcpyDeviceToHost, st));
        if (J_out) NBX_CUDA_OK(cudaMemcpyAsync(J_out + off * 6 * dof, dJ, sizeof(T) * (size_t)n * 6 * dof, cudaMemcpyDeviceToHost, st));
    }
    // always drain before the pipe goes back to the pool (also on the error path)
    for (int s = 0; s < HostPipe::kSlots; ++s) {
        const cudaError_t e = cudaStreamSynchronize(hp.st[s]);
        if (e != cudaSuccess && rc == NBX_OK) { nbx_set_error("cudaStreamSynchronize failed: %s", cudaGetErrorString(e)); rc = NBX_ERR_CUDA; }
    }
    return rc;
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
    {
        // roll-only offset rotation?  (exact structure: rpy = (r, 0, 0) gives exact zeros and an exact one)
        const bool roll = R[0] == 1.0 && R[1] == 0.0 && R[2] == 0.0 && R[3] == 0.0 && R[6] == 0.0;
        const bool ident = roll && R[4] == 1.0 && R[8] == 1.0 && R[5] == 0.0 && R[7] == 0.0;
        jk.okind = ident ? 2 : (roll ? 1 : 0);
        jk.oc = roll ? (T)R[4] : (T)1;
        jk.os = roll ? (T)R[7] : (T)0;
    }
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
