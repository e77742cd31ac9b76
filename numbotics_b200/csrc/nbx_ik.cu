// nbx_ik.cu -- batched damped-least-squares inverse kinematics, the whole iteration loop on the device.
//
// Replaces reference numbotics/robots/arm.py:464-552 (Arm.inverse_kinematics): there every iteration is
// a round trip through batched FK, batched Jacobian, np.linalg.solve over the "running" subset and
// boolean-mask bookkeeping.  Here one thread owns one problem and runs the reference's loop to its own
// termination (the reference's masks make the problems independent, so this is the same arithmetic):
//
//     diff = [p* - p ; vee(skew(R* R^T))]                                   arm.py:509-511, spatial.py:207-212
//     while |diff| > tol and failures < max_failures and it < max_iter:        arm.py:516-519
//         q += J^T (J J^T + lambda I)^-1 diff            (6x6 Cholesky)        arm.py:521-528
//         clip to limits (optional)                                            arm.py:529-530
//         diff' = ... at the new q                                             arm.py:532-534
//         worse = |diff'| > |diff| ; lambda *= worse ? 1.2 : 0.5               arm.py:539-542
//         failures = worse ? failures + 1 : 0                                  arm.py:544-545
//
// Lanes are persistent: a finished problem is retired and the next one fetched, so every trip through the loop is one
// uniform iteration for all lanes (see the kernel).  Registers stay O(1) in the chain depth: the Jacobian is never stored.  Each iteration walks the path
// twice: walk A accumulates the 21 entries of J J^T column by column (column i needs only p_e of the
// current q, known from the previous walk); walk B re-derives column i, applies dq_i = col_i . x, and
// advances a second pose with the UPDATED q_i, so it ends with FK(q_new) -- which is exactly what the
// next iteration needs.  q lives in a shared-memory tile ([config][dof], the global layout).
// Compute-bound (no HBM traffic inside the loop); fp64 by default because the reference's tol is 1e-6.
#include "nbx_common.cuh"
#include <limits>
#include "nbx_joint.cuh"
#include <algorithm>
#include <cstdlib>

using namespace nbxj;

namespace {

constexpr int kIkTile = 64;

template <typename T>
struct IkParams {
    T tol;
    int max_iter;
    int max_failures;
    int use_limits;
    T lower[NBX_MAX_PATH];
    T upper[NBX_MAX_PATH];    // indexed by PATH position (not by q column)
};

template <typename T> __device__ __forceinline__ T sqrt_t(T x);
template <> __device__ __forceinline__ float sqrt_t<float>(float x) { return sqrtf(x); }
template <> __device__ __forceinline__ double sqrt_t<double>(double x) { return sqrt(x); }

// error twist between the current pose X and the target (Rt, pt): rows [dp ; rot_diff]
template <typename T>
__device__ __forceinline__ T pose_error(const Pose<T>& X, const T* Rt, const T* pt, T* diff) {
    diff[0] = pt[0] - X.p[0]; diff[1] = pt[1] - X.p[1]; diff[2] = pt[2] - X.p[2];
    // M = Rt * R^T ; rot_diff = 0.5 * [M21 - M12, M02 - M20, M10 - M01]
    T M[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            M[r * 3 + c] = fma_t<T>(Rt[r * 3 + 0], X.R[c * 3 + 0], fma_t<T>(Rt[r * 3 + 1], X.R[c * 3 + 1], Rt[r * 3 + 2] * X.R[c * 3 + 2]));
    diff[3] = T(0.5) * (M[7] - M[5]);
    diff[4] = T(0.5) * (M[2] - M[6]);
    diff[5] = T(0.5) * (M[3] - M[1]);
    T s = T(0);
#pragma unroll
    for (int k = 0; k < 6; ++k) s = fma_t<T>(diff[k], diff[k], s);
    return sqrt_t<T>(s);
}

// Jacobian column of the joint whose post-joint frame is X, for end point pe: [v ; w]
template <typename T>
__device__ __forceinline__ void joint_column(const Pose<T>& X, const JointK<T>& jk, const T* pe, T* col) {
    T w[3], m[3];
    joint_screw<T, JointK<T>>(X, jk, w, m);
    col[0] = fma_t<T>(w[1], pe[2], fma_t<T>(-w[2], pe[1], -m[0]));
    col[1] = fma_t<T>(w[2], pe[0], fma_t<T>(-w[0], pe[2], -m[1]));
    col[2] = fma_t<T>(w[0], pe[1], fma_t<T>(-w[1], pe[0], -m[2]));
    col[3] = w[0]; col[4] = w[1]; col[5] = w[2];
}

// in-place Cholesky solve of the SPD 6x6 system A x = b (A given by its lower triangle, row-major packed)
template <typename T>
__device__ __forceinline__ void chol_solve6(T* L, const T* b, T* x) {
    // L index: row r, col c (c <= r) -> r*(r+1)/2 + c
#define LI(r, c) ((r) * ((r) + 1) / 2 + (c))
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        T d = L[LI(j, j)];
#pragma unroll
        for (int k = 0; k < j; ++k) d = fma_t<T>(-L[LI(j, k)], L[LI(j, k)], d);
        d = sqrt_t<T>(d > T(0) ? d : std::numeric_limits<T>::min());   // type-appropriate pivot floor (1e-300 is 0 in float)
        L[LI(j, j)] = d;
        const T inv = T(1) / d;
#pragma unroll
        for (int i = j + 1; i < 6; ++i) {
            T s = L[LI(i, j)];
#pragma unroll
            for (int k = 0; k < j; ++k) s = fma_t<T>(-L[LI(i, k)], L[LI(j, k)], s);
            L[LI(i, j)] = s * inv;
        }
    }
    T y[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        T s = b[i];
#pragma unroll
        for (int k = 0; k < i; ++k) s = fma_t<T>(-L[LI(i, k)], y[k], s);
        y[i] = s / L[LI(i, i)];
    }
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        T s = y[i];
#pragma unroll
        for (int k = i + 1; k < 6; ++k) s = fma_t<T>(-L[LI(k, i)], x[k], s);
        x[i] = s / L[LI(i, i)];
    }
#undef LI
}

// Persistent lanes.  Problems need 1..max_iter iterations (mean ~10 on the Kinova), so "one thread runs its problem to the
// end" left a warp waiting for its slowest lane (~40 % lane utilisation).  Here a lane that finishes retires its problem and
// fetches the next one from a global counter (one warp-aggregated atomic), and every trip through the loop is ONE uniform
// iteration for all lanes whatever their iteration number.  A freshly loaded problem spends its first trip as a degenerate
// iteration with x = 0: walk B then leaves q untouched and returns FK(q0), the error the reference computes before its
// loop -- so loading costs no divergent code beyond the copies themselves.  Results do not depend on the schedule.
// STORE: walk A parks the Jacobian columns in shared memory ([joint][row][thread]) and walk B reads them back instead of
// walking the OLD configuration a second time (a third of an iteration's instructions, and an iteration is a dependent chain:
// batches of a few 10^5 problems are bound by the latency of the problems that run to max_iter, float and double alike).
// Chains whose columns do not fit (6 n values per thread) recompute them.
template <typename T, bool STORE>
__global__ void __launch_bounds__(kIkTile)
ik_dls_kernel(const __grid_constant__ PathK<T, NBX_MAX_PATH> P, const __grid_constant__ IkParams<T> K,
              const T* __restrict__ target, int64_t target_stride, const T* __restrict__ q0, int64_t B,
              T* __restrict__ q_out, uint8_t* __restrict__ success, int32_t* __restrict__ iters_out,
              unsigned long long* __restrict__ next) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int dof = P.dof, n = P.n;
    const int lane = threadIdx.x & 31;
    T* myq = reinterpret_cast<T*>(smem_raw) + (size_t)threadIdx.x * dof;     // this lane's configuration
    T* mycol = reinterpret_cast<T*>(smem_raw) + (size_t)kIkTile * dof + threadIdx.x;   // STORE: [joint][row][thread]
    auto reset = [&](Pose<T>& X) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            X.R[r * 3 + 0] = P.base[r * 4 + 0]; X.R[r * 3 + 1] = P.base[r * 4 + 1]; X.R[r * 3 + 2] = P.base[r * 4 + 2];
            X.p[r] = P.base[r * 4 + 3];
        }
    };
    int64_t b = -1;                 // current problem (-1: none)
    bool exhausted = false, fresh = false;
    int it = 0, fails = 0;
    T Rt[9], pt[3], diff[6], pe[3] = {T(0), T(0), T(0)};
    T norm = T(0), lambda = T(0.1);
#pragma unroll
    for (int k = 0; k < 9; ++k) Rt[k] = T(0);
#pragma unroll
    for (int k = 0; k < 3; ++k) pt[k] = T(0);
#pragma unroll
    for (int k = 0; k < 6; ++k) diff[k] = T(0);

#pragma unroll 1
    while (true) {
        // ---- retire (reference loop condition, arm.py:516-519) -----------------------------------------
        if (b >= 0 && !fresh && (!(norm > K.tol) || fails >= K.max_failures || it >= K.max_iter)) {
            for (int c = 0; c < dof; ++c) q_out[b * dof + c] = myq[c];
            success[b] = norm < K.tol ? 1 : 0;
            if (iters_out != nullptr) iters_out[b] = it;
            b = -1;
        }
        // ---- refill: one atomic for all the idle lanes of the warp ----------------------------------
        const bool want = b < 0 && !exhausted;
        const unsigned wm = __ballot_sync(0xffffffffu, want);
        if (wm != 0) {
            unsigned long long base = 0;
            const int leader = __ffs(wm) - 1;
            if (lane == leader) base = atomicAdd(next, (unsigned long long)__popc(wm));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (want) {
                const int64_t nb = (int64_t)base + __popc(wm & ((1u << lane) - 1u));
                if (nb >= B) {
                    exhausted = true;
                } else {
                    b = nb;
                    for (int c = 0; c < dof; ++c) myq[c] = q0[b * dof + c];
                    const T* tg = target + b * target_stride;
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        Rt[r * 3 + 0] = tg[r * 4 + 0]; Rt[r * 3 + 1] = tg[r * 4 + 1]; Rt[r * 3 + 2] = tg[r * 4 + 2];
                        pt[r] = tg[r * 4 + 3];
                    }
                    fresh = true; it = 0; fails = 0; lambda = T(0.1); norm = T(0);
                    pe[0] = pe[1] = pe[2] = T(0);
#pragma unroll
                    for (int k = 0; k < 6; ++k) diff[k] = T(0);
                }
            }
        }
        if (__all_sync(0xffffffffu, b < 0)) break;
        if (b < 0) continue;          // idle lane at the tail of the run

        // ---- one iteration --------------------------------------------------------------------------
        // walk A: A = J J^T + lambda I (lower triangle); a fresh problem computes it on p_e = 0 and discards the result
        Pose<T> X;
        T A[21];
#pragma unroll
        for (int k = 0; k < 21; ++k) A[k] = T(0);
#pragma unroll
        for (int k = 0; k < 6; ++k) A[k * (k + 1) / 2 + k] = lambda;
        reset(X);
#pragma unroll 1
        for (int i = 0; i < n; ++i) {
            joint_step<T, JointK<T>>(X, P.j[i], myq[P.j[i].qidx]);
            T col[6];
            joint_column<T>(X, P.j[i], pe, col);
            if (STORE) {
#pragma unroll
                for (int r = 0; r < 6; ++r) mycol[(size_t)(i * 6 + r) * kIkTile] = col[r];
            }
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int c = 0; c <= r; ++c) A[r * (r + 1) / 2 + c] = fma_t<T>(col[r], col[c], A[r * (r + 1) / 2 + c]);
        }
        T x[6];
        chol_solve6<T>(A, diff, x);
        if (fresh) {
#pragma unroll
            for (int k = 0; k < 6; ++k) x[k] = T(0);
        }
        // walk B: q_i += col_i . x, and FK of the updated configuration in the same pass
        Pose<T> Xn;
        reset(X);
        reset(Xn);
#pragma unroll 1
        for (int i = 0; i < n; ++i) {
            const int col_q = P.j[i].qidx;
            const T qi = myq[col_q];
            T col[6];
            if (STORE) {
#pragma unroll
                for (int r = 0; r < 6; ++r) col[r] = mycol[(size_t)(i * 6 + r) * kIkTile];
            } else {
                joint_step<T, JointK<T>>(X, P.j[i], qi);
                joint_column<T>(X, P.j[i], pe, col);
            }
            T dq = T(0);
#pragma unroll
            for (int k = 0; k < 6; ++k) dq = fma_t<T>(col[k], x[k], dq);
            T qn = qi + dq;
            if (K.use_limits && !fresh) qn = qn < K.lower[i] ? K.lower[i] : (qn > K.upper[i] ? K.upper[i] : qn);
            if (fresh) qn = qi;        // (0 * col is 0 for finite columns; be exact regardless)
            myq[col_q] = qn;
            joint_step<T, JointK<T>>(Xn, P.j[i], qn);
        }
        pose_mul_const<T>(Xn, P.tail);
        const T new_norm = pose_error<T>(Xn, Rt, pt, diff);
        pe[0] = Xn.p[0]; pe[1] = Xn.p[1]; pe[2] = Xn.p[2];
        if (fresh) {
            fresh = false;            // this was FK(q0): the error before the reference's loop (arm.py:507-511)
        } else {
            const bool worse = new_norm > norm;
            lambda *= worse ? T(1.2) : T(0.5);
            fails = worse ? fails + 1 : 0;
            ++it;
        }
        norm = new_norm;
    }
}

template <typename T>
int launch_ik(const nbx_chain* ch, int frame, const T* target, int64_t target_stride, const T* q0, int64_t B,
              const double* base_pose, int use_com, int use_limits, const double* lower, const double* upper, double tol,
              int max_iter, int max_failures, T* q_out, uint8_t* success, int32_t* iters_out, cudaStream_t stream) {
    NBX_REQUIRE(ch != nullptr, "chain handle is NULL");
    NBX_REQUIRE(frame >= 0 && frame < ch->n_frames, "frame id %d out of range [0, %d)", frame, ch->n_frames);
    NBX_REQUIRE(B >= 0, "negative batch size");
    NBX_REQUIRE(max_iter >= 1, "max_iter must be greater than 0");
    NBX_REQUIRE(target_stride == 0 || target_stride == 16, "target_stride must be 0 or 16");
    if (B == 0) return NBX_OK;
    NBX_REQUIRE(target != nullptr && q0 != nullptr && q_out != nullptr && success != nullptr && base_pose != nullptr, "NULL buffer");
    NBX_REQUIRE(!use_limits || (lower != nullptr && upper != nullptr), "use_limits needs lower / upper");
    NBX_CUDA_OK(cudaSetDevice(ch->device));
    const std::vector<int>& path = ch->frame_path[frame];
    NBX_REQUIRE((int)path.size() <= NBX_MAX_PATH, "path too long");
    PathK<T, NBX_MAX_PATH> P;
    P.n = (int)path.size();
    P.dof = ch->dof;
    P.use_com = use_com;
    P.dense = 0;
    for (int i = 0; i < P.n; ++i) {
        const int j = path[i];
        nbx_fill_joint<T>(P.j[i], &ch->mov_offset[12 * j], &ch->mov_axis[3 * j], ch->mov_type[j], ch->mov_qidx[j]);
    }
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c) P.base[r * 4 + c] = (T)base_pose[r * 4 + c];
    double tail[12];
    for (int k = 0; k < 12; ++k) tail[k] = ch->frame_tail[12 * frame + k];
    if (use_com) pose12_mul(tail, &ch->frame_com[12 * frame], tail);
    for (int k = 0; k < 12; ++k) P.tail[k] = (T)tail[k];
    IkParams<T> K;
    K.tol = (T)tol; K.max_iter = max_iter; K.max_failures = max_failures; K.use_limits = use_limits;
    for (int i = 0; i < NBX_MAX_PATH; ++i) { K.lower[i] = (T)0; K.upper[i] = (T)0; }
    if (use_limits)
        for (int i = 0; i < P.n; ++i) { K.lower[i] = (T)lower[ch->mov_qidx[path[i]]]; K.upper[i] = (T)upper[ch->mov_qidx[path[i]]]; }
    const size_t smem_q = sizeof(T) * (size_t)kIkTile * std::max(ch->dof, 1);
    const size_t smem_cols = sizeof(T) * (size_t)kIkTile * 6 * std::max(P.n, 1);
    const bool store = smem_cols <= 24 * 1024;          // 7-DoF: 10.5 KB (float) / 21 KB (double) per CTA of 64 threads
    const size_t smem = smem_q + (store ? smem_cols : 0);
    NBX_REQUIRE(smem <= 200 * 1024, "dof %d too large for the IK kernel", ch->dof);
    auto kern = store ? ik_dls_kernel<T, true> : ik_dls_kernel<T, false>;
    NBX_SMEM_LIMIT(kern, smem);
    // persistent lanes: enough CTAs to fill the machine a few times over, never more than the problems need
    static const int cps_env = std::getenv("NBX_IK_CTAS_PER_SM") ? std::atoi(std::getenv("NBX_IK_CTAS_PER_SM")) : 0;   // tuning only
    const int cps = cps_env > 0 ? cps_env : 16;
    const int64_t grid = std::min<int64_t>((B + kIkTile - 1) / kIkTile, (int64_t)ch->sm_count * cps);
    unsigned long long* next = nullptr;
    cudaMemPool_t pool = nullptr;
    {
        const int prc = nbx_workspace_pool(ch->device, &pool);
        if (prc != NBX_OK) return prc;
    }
    NBX_CUDA_OK(cudaMallocFromPoolAsync(reinterpret_cast<void**>(&next), sizeof(unsigned long long), pool, stream));
    NBX_CUDA_OK(cudaMemsetAsync(next, 0, sizeof(unsigned long long), stream));
    kern<<<(unsigned)grid, kIkTile, smem, stream>>>(P, K, target, target_stride, q0, B, q_out, success, iters_out, next);
    nbx_count_launch();
    const cudaError_t launch_err = cudaGetLastError();
    NBX_CUDA_OK(cudaFreeAsync(next, stream));
    NBX_CUDA_OK(launch_err);
    return NBX_OK;
}

}  // namespace

extern "C" {

int nbx_ik_dls_f64(const nbx_chain* chain, int frame, const double* target_pose, int64_t target_stride, const double* q0, int64_t B,
                   const double* base_pose, int use_com, int use_limits, const double* lower, const double* upper, double tol,
                   int max_iter, int max_failures, double* q_out, uint8_t* success, int32_t* iters_out, void* stream) {
    return launch_ik<double>(chain, frame, target_pose, target_stride, q0, B, base_pose, use_com, use_limits, lower, upper, tol,
                             max_iter, max_failures, q_out, success, iters_out, static_cast<cudaStream_t>(stream));
}

int nbx_ik_dls_f32(const nbx_chain* chain, int frame, const float* target_pose, int64_t target_stride, const float* q0, int64_t B,
                   const double* base_pose, int use_com, int use_limits, const double* lower, const double* upper, double tol,
                   int max_iter, int max_failures, float* q_out, uint8_t* success, int32_t* iters_out, void* stream) {
    return launch_ik<float>(chain, frame, target_pose, target_stride, q0, B, base_pose, use_com, use_limits, lower, upper, tol,
                            max_iter, max_failures, q_out, success, iters_out, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
