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
// Registers stay O(1) in the chain depth: the Jacobian is never stored.  Each iteration walks the path
// twice: walk A accumulates the 21 entries of J J^T column by column (column i needs only p_e of the
// current q, known from the previous walk); walk B re-derives column i, applies dq_i = col_i . x, and
// advances a second pose with the UPDATED q_i, so it ends with FK(q_new) -- which is exactly what the
// next iteration needs.  q lives in a shared-memory tile ([config][dof], the global layout).
// Compute-bound (no HBM traffic inside the loop); fp64 by default because the reference's tol is 1e-6.
#include "nbx_common.cuh"
#include "nbx_joint.cuh"
#include <algorithm>

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
        d = sqrt_t<T>(d > T(0) ? d : T(1e-300));
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

template <typename T>
__global__ void __launch_bounds__(kIkTile)
ik_dls_kernel(const __grid_constant__ PathK<T, NBX_MAX_PATH> P, const __grid_constant__ IkParams<T> K,
              const T* __restrict__ target, int64_t target_stride, const T* __restrict__ q0, int64_t B,
              T* __restrict__ q_out, uint8_t* __restrict__ success, int32_t* __restrict__ iters_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* sq = reinterpret_cast<T*>(smem_raw);          // [kIkTile][dof]
    const int dof = P.dof, n = P.n;
    const int64_t tile0 = (int64_t)blockIdx.x * kIkTile;
    const int nvalid = (int)min((int64_t)kIkTile, B - tile0);
    const int tid = threadIdx.x;
    for (int i = tid; i < nvalid * dof; i += kIkTile) sq[i] = q0[tile0 * dof + i];
    __syncthreads();

    int it = 0;
    bool ok = false;
    if (tid < nvalid) {
        T* myq = sq + tid * dof;
        const T* tg = target + (tile0 + tid) * target_stride;
        T Rt[9], pt[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            Rt[r * 3 + 0] = tg[r * 4 + 0]; Rt[r * 3 + 1] = tg[r * 4 + 1]; Rt[r * 3 + 2] = tg[r * 4 + 2];
            pt[r] = tg[r * 4 + 3];
        }
        auto reset = [&](Pose<T>& X) {
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                X.R[r * 3 + 0] = P.base[r * 4 + 0]; X.R[r * 3 + 1] = P.base[r * 4 + 1]; X.R[r * 3 + 2] = P.base[r * 4 + 2];
                X.p[r] = P.base[r * 4 + 3];
            }
        };
        // FK at q0
        Pose<T> X;
        reset(X);
#pragma unroll 1
        for (int i = 0; i < n; ++i) joint_step<T, JointK<T>>(X, P.j[i], myq[P.j[i].qidx]);
        pose_mul_const<T>(X, P.tail);
        T diff[6];
        T norm = pose_error<T>(X, Rt, pt, diff);
        T pe[3] = {X.p[0], X.p[1], X.p[2]};
        T lambda = T(0.1);
        int fails = 0;
#pragma unroll 1
        for (; it < K.max_iter; ++it) {
            if (!(norm > K.tol) || fails >= K.max_failures) break;
            // ---- walk A: A = J J^T + lambda I (lower triangle) ------------------------------------
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
#pragma unroll
                for (int r = 0; r < 6; ++r)
#pragma unroll
                    for (int c = 0; c <= r; ++c) A[r * (r + 1) / 2 + c] = fma_t<T>(col[r], col[c], A[r * (r + 1) / 2 + c]);
            }
            T x[6];
            chol_solve6<T>(A, diff, x);
            // ---- walk B: q_i += col_i . x, and FK of the updated configuration in the same pass -------
            Pose<T> Xn;
            reset(X);
            reset(Xn);
#pragma unroll 1
            for (int i = 0; i < n; ++i) {
                const int col_q = P.j[i].qidx;
                const T qi = myq[col_q];
                joint_step<T, JointK<T>>(X, P.j[i], qi);
                T col[6];
                joint_column<T>(X, P.j[i], pe, col);
                T dq = T(0);
#pragma unroll
                for (int k = 0; k < 6; ++k) dq = fma_t<T>(col[k], x[k], dq);
                T qn = qi + dq;
                if (K.use_limits) qn = qn < K.lower[i] ? K.lower[i] : (qn > K.upper[i] ? K.upper[i] : qn);
                myq[col_q] = qn;
                joint_step<T, JointK<T>>(Xn, P.j[i], qn);
            }
            pose_mul_const<T>(Xn, P.tail);
            const T new_norm = pose_error<T>(Xn, Rt, pt, diff);
            pe[0] = Xn.p[0]; pe[1] = Xn.p[1]; pe[2] = Xn.p[2];
            const bool worse = new_norm > norm;
            lambda *= worse ? T(1.2) : T(0.5);
            fails = worse ? fails + 1 : 0;
            norm = new_norm;
        }
        ok = norm < K.tol;
    }
    __syncthreads();
    for (int i = tid; i < nvalid * dof; i += kIkTile) q_out[tile0 * dof + i] = sq[i];
    if (tid < nvalid) {
        success[tile0 + tid] = ok ? 1 : 0;
        if (iters_out != nullptr) iters_out[tile0 + tid] = it;
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
    const size_t smem = sizeof(T) * (size_t)kIkTile * std::max(ch->dof, 1);
    NBX_REQUIRE(smem <= 200 * 1024, "dof %d too large for the IK kernel", ch->dof);
    auto kern = ik_dls_kernel<T>;
    NBX_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t grid = (B + kIkTile - 1) / kIkTile;
    NBX_REQUIRE(grid <= 0x7fffffffLL, "batch too large for one launch");
    kern<<<(unsigned)grid, kIkTile, smem, stream>>>(P, K, target, target_stride, q0, B, q_out, success, iters_out);
    nbx_count_launch();
    NBX_CUDA_OK(cudaGetLastError());
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
