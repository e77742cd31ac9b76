// nbx_joint.cuh -- one joint of the kinematic walk, shared by the FK/Jacobian and collision kernels.
//
// Restates reference numbotics/robots/helpers.py:33-88 (nb_joint_transform) fused with the chain product
// of :105-108:  pose <- pose * X(q).
#pragma once
#include <cuda_runtime.h>
#include "nbx_common.cuh"

namespace nbxj {

template <typename T> __device__ __forceinline__ T fma_t(T a, T b, T c);
template <> __device__ __forceinline__ float fma_t<float>(float a, float b, float c) { return fmaf(a, b, c); }
template <> __device__ __forceinline__ double fma_t<double>(double a, double b, double c) { return fma(a, b, c); }

// sin and cos with one shared Cody-Waite reduction (three-constant pi/2 split) and the classic
// degree-7 / degree-8 minimax polynomials on [-pi/4, pi/4]: <= 1.5 ulp for |x| < 1e5, about half the
// instructions of sincosf() (no Payne-Hanek path, no local stack).  Larger arguments take sincosf().
// (returned by value: results handed back through pointers would live in local memory and put an STL/LDL pair
// on every call site's fast path as well)
static __device__ __noinline__ float2 sincos_huge(float x) {
    float s, c;
    sincosf(x, &s, &c);
    return make_float2(s, c);
}

// the fast path alone: valid for |x| <= 1e5 (callers that screen their arguments once per tile use it directly:
// no compare / branch / reconvergence per joint)
__device__ __forceinline__ void sincos_core(float x, float* sp, float* cp) {
    // round-to-nearest via the 1.5 * 2^23 magic constant: FFMA / FADD only (no FRND / F2I on the XU pipe),
    // the integer part sits in the low mantissa bits of `t`
    const float t = fmaf(x, 0.636619747f, 12582912.0f);
    const int i = __float_as_int(t);
    const float j = t - 12582912.0f;
    float r = fmaf(j, -1.57079601e+00f, x);
    r = fmaf(j, -3.13916473e-07f, r);
    r = fmaf(j, -5.39030253e-15f, r);
    const float r2 = r * r;
    float s = fmaf(r2, -1.95152959e-04f, 8.33216087e-03f);
    s = fmaf(s, r2, -1.66666546e-01f);
    s = fmaf(s * r2, r, r);
    float c = fmaf(r2, 2.44331571e-05f, -1.38873163e-03f);
    c = fmaf(c, r2, 4.16666418e-02f);
    c = fmaf(c, r2, -0.5f);
    c = fmaf(c, r2, 1.0f);
    const bool swap = (i & 1) != 0;
    float so = swap ? c : s;
    float co = swap ? s : c;
    so = __int_as_float(__float_as_int(so) ^ ((i & 2) << 30));          // quadrants 2, 3: sin < 0
    co = __int_as_float(__float_as_int(co) ^ (((i + 1) & 2) << 30));    // quadrants 1, 2: cos < 0
    *sp = so;
    *cp = co;
}

__device__ __forceinline__ void sincos_fast(float x, float* sp, float* cp) {
    if (fabsf(x) > 1.0e5f) {   // cold, kept out of line
        const float2 r = sincos_huge(x);
        *sp = r.x;
        *cp = r.y;
        return;
    }
    sincos_core(x, sp, cp);
}

// Cold companion of sincos_core: replace every |x| > 1e5 of a row by the equivalent angle in [-pi, pi]
// (double-precision sine / cosine of the exact float, then atan2: half an ulp of the reduced angle).
static __device__ __noinline__ void reduce_huge_angles(float* row, int n) {
    for (int i = 0; i < n; ++i) {
        const float x = row[i];
        if (fabsf(x) > 1.0e5f) {
            double s, c;
            sincos((double)x, &s, &c);
            row[i] = (float)atan2(s, c);      // NaN / inf stay NaN
        }
    }
}

template <typename T> __device__ __forceinline__ void sincos_t(T x, T* s, T* c);
template <> __device__ __forceinline__ void sincos_t<float>(float x, float* s, float* c) { sincos_fast(x, s, c); }
template <> __device__ __forceinline__ void sincos_t<double>(double x, double* s, double* c) { sincos(x, s, c); }

// R (3x3 row-major) and p: world pose of the current joint frame.
template <typename T>
struct Pose {
    T R[9];
    T p[3];
};

template <typename T>
__device__ __forceinline__ void pose_mul_const(Pose<T>& X, const T* c12) {
    // X <- X * [C | t] with a constant pose12
    T Rn[9], pn[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        pn[r] = fma_t<T>(X.R[r * 3 + 0], c12[3], fma_t<T>(X.R[r * 3 + 1], c12[7], fma_t<T>(X.R[r * 3 + 2], c12[11], X.p[r])));
#pragma unroll
        for (int c = 0; c < 3; ++c)
            Rn[r * 3 + c] = fma_t<T>(X.R[r * 3 + 0], c12[c], fma_t<T>(X.R[r * 3 + 1], c12[4 + c], X.R[r * 3 + 2] * c12[8 + c]));
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) X.R[k] = Rn[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) X.p[k] = pn[k];
}

// rotate columns (A, B) of X.R by the plane rotation (c, s):  colA' = c colA + s colB ; colB' = c colB - s colA
template <typename T, int A, int B>
__device__ __forceinline__ void rot_cols(Pose<T>& X, T c, T s) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const T a = X.R[r * 3 + A], b = X.R[r * 3 + B];
        X.R[r * 3 + A] = fma_t<T>(c, a, s * b);
        X.R[r * 3 + B] = fma_t<T>(c, b, -(s * a));
    }
}

// pose <- pose * X(q).  JK is JointK<T> or TreeJointK (same field names).
template <typename T, typename JK>
__device__ __forceinline__ void joint_step_sc(Pose<T>& X, const JK& jk, T qv, T s, T c) {
    if (jk.kind == NBX_KIND_GENERIC) {
        T M[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) M[k] = fma_t<T>(s, (T)jk.Cm[k], fma_t<T>(-c, (T)jk.Bm[k], (T)jk.A[k]));
        T Rn[9];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            X.p[r] = fma_t<T>(X.R[r * 3 + 0], (T)jk.t[0], fma_t<T>(X.R[r * 3 + 1], (T)jk.t[1], fma_t<T>(X.R[r * 3 + 2], (T)jk.t[2], X.p[r])));
#pragma unroll
            for (int cc = 0; cc < 3; ++cc)
                Rn[r * 3 + cc] = fma_t<T>(X.R[r * 3 + 0], M[cc], fma_t<T>(X.R[r * 3 + 1], M[3 + cc], X.R[r * 3 + 2] * M[6 + cc]));
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) X.R[k] = Rn[k];
        return;
    }
    // canonical axis or prismatic: first the constant part  X <- X * [R_off | t (+ q d)]
    T tt[3] = {(T)jk.t[0], (T)jk.t[1], (T)jk.t[2]};
    if (jk.kind == NBX_KIND_PRISMATIC) {
#pragma unroll
        for (int k = 0; k < 3; ++k) tt[k] = fma_t<T>(qv, (T)jk.d[k], tt[k]);
    }
    T Rn[9];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        X.p[r] = fma_t<T>(X.R[r * 3 + 0], tt[0], fma_t<T>(X.R[r * 3 + 1], tt[1], fma_t<T>(X.R[r * 3 + 2], tt[2], X.p[r])));
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
            Rn[r * 3 + cc] = fma_t<T>(X.R[r * 3 + 0], (T)jk.A[cc], fma_t<T>(X.R[r * 3 + 1], (T)jk.A[3 + cc], X.R[r * 3 + 2] * (T)jk.A[6 + cc]));
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) X.R[k] = Rn[k];
    if (jk.kind == NBX_KIND_PRISMATIC) return;
    s *= (T)jk.sgn;                      // rotation by sgn * q
    if (jk.kind == NBX_KIND_Z) rot_cols<T, 0, 1>(X, c, s);
    else if (jk.kind == NBX_KIND_X) rot_cols<T, 1, 2>(X, c, s);
    else rot_cols<T, 2, 0>(X, c, s);
}

template <typename T, typename JK>
__device__ __forceinline__ void joint_step(Pose<T>& X, const JK& jk, T qv) {
    T s = T(0), c = T(1);
    if (jk.kind != NBX_KIND_PRISMATIC) sincos_t<T>(qv, &s, &c);
    joint_step_sc<T, JK>(X, jk, qv, s, c);
}

// Branch-free special case: revolute about the joint's own +-z (kind == NBX_KIND_Z).
template <typename T, typename JK>
__device__ __forceinline__ void joint_step_z(Pose<T>& X, const JK& jk, T s, T c) {
    T Rn[9];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        X.p[r] = fma_t<T>(X.R[r * 3 + 0], (T)jk.t[0], fma_t<T>(X.R[r * 3 + 1], (T)jk.t[1], fma_t<T>(X.R[r * 3 + 2], (T)jk.t[2], X.p[r])));
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
            Rn[r * 3 + cc] = fma_t<T>(X.R[r * 3 + 0], (T)jk.A[cc], fma_t<T>(X.R[r * 3 + 1], (T)jk.A[3 + cc], X.R[r * 3 + 2] * (T)jk.A[6 + cc]));
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) X.R[k] = Rn[k];
    s *= (T)jk.sgn;
    rot_cols<T, 0, 1>(X, c, s);
}

// Same joint when the offset rotation is a roll about x (okind 1; identity = roll by 0):
// R <- R * Rx(alpha) is a plane rotation of columns (1, 2): 33 instead of 48 instructions per joint.
template <typename T, typename JK>
__device__ __forceinline__ void joint_step_zx(Pose<T>& X, const JK& jk, T s, T c) {
#pragma unroll
    for (int r = 0; r < 3; ++r)
        X.p[r] = fma_t<T>(X.R[r * 3 + 0], (T)jk.t[0], fma_t<T>(X.R[r * 3 + 1], (T)jk.t[1], fma_t<T>(X.R[r * 3 + 2], (T)jk.t[2], X.p[r])));
    rot_cols<T, 1, 2>(X, (T)jk.oc, (T)jk.os);
    s *= (T)jk.sgn;
    rot_cols<T, 0, 1>(X, c, s);
}

// joint_step_zx / joint_screw_z for paths whose axis signs were folded into the offsets on the host
// (fold_axis_signs, nbx_fkjac.cu): every joint turns about its own +z, no multiplication by sgn.
template <typename T, typename JK>
__device__ __forceinline__ void joint_step_zxp(Pose<T>& X, const JK& jk, T s, T c) {
#pragma unroll
    for (int r = 0; r < 3; ++r)
        X.p[r] = fma_t<T>(X.R[r * 3 + 0], (T)jk.t[0], fma_t<T>(X.R[r * 3 + 1], (T)jk.t[1], fma_t<T>(X.R[r * 3 + 2], (T)jk.t[2], X.p[r])));
    rot_cols<T, 1, 2>(X, (T)jk.oc, (T)jk.os);
    rot_cols<T, 0, 1>(X, c, s);
}

template <typename T>
__device__ __forceinline__ void joint_screw_zp(const Pose<T>& X, T* w, T* m) {
    w[0] = X.R[2]; w[1] = X.R[5]; w[2] = X.R[8];
    m[0] = fma_t<T>(w[1], X.p[2], -w[2] * X.p[1]);
    m[1] = fma_t<T>(w[2], X.p[0], -w[0] * X.p[2]);
    m[2] = fma_t<T>(w[0], X.p[1], -w[1] * X.p[0]);
}

template <typename T, typename JK>
__device__ __forceinline__ void joint_screw_z(const Pose<T>& X, const JK& jk, T* w, T* m) {
    w[0] = (T)jk.sgn * X.R[2]; w[1] = (T)jk.sgn * X.R[5]; w[2] = (T)jk.sgn * X.R[8];
    m[0] = fma_t<T>(w[1], X.p[2], -w[2] * X.p[1]);
    m[1] = fma_t<T>(w[2], X.p[0], -w[0] * X.p[2]);
    m[2] = fma_t<T>(w[0], X.p[1], -w[1] * X.p[0]);
}

// world axis w = R a and moment m = w x p of the joint frame just reached.  Prismatic joints are
// stored as (w_store, m_store) = (0, -w) so the common finish  v = w_store x p_e - m_store,
// omega = w_store  yields [w; 0] without a branch (reference helpers.py:181-185).
template <typename T, typename JK>
__device__ __forceinline__ void joint_screw(const Pose<T>& X, const JK& jk, T* w, T* m) {
    T a[3];
    if (jk.kind == NBX_KIND_GENERIC || jk.kind == NBX_KIND_PRISMATIC) {
#pragma unroll
        for (int r = 0; r < 3; ++r)
            a[r] = fma_t<T>(X.R[r * 3 + 0], (T)jk.ax[0], fma_t<T>(X.R[r * 3 + 1], (T)jk.ax[1], X.R[r * 3 + 2] * (T)jk.ax[2]));
    } else {
        const int k = jk.kind == NBX_KIND_X ? 0 : (jk.kind == NBX_KIND_Y ? 1 : 2);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const T col = k == 0 ? X.R[r * 3 + 0] : (k == 1 ? X.R[r * 3 + 1] : X.R[r * 3 + 2]);
            a[r] = (T)jk.sgn * col;
        }
    }
    if (jk.type == NBX_JOINT_REVOLUTE) {
        w[0] = a[0]; w[1] = a[1]; w[2] = a[2];
        m[0] = fma_t<T>(a[1], X.p[2], -a[2] * X.p[1]);
        m[1] = fma_t<T>(a[2], X.p[0], -a[0] * X.p[2]);
        m[2] = fma_t<T>(a[0], X.p[1], -a[1] * X.p[0]);
    } else {
        w[0] = w[1] = w[2] = T(0);
        m[0] = -a[0]; m[1] = -a[1]; m[2] = -a[2];
    }
}

}  // namespace nbxj
