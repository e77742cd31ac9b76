// nbx_gjk.cuh -- register-resident GJK distance between rounded convex primitives (fp32).
//
// Replaces, for the verdict `min distance < threshold`, the Bullet narrow phase that the
// reference reaches through pybullet.getClosestPoints (numbotics/physics/chain.py:944-951).
// Geometry model: every primitive is a sharp *core* (point / box /
// z-cylinder) swept by a sphere of radius `margin`;  d(A,B) = d(coreA, coreB) - mA - mB.
//
// Everything is statically indexed so the simplex stays in registers.
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include "nbx.h"

namespace nbxg {

struct V3 {
    float x, y, z;
};
__device__ __forceinline__ V3 mk(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return mk(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
    return mk(fmaf(a.y, b.z, -a.z * b.y), fmaf(a.z, b.x, -a.x * b.z), fmaf(a.x, b.y, -a.y * b.x));
}
__device__ __forceinline__ V3 madd(V3 a, float s, V3 b) { return mk(fmaf(s, b.x, a.x), fmaf(s, b.y, a.y), fmaf(s, b.z, a.z)); }

// World-space core of a primitive.  `ax` holds the box axes as columns of a row-major 3x3
// (axis k = (ax[k], ax[3+k], ax[6+k])); for a cylinder only the z axis is stored in ax[0..2].
struct ShapeW {
    V3 c;
    float ax[9];
    float d[3];   // core dims: box half extents / cylinder {radius, half height} / sphere: unused
    int type;
};

__device__ __forceinline__ V3 support(const ShapeW& S, V3 dir) {
    if (S.type == NBX_PRIM_SPHERE) return S.c;
    if (S.type == NBX_PRIM_BOX) {
        V3 p = S.c;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const V3 a = mk(S.ax[k], S.ax[3 + k], S.ax[6 + k]);
            const float s = dot(dir, a);
            p = madd(p, copysignf(S.d[k], s), a);
        }
        return p;
    }
    // cylinder: segment along u plus a disc of radius r orthogonal to u
    const V3 u = mk(S.ax[0], S.ax[1], S.ax[2]);
    const float du = dot(dir, u);
    V3 perp = madd(dir, -du, u);
    // dir nearly parallel to the axis (cap contacts: the common case) leaves only rounding noise in
    // `perp`, and that noise is NOT orthogonal to u: normalising it would push the support point up
    // to one radius beyond the cap.  Re-orthogonalise once and drop it below fp32 resolution of dir;
    // every cap point is a support point then, so the cap centre is exact to 1e-5 * radius.
    perp = madd(perp, -dot(perp, u), u);
    const float l2 = dot(perp, perp);
    V3 p = madd(S.c, copysignf(S.d[1], du), u);
    if (l2 > 1e-10f * dot(dir, dir)) p = madd(p, S.d[0] * rsqrtf(l2), perp);
    return p;
}

// Closest point to the origin on triangle (a,b,c); returns it and the bitmask (bit0=a, bit1=b, bit2=c)
// of the vertices that support it (Ericson, Real-Time Collision Detection, 5.1.5).
__device__ __forceinline__ V3 closest_tri(V3 a, V3 b, V3 c, int* mask) {
    const V3 ab = b - a, ac = c - a;
    const float d1 = -dot(ab, a), d2 = -dot(ac, a);
    if (d1 <= 0.f && d2 <= 0.f) { *mask = 1; return a; }
    const float d3 = -dot(ab, b), d4 = -dot(ac, b);
    if (d3 >= 0.f && d4 <= d3) { *mask = 2; return b; }
    const float vc = d1 * d4 - d3 * d2;
    if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) {
        *mask = 3;
        return madd(a, d1 / fmaxf(d1 - d3, 1e-37f), ab);      // (max: duplicate vertices give 0/0)
    }
    const float d5 = -dot(ab, c), d6 = -dot(ac, c);
    if (d6 >= 0.f && d5 <= d6) { *mask = 4; return c; }
    const float vb = d5 * d2 - d1 * d6;
    if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) {
        *mask = 5;
        return madd(a, d2 / fmaxf(d2 - d6, 1e-37f), ac);
    }
    const float va = d3 * d6 - d5 * d4;
    if (va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f) {
        *mask = 6;
        return madd(b, (d4 - d3) / fmaxf((d4 - d3) + (d5 - d6), 1e-37f), c - b);
    }
    *mask = 7;
    // interior: barycentric combination.  Unlike the plane projection this is a point OF the triangle
    // whatever the rounding in va/vb/vc did, so |result| stays a valid upper bound even on slivers
    // (collinear support points are the norm on cylinder sides and box faces).
    const float sum = va + vb + vc;
    if (!(sum > 0.f)) { *mask = 1; return a; }
    const float inv = 1.f / sum;
    return madd(madd(a, vb * inv, ab), vc * inv, ac);
}

// Result of one GJK query on the cores.
//   d    : distance estimate.  If `pruned` it is a LOWER bound that already exceeds stop_above;
//          otherwise an UPPER bound (|v| of a point of the Minkowski difference) that is either
//          below stop_below or converged.  Overlapping cores report d = 0 and overlap = true.
struct GjkOut {
    float d;
    float lb;       // best separating-plane LOWER bound seen (0 if none): the distance lies in [lb, d] when !pruned
    bool overlap;   // proven: the simplex encloses the origin
    bool contact;   // |v| fell below the fp32 contact resolution (grazing or shallow penetration)
    bool pruned;
};

template <int MAX_ITER>
__device__ GjkOut gjk_distance(const ShapeW& A, const ShapeW& B, float stop_above, float stop_below,
                                float conv_abs = 1e-6f, float conv_rel = 1e-5f, int max_iter = MAX_ITER) {
    GjkOut out;
    out.lb = 0.f;
    out.overlap = false;
    out.contact = false;
    out.pruned = false;
    V3 v = A.c - B.c;
    float vv = dot(v, v);
    if (vv < 1e-20f) { v = mk(1.f, 0.f, 0.f); vv = 1.f; }
    V3 s0 = mk(0, 0, 0), s1 = s0, s2 = s0;
    int n = 0;
    bool have_ub = false;   // v is a point of the difference only after the first support call
    float ub = FLT_MAX;
#pragma unroll 1
    for (int it = 0; it < max_iter; ++it) {
        const V3 w = support(A, -v) - support(B, v);
        const float vw = dot(v, w);
        // separating-plane lower bound: the difference lies in the half space {x : x.v >= v.w}
        if (vw > 0.f) {
            const float lb = vw * rsqrtf(vv);
            out.lb = fmaxf(out.lb, lb);
            if (lb > stop_above) { out.d = lb; out.pruned = true; return out; }
            if (have_ub && ub - lb <= conv_abs + conv_rel * ub) break;   // converged: the distance lies in [lb, ub]
        }
        if (have_ub && vv - vw <= 1e-7f * vv) break;               // no progress possible
        // closest point of conv(simplex + w) to the origin
        V3 nv;
        if (n == 0) {
            s0 = w; n = 1; nv = w;
        } else if (n == 1) {
            const V3 ab = s0 - w;              // a = w (newest), b = s0
            const float t = -dot(w, ab);
            const float den = dot(ab, ab);
            if (t <= 0.f || den <= 0.f) { s0 = w; n = 1; nv = w; }
            else if (t >= den) { nv = s0; }
            else { nv = madd(w, t / den, ab); s1 = w; n = 2; }
        } else if (n == 2) {
            int m;
            nv = closest_tri(w, s0, s1, &m);
            // compact: bit0 = w, bit1 = s0, bit2 = s1
            V3 q0 = s0, q1 = s1;
            int k = 0;
            if (m & 2) { q0 = s0; k = 1; }
            if (m & 4) { if (k == 0) q0 = s1; else q1 = s1; ++k; }
            if (m & 1) { if (k == 0) q0 = w; else if (k == 1) q1 = w; else s2 = w; ++k; }
            s0 = q0; s1 = q1; n = k;
        } else {
            // tetrahedron (w, s0, s1, s2): test the three faces that contain the new vertex and the old face
            const V3 a = w, b = s0, c = s1, d = s2;
            float best = FLT_MAX;
            int bm = 0;
            V3 bp = mk(0, 0, 0);
            bool outside_any = false;
#define NBX_FACE(P0, P1, P2, OPP, B0, B1, B2)                                         \
            {                                                                          \
                const V3 nrm = cross(P1 - P0, P2 - P0);                                \
                const float sp = -dot(P0, nrm);                                        \
                const float sd = dot(OPP - P0, nrm);                                   \
                /* height of OPP over the face below fp32 resolution of the coordinates: the   */ \
                /* tetrahedron is flat (coplanar support points) and its 'inside' test is noise */ \
                const float flat = 4e-6f * (fabsf(P0.x) + fabsf(P0.y) + fabsf(P0.z) + 1e-3f);  \
                const bool degenerate = sd * sd <= flat * flat * dot(nrm, nrm);        \
                if (degenerate || sp * sd < 0.f) {                                     \
                    outside_any = true;                                                \
                    int m3;                                                            \
                    const V3 cp = closest_tri(P0, P1, P2, &m3);                        \
                    const float dd = dot(cp, cp);                                      \
                    if (dd < best) {                                                   \
                        best = dd; bp = cp;                                            \
                        bm = ((m3 & 1) ? B0 : 0) | ((m3 & 2) ? B1 : 0) | ((m3 & 4) ? B2 : 0); \
                    }                                                                  \
                }                                                                      \
            }
            NBX_FACE(a, b, c, d, 1, 2, 4)
            NBX_FACE(a, c, d, b, 1, 4, 8)
            NBX_FACE(a, d, b, c, 1, 8, 2)
            NBX_FACE(b, d, c, a, 2, 8, 4)
#undef NBX_FACE
            if (!outside_any) { out.d = 0.f; out.overlap = true; return out; }
            nv = bp;
            V3 q0 = s0, q1 = s1, q2 = s2;
            int k = 0;
            if (bm & 2) { q0 = b; k = 1; }
            if (bm & 4) { if (k == 0) q0 = c; else q1 = c; ++k; }
            if (bm & 8) { if (k == 0) q0 = d; else if (k == 1) q1 = d; else q2 = d; ++k; }
            if (bm & 1) { if (k == 0) q0 = a; else if (k == 1) q1 = a; else q2 = a; ++k; }
            s0 = q0; s1 = q1; s2 = q2; n = k;
            if (n == 4) { out.d = 0.f; out.overlap = true; return out; }
        }
        float nvv = dot(nv, nv);
        if (have_ub && nvv >= vv) {
            // The sub-simplex classification lost to rounding (flat / collinear simplices).  Fall back to
            // one Frank-Wolfe step: the old closest point v is itself a point of the difference, so
            // [v, w] is a valid simplex and its closest point is strictly closer whenever v.w < v.v.
            const V3 e = v - w;
            const float den = dot(e, e);
            const float t = -dot(w, e);
            if (!(den > 0.f) || t <= 0.f || t >= den) break;
            nv = madd(w, t / den, e);
            nvv = dot(nv, nv);
            if (nvv >= vv) break;
            s0 = v; s1 = w; n = 2;
        }
        v = nv; vv = nvv; have_ub = true;
        ub = sqrtf(vv);
        // Contact regime: below 50 um the fp32 simplex arithmetic cannot tell grazing separation from shallow
        // penetration (the iteration stalls at |v| ~ 1e-5 instead of enclosing the origin), so it is reported
        // as contact.  Half of the stated 1e-4 m band.
        if (vv <= 2.5e-9f) { out.d = 0.f; out.contact = true; return out; }
        if (ub < stop_below) { out.d = ub; return out; }
    }
    out.d = have_ub ? ub : sqrtf(vv);
    return out;
}

// ---------------------------------------------------------------------------------------------
// GJK with witness points (closest-pair query; reference Arm.closest_to / Proximity, robots/arm.py:599-601,
// physics/chain.py:944-969).  Not a hot path: one call per configuration, on the pair that the distance pass found
// closest.  The simplex carries, next to each vertex w_i = a_i - b_i of the Minkowski difference, the support point
// a_i of A; the closest point v = sum lam_i w_i gives the witness pair  pA = sum lam_i a_i,  pB = pA - v.
// ---------------------------------------------------------------------------------------------
struct GjkWit {
    float d;        // distance between the cores (0 when they overlap)
    V3 pa, pb;      // closest points on core A / core B
    bool overlap;
};

// barycentric coordinates of the point of triangle (a,b,c) closest to the origin (Ericson 5.1.5)
__device__ __forceinline__ void closest_tri_bary(V3 a, V3 b, V3 c, float* la, float* lb, float* lc) {
    const V3 ab = b - a, ac = c - a;
    const float d1 = -dot(ab, a), d2 = -dot(ac, a);
    *la = 1.f; *lb = 0.f; *lc = 0.f;
    if (d1 <= 0.f && d2 <= 0.f) return;
    const float d3 = -dot(ab, b), d4 = -dot(ac, b);
    if (d3 >= 0.f && d4 <= d3) { *la = 0.f; *lb = 1.f; return; }
    const float vc = d1 * d4 - d3 * d2;
    if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) { const float t = d1 / fmaxf(d1 - d3, 1e-37f); *la = 1.f - t; *lb = t; return; }
    const float d5 = -dot(ab, c), d6 = -dot(ac, c);
    if (d6 >= 0.f && d5 <= d6) { *la = 0.f; *lc = 1.f; return; }
    const float vb = d5 * d2 - d1 * d6;
    if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) { const float t = d2 / fmaxf(d2 - d6, 1e-37f); *la = 1.f - t; *lc = t; return; }
    const float va = d3 * d6 - d5 * d4;
    if (va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f) {
        const float t = (d4 - d3) / fmaxf((d4 - d3) + (d5 - d6), 1e-37f);
        *la = 0.f; *lb = 1.f - t; *lc = t;
        return;
    }
    const float sum = va + vb + vc;
    if (!(sum > 0.f)) return;
    const float inv = 1.f / sum;
    *lb = vb * inv; *lc = vc * inv; *la = 1.f - *lb - *lc;
}

template <int MAX_ITER>
__device__ GjkWit gjk_witness(const ShapeW& A, const ShapeW& B) {
    GjkWit out;
    out.overlap = false;
    V3 w[4], a[4];
    float lam[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { w[i] = mk(0, 0, 0); a[i] = mk(0, 0, 0); lam[i] = 0.f; }
    int n = 0;
    V3 v = A.c - B.c;
    float vv = dot(v, v);
    if (vv < 1e-20f) { v = mk(1.f, 0.f, 0.f); vv = 1.f; }
    V3 best_pa = A.c, best_v = v;
    float best_vv = FLT_MAX;
#pragma unroll 1
    for (int it = 0; it < MAX_ITER; ++it) {
        const V3 sa = support(A, -v);
        const V3 sw = sa - support(B, v);
        const float vw = dot(v, sw);
        if (best_vv < FLT_MAX) {
            if (vv - vw <= 1e-7f * vv) break;                                // no support point beyond the current plane
            // bracket [v.w / |v|, |v|] closed: stop BEFORE the simplex degenerates (continuing past convergence is what
            // produces flat tetrahedra and false overlaps)
            const float ubd = sqrtf(vv);
            if (vw > 0.f && ubd - vw / ubd <= 1e-6f + 1e-5f * ubd) break;
        }
        // append the new vertex at index n (n <= 3 here)
#pragma unroll
        for (int i = 0; i < 4; ++i) if (i == n) { w[i] = sw; a[i] = sa; }
        ++n;
        // closest point of the simplex to the origin, as barycentric weights
#pragma unroll
        for (int i = 0; i < 4; ++i) lam[i] = 0.f;
        if (n == 1) {
            lam[0] = 1.f;
        } else if (n == 2) {
            const V3 e = w[1] - w[0];
            const float den = dot(e, e);
            const float t = den > 0.f ? fminf(fmaxf(-dot(w[0], e) / den, 0.f), 1.f) : 0.f;
            lam[0] = 1.f - t; lam[1] = t;
        } else if (n == 3) {
            closest_tri_bary(w[0], w[1], w[2], &lam[0], &lam[1], &lam[2]);
        } else {
            // tetrahedron: the closest of the faces the origin lies outside of; inside all four = overlap
            float bd = FLT_MAX;
            bool outside_any = false;
#define NBX_WFACE(I0, I1, I2, IO)                                                             \
            {                                                                                  \
                const V3 nrm = cross(w[I1] - w[I0], w[I2] - w[I0]);                            \
                const float sp = -dot(w[I0], nrm);                                             \
                const float sd = dot(w[IO] - w[I0], nrm);                                      \
                const float flat = 4e-6f * (fabsf(w[I0].x) + fabsf(w[I0].y) + fabsf(w[I0].z) + 1e-3f); \
                if (sd * sd <= flat * flat * dot(nrm, nrm) || sp * sd < 0.f) {                 \
                    outside_any = true;                                                        \
                    float l0, l1, l2;                                                          \
                    closest_tri_bary(w[I0], w[I1], w[I2], &l0, &l1, &l2);                      \
                    const V3 cp = madd(madd(l0 * w[I0], l1, w[I1]), l2, w[I2]);                \
                    const float dd = dot(cp, cp);                                              \
                    if (dd < bd) { bd = dd; lam[I0] = l0; lam[I1] = l1; lam[I2] = l2; lam[IO] = 0.f; } \
                }                                                                              \
            }
            NBX_WFACE(0, 1, 2, 3)
            NBX_WFACE(0, 1, 3, 2)
            NBX_WFACE(0, 2, 3, 1)
            NBX_WFACE(1, 2, 3, 0)
#undef NBX_WFACE
            if (!outside_any) { out.overlap = true; break; }
        }
        V3 nv = mk(0, 0, 0), pa = mk(0, 0, 0);
#pragma unroll
        for (int i = 0; i < 4; ++i) { nv = madd(nv, lam[i], w[i]); pa = madd(pa, lam[i], a[i]); }
        const float nvv = dot(nv, nv);
        if (nvv >= best_vv) break;                       // rounding stopped the descent: keep the best point so far
        best_vv = nvv; best_v = nv; best_pa = pa;
        v = nv; vv = nvv;
        if (vv <= 2.5e-9f) { out.overlap = true; break; }   // contact regime (see gjk_distance)
        // drop the vertices that do not support the closest point (stable compaction, static indices)
#pragma unroll
        for (int pass = 0; pass < 3; ++pass)
#pragma unroll
            for (int i = 0; i < 3; ++i)
                if (lam[i] <= 0.f && lam[i + 1] > 0.f) {
                    const V3 tw = w[i]; w[i] = w[i + 1]; w[i + 1] = tw;
                    const V3 ta = a[i]; a[i] = a[i + 1]; a[i + 1] = ta;
                    lam[i] = lam[i + 1]; lam[i + 1] = 0.f;
                }
        n = (lam[0] > 0.f) + (lam[1] > 0.f) + (lam[2] > 0.f) + (lam[3] > 0.f);
        if (n == 4) { out.overlap = true; break; }
    }
    if (out.overlap || best_vv == FLT_MAX) {
        out.overlap = true;
        out.d = 0.f;
        out.pa = A.c; out.pb = B.c;
        return out;
    }
    out.d = sqrtf(best_vv);
    out.pa = best_pa;
    out.pb = best_pa - best_v;
    return out;
}

}  // namespace nbxg
