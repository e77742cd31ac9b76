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

// out-of-line copy for instantiations that care more about their instruction footprint than about a call (the fused
// distance kernel: five inlined copies of closest_tri are ~750 instructions of a GJK loop that does not fit the L0 I-cache)
__device__ __noinline__ V3 closest_tri_call(V3 a, V3 b, V3 c, int* mask) { return closest_tri(a, b, c, mask); }

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

// host builds of this header (tests/host_harness, scripts/) may define NBX_GJK_COUNT_ITER(it) to record iteration counts
#ifndef NBX_GJK_COUNT_ITER
#define NBX_GJK_COUNT_ITER(it)
#endif
#ifndef NBX_EPA_COUNT_EXPANSION
#define NBX_EPA_COUNT_EXPANSION(it)
#endif
// the EPA stops when the bracket [closest face, smallest support value seen] is this narrow
#ifndef NBX_EPA_TOL_ABS
#define NBX_EPA_TOL_ABS 1e-5f
#define NBX_EPA_TOL_REL 5e-5f
#endif

template <int MAX_ITER, bool COMPACT = false>
__device__ GjkOut gjk_distance(const ShapeW& A, const ShapeW& B, float stop_above, float stop_below,
                                float conv_abs = 1e-6f, float conv_rel = 1e-5f, int max_iter = MAX_ITER) {
    auto ctri = [](V3 a, V3 b, V3 c, int* m) {
        if constexpr (COMPACT) return closest_tri_call(a, b, c, m);
        else return closest_tri(a, b, c, m);
    };
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
        NBX_GJK_COUNT_ITER(it)
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
            nv = ctri(w, s0, s1, &m);
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
                    const V3 cp = ctri(P0, P1, P2, &m3);                               \
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
    bool enclosed;  // overlap proven by a tetrahedron around the origin: tw / ta hold it (the seed of the EPA below)
    int n;          // !enclosed overlap (contact regime): the origin lies within 5e-5 of the hull of tw[0..n-1]
    V3 tw[4], ta[4];
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

// out-of-line copy for gjk_witness (never on a verdict path; its callers are instruction-fetch-bound: ncu no_instruction 2.2
// stalls per issue in pair_distances_kernel): five inlined copies are ~500 instructions of its loop
__device__ __noinline__ V3 closest_tri_bary_call(V3 a, V3 b, V3 c) {
    float l0, l1, l2;
    closest_tri_bary(a, b, c, &l0, &l1, &l2);
    return mk(l0, l1, l2);
}

template <int MAX_ITER>
__device__ GjkWit gjk_witness(const ShapeW& A, const ShapeW& B, bool generic_start = false) {
    GjkWit out;
    out.overlap = false;
    out.enclosed = false;
    V3 w[4], a[4];
    float lam[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { w[i] = mk(0, 0, 0); a[i] = mk(0, 0, 0); lam[i] = 0.f; }
    int n = 0;
    V3 v = A.c - B.c;
    float vv = dot(v, v);
    if (vv < 1e-20f) { v = mk(1.f, 0.f, 0.f); vv = 1.f; }
    if (generic_start) { v = mk(-0.3034f, 0.2939f, -0.9065f); vv = dot(v, v); }   // retries of core_penetration: break symmetry
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
            const V3 l = closest_tri_bary_call(w[0], w[1], w[2]);
            lam[0] = l.x; lam[1] = l.y; lam[2] = l.z;
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
                    const V3 l = closest_tri_bary_call(w[I0], w[I1], w[I2]);                   \
                    const float l0 = l.x, l1 = l.y, l2 = l.z;                                  \
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
            if (!outside_any) { out.overlap = true; out.enclosed = true; break; }
        }
        V3 nv = mk(0, 0, 0), pa = mk(0, 0, 0);
#pragma unroll
        for (int i = 0; i < 4; ++i) { nv = madd(nv, lam[i], w[i]); pa = madd(pa, lam[i], a[i]); }
        const float nvv = dot(nv, nv);
        if (nvv >= best_vv) break;                       // rounding stopped the descent: keep the best point so far
        best_vv = nvv; best_v = nv; best_pa = pa;
        v = nv; vv = nvv;
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
        if (n == 4) { out.overlap = true; out.enclosed = true; break; }
        if (vv <= 2.5e-9f) { out.overlap = true; break; }   // contact regime (see gjk_distance): origin on the hull of w[0..n-1]
    }
    out.n = n;
    if (out.overlap || best_vv == FLT_MAX) {
        out.overlap = true;
        out.d = 0.f;
        out.pa = A.c; out.pb = B.c;
#pragma unroll
        for (int i = 0; i < 4; ++i) { out.tw[i] = w[i]; out.ta[i] = a[i]; }
        return out;
    }
    out.d = sqrtf(best_vv);
    out.pa = best_pa;
    out.pb = best_pa - best_v;
    return out;
}

// ---------------------------------------------------------------------------------------------
// EPA: penetration depth of two OVERLAPPING cores = distance from the origin to the boundary of the Minkowski difference
// D = A - B = min over unit n of h_D(n), h_D(n) = n . (s_A(n) - s_B(-n)) (van den Bergen's expanding polytope; what Bullet
// runs behind getClosestPoints once its GJK reports an intersection, reference numbotics/physics/chain.py:944-951).  The
// signed distance of the rounded-core model continues below -(mA + mB) as  -(mA + mB) - depth(cores)  (DESIGN.md section 4).
// Rare path (only pairs whose cores interpenetrate reach it): out of line, the polytope lives in local memory --
// kEpaV vertices, 2 kEpaV - 4 faces -- and the iteration count is bounded by the vertex budget.
// Result: depth >= 0, n = unit normal from B (target) to A (subject), witness points with  pa - pb = -depth * n.
// ---------------------------------------------------------------------------------------------
constexpr int kEpaV = 48;
constexpr int kEpaF = 2 * kEpaV;          // a closed triangulated polytope of V vertices has 2V - 4 faces

struct EpaOut {
    float depth;
    V3 n, pa, pb;
    bool ok;
};

struct EpaFace {
    unsigned char v0, v1, v2, pad;
    float nx, ny, nz, d;
    float pd;       // plane offset n . W[v0]: the visibility test of an expansion needs no vertex gather
};

__device__ __forceinline__ bool epa_make_face(EpaFace& f, const V3* W, int i0, int i1, int i2, V3 inside) {
    V3 n = cross(W[i1] - W[i0], W[i2] - W[i0]);
    const float l2 = dot(n, n);
    if (!(l2 > 1e-30f)) return false;
    n = rsqrtf(l2) * n;
    f.v0 = (unsigned char)i0; f.v1 = (unsigned char)i1; f.v2 = (unsigned char)i2;
    if (dot(n, W[i0] - inside) < 0.f) { f.v1 = (unsigned char)i2; f.v2 = (unsigned char)i1; n = -n; }
    f.nx = n.x; f.ny = n.y; f.nz = n.z;
    // distance from the origin to the TRIANGLE, not to its plane: the two agree on a convex polytope (the origin projects
    // into its closest face), but in float32 a sliver whose plane passes near the origin while the triangle itself lies far
    // away would otherwise be taken for the closest face -- and witnesses read off it are not points of the shapes
    float l0, l1, l2f;
    closest_tri_bary(W[f.v0], W[f.v1], W[f.v2], &l0, &l1, &l2f);
    const V3 cp = madd(madd(l0 * W[f.v0], l1, W[f.v1]), l2f, W[f.v2]);
    f.pd = dot(n, W[i0]);
    f.d = fmaxf(sqrtf(dot(cp, cp)), f.pd);
    f.pad = 0;
    return true;
}

__device__ __noinline__ EpaOut epa_penetration(const ShapeW& A, const ShapeW& B, const V3* sw, const V3* sa) {
    EpaOut out;
    out.ok = false;
    out.depth = 0.f;
    out.n = mk(0.f, 0.f, 1.f);
    out.pa = A.c; out.pb = B.c;
    V3 W[kEpaV], PA[kEpaV];
    EpaFace F[kEpaF];
    unsigned char E0[3 * 24], E1[3 * 24];         // horizon candidates: edges of the faces that see the new vertex
    unsigned long long adj[kEpaV];                // ... as a bit matrix: bit b of adj[a] = the directed edge a -> b is among them
#pragma unroll 1
    for (int i = 0; i < kEpaV; ++i) adj[i] = 0ull;
    int nv = 4, nf = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { W[i] = sw[i]; PA[i] = sa[i]; }
    const V3 inside = 0.25f * (W[0] + W[1] + W[2] + W[3]);
    if (!epa_make_face(F[nf++], W, 0, 1, 2, inside)) return out;
    if (!epa_make_face(F[nf++], W, 0, 1, 3, inside)) return out;
    if (!epa_make_face(F[nf++], W, 0, 2, 3, inside)) return out;
    if (!epa_make_face(F[nf++], W, 1, 2, 3, inside)) return out;
    float best_h = FLT_MAX, lower = 0.f;
    EpaFace bf = F[0];
#pragma unroll 1
    for (int it = 0; it < kEpaV; ++it) {
        int f = -1;
        float fd = FLT_MAX;
        for (int i = 0; i < nf; ++i)
            if (F[i].d < fd) { fd = F[i].d; f = i; }
        if (f < 0) break;
        const V3 n = mk(F[f].nx, F[f].ny, F[f].nz);
        const V3 a = support(A, n);
        const V3 w = a - support(B, -n);
        const float h = dot(n, w);
        lower = fd;
        if (h < best_h) { best_h = h; bf = F[f]; }
        // the bracket [closest face, smallest h] is closed to a fraction of the stated 1e-4 m band: stop BEFORE the new vertex
        // lands in the plane of its face (the slivers that come out of such an expansion have garbage normals).  Converged: the
        // CLOSEST face carries the normal and the witness points (an earlier face can own the smallest h by a hair while its
        // own plane is millimetres closer to the origin -- witnesses taken from it would not be points of the shapes)
        NBX_EPA_COUNT_EXPANSION(it)
        if (h - fd <= NBX_EPA_TOL_ABS + NBX_EPA_TOL_REL * fabsf(h)) { bf = F[f]; break; }
        if (nv >= kEpaV) break;
        W[nv] = w; PA[nv] = a;
        int ne = 0, nvis = 0;
        const float eps = 1e-6f * (1.f + fabsf(h));
        // faces that see the new vertex leave the list at once (the last face takes their slot): F[0 .. nf) is always exactly
        // the current polytope, <= 2 nv - 4 faces, and both scans of an expansion are as short as they can be
        for (int i = 0; i < nf;) {
            if (dot(mk(F[i].nx, F[i].ny, F[i].nz), w) - F[i].pd > eps) {
                if (ne + 3 > 3 * 24) { ne = -1; break; }       // absurdly large visible region: stop with the bracket found so far
                ++nvis;
                const int a0 = F[i].v0, a1 = F[i].v1, a2 = F[i].v2;
                E0[ne] = (unsigned char)a0; E1[ne] = (unsigned char)a1; ++ne;
                E0[ne] = (unsigned char)a1; E1[ne] = (unsigned char)a2; ++ne;
                E0[ne] = (unsigned char)a2; E1[ne] = (unsigned char)a0; ++ne;
                adj[a0] |= 1ull << a1;
                adj[a1] |= 1ull << a2;
                adj[a2] |= 1ull << a0;
                --nf;
                if (i != nf) F[i] = F[nf];
            } else {
                ++i;
            }
        }
#ifdef EPA_TRACE
        printf("  it %d nv %d nf %d fd %.6f h %.6f nvis %d ne %d\n", it, nv, nf, fd, h, nvis, ne);
#endif
        if (ne <= 0 || nvis == 0) break;
        // an edge whose reverse belongs to another removed face is interior to the visible region, not on its horizon
        // (one bit test per edge; the pairwise search this replaces was a tenth of the distance kernel's instructions)
        unsigned long long interior = 0ull;        // bit i: candidate i is interior (ne <= 72 > 64: two words)
        unsigned long long interior_hi = 0ull;
        for (int i = 0; i < ne; ++i) {
            const bool twin = (adj[E1[i]] >> E0[i]) & 1ull;
            if (i < 64) interior |= (unsigned long long)twin << i; else interior_hi |= (unsigned long long)twin << (i - 64);
        }
        for (int i = 0; i < ne; ++i) adj[E0[i]] = 0ull;
        if (nf + ne > kEpaF) break;                // cannot happen for a closed polytope (2 nv - 4 faces); a guard, not a path
        int added = 0;
        bool garbage = false;
        for (int i = 0; i < ne; ++i) {
            if ((i < 64 ? interior >> i : interior_hi >> (i - 64)) & 1ull) continue;
            if (epa_make_face(F[nf], W, E0[i], E1[i], nv, inside)) {
                // a face of the grown (convex) polytope cannot be closer to the origin than the closest face before
                garbage = garbage || F[nf].d < fd - (1e-5f + 1e-4f * fabsf(h));
                ++nf; ++added;
            }
        }
        ++nv;
#ifdef EPA_TRACE
        if (added < 3 || garbage) printf("  stop: added %d garbage %d\n", added, (int)garbage);
#endif
        if (added < 3 || garbage) break;
    }
    if (best_h == FLT_MAX) return out;
    const V3 n = mk(bf.nx, bf.ny, bf.nz);
    // The depth lies in [lower, best_h]: closest face of the inscribed polytope / smallest support value seen.  best_h is
    // reported: h_D(n) of an actual direction is attained exactly where the minimum sits on a flat feature or on a whole
    // circle of normals (coaxial cylinders, a point on a cylinder's axis), where the polygonal lower bound of a 48-vertex
    // polytope is still 0.5 % short; on generic curved contacts both ends are within 1e-5 at termination anyway.
    (void)lower;
    const float d = fmaxf(best_h, 0.f);
    float l0, l1, l2;
    closest_tri_bary(W[bf.v0], W[bf.v1], W[bf.v2], &l0, &l1, &l2);       // the point of the best face closest to the origin
    out.pa = madd(madd(l0 * PA[bf.v0], l1, PA[bf.v1]), l2, PA[bf.v2]);
    out.n = -n;
    out.pb = madd(out.pa, -d, n);
    out.depth = d;
    out.ok = true;
    return out;
}

// Penetration of two cores known (or suspected) to overlap: seed GJK with witness bookkeeping, then the EPA.  A seed that
// degenerates -- the origin within the contact resolution of a vertex / edge / face of the GJK simplex: symmetric scenes
// produce that exactly, and a triangle of three early support points passes within 5e-5 of the origin about once in 10^4
// generic deep overlaps -- is retried with A shifted by a generic 0.3 mm (then 2 mm) and a generic first search direction.
// The shift is taken out again at the end: with the face normal n found on the shifted problem the depth is re-evaluated
// on the UNSHIFTED cores, depth = h_D(n) = n . (s_A(n) - s_B(-n)), which is exact for the same contact feature and off by
// O(shift^2 / curvature radius) otherwise.
// point core (sphere primitive) inside a box / cylinder core: closed form.  Returns the depth, `dir` = the unit direction
// that takes the POINT out of the shape by the shortest way.
__device__ __forceinline__ float point_in_core(V3 p, const ShapeW& S, V3* dir) {
    const V3 r = p - S.c;
    if (S.type == NBX_PRIM_BOX) {
        float best = FLT_MAX;
        V3 bd = mk(0.f, 0.f, 1.f);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const V3 a = mk(S.ax[k], S.ax[3 + k], S.ax[6 + k]);
            const float x = dot(r, a);
            const float g = S.d[k] - fabsf(x);
            if (g < best) { best = g; bd = copysignf(1.f, x) * a; }
        }
        *dir = bd;
        return fmaxf(best, 0.f);
    }
    const V3 u = mk(S.ax[0], S.ax[1], S.ax[2]);
    const float z = dot(r, u);
    const V3 rad = madd(r, -z, u);
    const float rho = sqrtf(dot(rad, rad));
    const float gr = S.d[0] - rho, gz = S.d[1] - fabsf(z);
    if (gr <= gz) {
        // on the axis every radial direction is equally short: pick one square to u
        V3 e = fabsf(u.x) <= fabsf(u.y) && fabsf(u.x) <= fabsf(u.z) ? mk(1.f, 0.f, 0.f) : (fabsf(u.y) <= fabsf(u.z) ? mk(0.f, 1.f, 0.f) : mk(0.f, 0.f, 1.f));
        e = cross(u, e);
        *dir = rho > 1e-7f * S.d[0] ? (1.f / rho) * rad : rsqrtf(dot(e, e)) * e;
        return fmaxf(gr, 0.f);
    }
    *dir = copysignf(1.f, z) * u;
    return fmaxf(gz, 0.f);
}

__device__ __noinline__ EpaOut core_penetration(const ShapeW& A, const ShapeW& B) {
    EpaOut out;
    out.ok = false; out.depth = 0.f; out.n = mk(0.f, 0.f, 1.f); out.pa = A.c; out.pb = B.c;
    if (A.type == NBX_PRIM_SPHERE || B.type == NBX_PRIM_SPHERE) {
        // a point core in a box / cylinder core (two points never overlap): exact, no polytope
        out.ok = true;
        if (A.type == NBX_PRIM_SPHERE && B.type == NBX_PRIM_SPHERE) return out;
        V3 dir;
        if (A.type == NBX_PRIM_SPHERE) {
            out.depth = point_in_core(A.c, B, &dir);
            out.n = dir;                                   // A (the point) leaves B along dir
            out.pa = A.c;
            out.pb = madd(A.c, out.depth, dir);
        } else {
            out.depth = point_in_core(B.c, A, &dir);
            out.n = -dir;                                  // A moves the other way
            out.pb = B.c;
            out.pa = madd(B.c, out.depth, dir);            // pa - pb = -depth * n
        }
        return out;
    }
    if (A.type == NBX_PRIM_CYLINDER && B.type == NBX_PRIM_CYLINDER) {
        // parallel axes (to 1e-4 rad): A - B is itself a cylinder of radius rA + rB and half height hA + hB about cA - cB, and
        // the minimum over the normals sits on a whole circle -- exact in closed form, a 0.1 % error for a 48-vertex polytope
        const V3 ua = mk(A.ax[0], A.ax[1], A.ax[2]), ub = mk(B.ax[0], B.ax[1], B.ax[2]);
        const V3 cx = cross(ua, ub);
        if (dot(cx, cx) < 1e-8f) {
            ShapeW D = B;
            D.d[0] = A.d[0] + B.d[0]; D.d[1] = A.d[1] + B.d[1];
            V3 dir;
            out.depth = point_in_core(A.c, D, &dir);
            out.n = dir;
            out.ok = true;
            const V3 r = A.c - B.c;
            const float z = dot(r, ub);
            const V3 rad = madd(r, -z, ub);
            const float rho = sqrtf(dot(rad, rad));
            if (fabsf(dot(dir, ub)) < 0.5f) {
                // side contact: on A's side facing B's axis, half way along the axial overlap
                const float zlo = fmaxf(z - A.d[1], -B.d[1]), zhi = fminf(z + A.d[1], B.d[1]);
                out.pa = madd(madd(A.c, -A.d[0], dir), 0.5f * (zlo + zhi) - z, ub);
            } else {
                // cap contact: A's cap centre, pulled towards B's axis until it is over B's cap
                const float t = fmaxf(rho - B.d[0], 0.f);
                out.pa = madd(A.c, -A.d[1], dir);
                if (rho > 0.f) out.pa = madd(out.pa, -t / rho, rad);
            }
            out.pb = madd(out.pa, out.depth, out.n);
            return out;
        }
    }
    ShapeW As = A;
#pragma unroll 1
    for (int attempt = 0; attempt < 3; ++attempt) {
        const GjkWit g = gjk_witness<64>(As, B, attempt > 0);
#ifdef EPA_TRACE
        printf(" attempt %d overlap %d enclosed %d d %.6g\n", attempt, (int)g.overlap, (int)g.enclosed, g.d);
#endif
        if (!g.overlap) {
            // Separated: a grazing contact (attempt 0), or the shift pulled a shallow overlap apart.  The witness normal of the
            // (shifted) pair is the contact normal; the penetration along it on the UNSHIFTED cores, h_D(-n), is an upper bound
            // of the depth and exact when the contact feature is the same -- no polytope needed for overlaps this shallow.
            const float l = g.d > 0.f ? g.d : 1.f;
            out.n = (1.f / l) * (g.pa - g.pb);
            if (attempt > 0) {
                const V3 nf = -out.n;
                const V3 a = support(A, nf);
                const V3 sh = As.c - A.c;
                // the shifted pair is separated, so the unshifted one cannot be deeper than the shift is long
                out.depth = fminf(fmaxf(dot(nf, a - support(B, -nf)), 0.f), sqrtf(dot(sh, sh)));
            }
            out.pa = g.pa - (As.c - A.c);
            out.pb = madd(out.pa, out.depth, out.n);
            out.ok = true;
            return out;
        }
        V3 tw[4], ta[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { tw[i] = g.tw[i]; ta[i] = g.ta[i]; }
        bool seeded = g.enclosed;
        if (!seeded && g.n >= 2) {
            // The origin sits (to 5e-5) ON a segment or triangle of the seed simplex.  That is grazing contact only if the
            // simplex is part of the boundary; for a deep overlap it is a chord through the interior (a triangle of three early
            // support points of two big cylinders does that routinely).  Complete it to a tetrahedron: a support point square
            // to the segment, then one beyond the triangle's plane on the origin's side.
            if (g.n == 2) {
                const V3 e = tw[1] - tw[0];
                const V3 ax = fabsf(e.x) <= fabsf(e.y) && fabsf(e.x) <= fabsf(e.z) ? mk(1.f, 0.f, 0.f)
                              : (fabsf(e.y) <= fabsf(e.z) ? mk(0.f, 1.f, 0.f) : mk(0.f, 0.f, 1.f));
                const V3 d = cross(e, ax);
                ta[2] = support(As, d);
                tw[2] = ta[2] - support(B, -d);
            }
            V3 nrm = cross(tw[1] - tw[0], tw[2] - tw[0]);
            const float l2 = dot(nrm, nrm);
            if (l2 > 1e-20f) {
                if (dot(nrm, tw[0]) > 0.f) nrm = -nrm;        // towards the side of the plane the origin is on
                ta[3] = support(As, nrm);
                tw[3] = ta[3] - support(B, -nrm);
                // a real tetrahedron (the 4th point clear of the plane), and the origin not beyond it: else this IS the boundary
                const float hgt = dot(nrm, tw[3] - tw[0]) * rsqrtf(l2);
                seeded = hgt > 1e-4f;
            }
        }
        if (seeded) {
            out = epa_penetration(As, B, tw, ta);
            if (out.ok) {
                if (attempt > 0) {
                    const V3 nf = -out.n;           // outward normal of the face of A - B closest to the origin
                    const V3 a = support(A, nf);
                    out.depth = fmaxf(dot(nf, a - support(B, -nf)), 0.f);
                    out.pa = out.pa - (As.c - A.c);
                }
                out.pb = madd(out.pa, out.depth, out.n);      // pa - pb = -depth * n
                return out;
            }
        }
        As.c = A.c + (attempt == 0 ? 3e-4f : 2e-3f) * (attempt == 0 ? mk(0.5377f, -0.8311f, 0.1418f) : mk(-0.6245f, -0.2171f, 0.7502f));
    }
    // no usable seed: report contact along the line of centres
    const V3 dc = A.c - B.c;
    const float l = sqrtf(dot(dc, dc));
    out.n = l > 0.f ? (1.f / l) * dc : mk(0.f, 0.f, 1.f);
    out.ok = true;
    return out;
}

}  // namespace nbxg
