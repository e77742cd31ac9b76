/*
 * nbx_oracle.c -- float64 CPU restatement of the hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing under numbotics_b200/ may import, link or execute this file; it exists so that
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs have
 * something to compare the CUDA path with (and to time beside it).
 *
 * What it restates (paths relative to the reference checkout):
 *   - forward kinematics + geometric Jacobian      numbotics/robots/helpers.py:33-187
 *     (Rodrigues joint transform :43-55, chain product :105-108, Jacobian columns :172-185)
 *   - Arm.in_collision verdict                      numbotics/robots/arm.py:555-604
 *     min over allowed pairs of closest distance < threshold.  The distance itself comes from
 *     PyBullet 3.2.7 (poetry.lock:1659) getClosestPoints -- third-party, not vendored, not
 *     installable offline -- so it is restated from the published algorithm: GJK closest
 *     distance between convex cores, minus the two collision margins (Bullet's
 *     btGjkPairDetector convention).  PARITY UNPINNED against real Bullet; pinned instead
 *     by analytic known answers and Monte-Carlo / SLSQP cross-checks in tests/.
 *   - DiscreteConnector.connect / steer             numbotics/planning/sampling_based/connectors.py:57-100
 *
 * The GJK here is deliberately NOT the one in numbotics_b200/csrc/nbx_gjk.cuh: the closest
 * point of a simplex is found by enumerating all its sub-simplices and solving the affine
 * projection of each (Johnson's sub-algorithm in its brute-force form), in float64, run to
 * convergence.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PRIM_SPHERE 0
#define PRIM_BOX 1
#define PRIM_CYL 2
#define JOINT_REVOLUTE 0
#define JOINT_PRISMATIC 1

typedef struct {
    int type;
    double c[3];
    double R[9];    /* world rotation, row-major; column k = local axis k */
    double d[3];    /* core dims */
} oshape;

static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

static void support(const oshape* S, const double* dir, double* out) {
    out[0] = S->c[0]; out[1] = S->c[1]; out[2] = S->c[2];
    if (S->type == PRIM_SPHERE) return;
    if (S->type == PRIM_BOX) {
        for (int k = 0; k < 3; ++k) {
            const double a[3] = {S->R[k], S->R[3 + k], S->R[6 + k]};
            const double s = dot3(dir, a) >= 0.0 ? S->d[k] : -S->d[k];
            for (int i = 0; i < 3; ++i) out[i] += s * a[i];
        }
        return;
    }
    {
        const double u[3] = {S->R[2], S->R[5], S->R[8]};
        const double du = dot3(dir, u);
        double perp[3] = {dir[0] - du * u[0], dir[1] - du * u[1], dir[2] - du * u[2]};
        /* remove the rounding residue along the axis (dir almost parallel to u: cap contacts).  u comes from a rotation
         * matrix that may have been rounded to float32 (|u|^2 - 1 ~ 6e-8), so one pass leaves a residue of that relative
         * size along u: two passes, and a direction within 1e-9 rad of the axis is treated as the axis (every cap point is
         * a support point then; the support VALUE changes by radius * 1e-9).  With a 1e-13 cut-off an EPA face lying
         * exactly in the cap plane produced a "support point" 3 mm outside the cylinder (noise direction with a 1 %
         * axial component times the radius), i.e. a polytope that was not inscribed and a depth 4 % too large. */
        for (int pass = 0; pass < 2; ++pass) {
            const double pu = dot3(perp, u) / dot3(u, u);
            for (int i = 0; i < 3; ++i) perp[i] -= pu * u[i];
        }
        const double l = sqrt(dot3(perp, perp));
        const double s = du >= 0.0 ? S->d[1] : -S->d[1];
        for (int i = 0; i < 3; ++i) out[i] += s * u[i];
        if (l > 1e-9 * sqrt(dot3(dir, dir)))
            for (int i = 0; i < 3; ++i) out[i] += S->d[0] / l * perp[i];
    }
}

/* Solve the k x k system for the affine projection of the origin on the hull of pts[idx[0..k-1]].
 * Returns 1 and barycentric weights lam if the system is regular. */
static int affine_project(const double pts[4][3], const int* idx, int k, double* lam) {
    if (k == 1) { lam[0] = 1.0; return 1; }
    /* unknowns lam_1..lam_{k-1} relative to p0:  x = p0 + sum lam_i (p_i - p0),  (p_j - p0).x = 0 */
    double E[3][3], G[3][3], r[3];
    const double* p0 = pts[idx[0]];
    const int m = k - 1;
    for (int i = 0; i < m; ++i)
        for (int c = 0; c < 3; ++c) E[i][c] = pts[idx[i + 1]][c] - p0[c];
    for (int i = 0; i < m; ++i) {
        for (int j = 0; j < m; ++j) G[i][j] = dot3(E[i], E[j]);
        r[i] = -dot3(E[i], p0);
    }
    /* Gaussian elimination with partial pivoting on the (<= 3x3) Gram system */
    double scale = 0.0;
    for (int i = 0; i < m; ++i) if (G[i][i] > scale) scale = G[i][i];
    if (scale <= 0.0) return 0;
    for (int c = 0; c < m; ++c) {
        int piv = c;
        for (int i = c + 1; i < m; ++i) if (fabs(G[i][c]) > fabs(G[piv][c])) piv = i;
        if (fabs(G[piv][c]) <= 1e-14 * scale) return 0;
        if (piv != c) {
            for (int j = 0; j < m; ++j) { double t = G[c][j]; G[c][j] = G[piv][j]; G[piv][j] = t; }
            double t = r[c]; r[c] = r[piv]; r[piv] = t;
        }
        for (int i = c + 1; i < m; ++i) {
            const double f = G[i][c] / G[c][c];
            for (int j = c; j < m; ++j) G[i][j] -= f * G[c][j];
            r[i] -= f * r[c];
        }
    }
    double x[3];
    for (int i = m - 1; i >= 0; --i) {
        double s = r[i];
        for (int j = i + 1; j < m; ++j) s -= G[i][j] * x[j];
        x[i] = s / G[i][i];
    }
    double sum = 0.0;
    for (int i = 0; i < m; ++i) { lam[i + 1] = x[i]; sum += x[i]; }
    lam[0] = 1.0 - sum;
    return 1;
}

/* closest point of conv(pts[0..n-1]) to the origin by sub-simplex enumeration; shrinks the simplex.  `apts` (nullable):
 * the support points on A that belong to the simplex vertices, compacted alongside. */
static void simplex_closest(double pts[4][3], double (*apts)[3], int* n, double* v) {
    double best = INFINITY;
    int best_mask = 1;
    double best_lam[4] = {1, 0, 0, 0};
    for (int mask = 1; mask < (1 << *n); ++mask) {
        int idx[4], k = 0;
        for (int i = 0; i < *n; ++i) if (mask & (1 << i)) idx[k++] = i;
        double lam[4];
        if (!affine_project(pts, idx, k, lam)) continue;
        int ok = 1;
        for (int i = 0; i < k; ++i) if (lam[i] < -1e-12) { ok = 0; break; }
        if (!ok) continue;
        double x[3] = {0, 0, 0};
        for (int i = 0; i < k; ++i)
            for (int c = 0; c < 3; ++c) x[c] += lam[i] * pts[idx[i]][c];
        const double dd = dot3(x, x);
        if (dd < best) {
            best = dd; best_mask = mask;
            for (int i = 0; i < 4; ++i) best_lam[i] = 0.0;
            for (int i = 0; i < k; ++i) best_lam[idx[i]] = lam[i];
        }
    }
    double np[4][3], na[4][3];
    int m = 0;
    v[0] = v[1] = v[2] = 0.0;
    for (int i = 0; i < *n; ++i) {
        if (!(best_mask & (1 << i))) continue;
        for (int c = 0; c < 3; ++c) {
            v[c] += best_lam[i] * pts[i][c]; np[m][c] = pts[i][c];
            if (apts) na[m][c] = apts[i][c];
        }
        ++m;
    }
    for (int i = 0; i < m; ++i) for (int c = 0; c < 3; ++c) { pts[i][c] = np[i][c]; if (apts) apts[i][c] = na[i][c]; }
    *n = m;
}

/* GJK distance between two convex cores, run to convergence.  *overlap = 1 when they intersect.
 * simplex_w / simplex_a / simplex_n (nullable): the final simplex (points of A - B, their supports on A, count) -- a
 * tetrahedron that encloses the origin when *overlap and *simplex_n == 4 (the seed of the EPA below). */
static double core_distance_ex(const oshape* A, const oshape* B, int* overlap, double (*simplex_w)[3], double (*simplex_a)[3],
                               int* simplex_n, const double* v_init) {
    double v[3] = {A->c[0] - B->c[0], A->c[1] - B->c[1], A->c[2] - B->c[2]};
    if (v_init) { v[0] = v_init[0]; v[1] = v_init[1]; v[2] = v_init[2]; }
    double pts[4][3], apts[4][3];
    int n = 0;
    *overlap = 0;
    if (simplex_n) *simplex_n = 0;
    if (dot3(v, v) < 1e-30) { v[0] = 1.0; v[1] = v[2] = 0.0; }
    double vv = dot3(v, v);
    int have = 0;
    for (int it = 0; it < 500; ++it) {
        double nd[3] = {-v[0], -v[1], -v[2]}, sa[3], sb[3], w[3];
        support(A, nd, sa);
        support(B, v, sb);
        for (int c = 0; c < 3; ++c) w[c] = sa[c] - sb[c];
        const double vw = dot3(v, w);
        if (have && vv - vw <= 1e-14 * vv + 1e-300) break;      /* |v|^2 - v.w is the duality gap */
        int dup = 0;
        for (int i = 0; i < n; ++i) {
            double e[3] = {pts[i][0] - w[0], pts[i][1] - w[1], pts[i][2] - w[2]};
            if (dot3(e, e) <= 1e-28) dup = 1;
        }
        if (dup) break;
        for (int c = 0; c < 3; ++c) { pts[n][c] = w[c]; apts[n][c] = sa[c]; }
        ++n;
        double nv[3];
        simplex_closest(pts, apts, &n, nv);
        const double nvv = dot3(nv, nv);
        if (n == 4 || nvv <= 1e-24) {
            *overlap = 1;
            if (simplex_n) {
                *simplex_n = n;
                for (int i = 0; i < n; ++i) for (int c = 0; c < 3; ++c) { simplex_w[i][c] = pts[i][c]; simplex_a[i][c] = apts[i][c]; }
            }
            return 0.0;
        }
        if (have && nvv >= vv) break;
        v[0] = nv[0]; v[1] = nv[1]; v[2] = nv[2];
        vv = nvv; have = 1;
    }
    return sqrt(vv);
}

static double core_distance(const oshape* A, const oshape* B, int* overlap) {
    return core_distance_ex(A, B, overlap, NULL, NULL, NULL, NULL);
}

/* ------------------------------------------------------------------------------------------
 * EPA: penetration depth of two OVERLAPPING cores = the distance from the origin to the boundary of the Minkowski
 * difference D = A - B, min over unit n of h_D(n), h_D(n) = n . (s_A(n) - s_B(-n)).  Restates what Bullet does behind
 * getClosestPoints once its GJK reports an intersection (btGjkEpaPenetrationDepthSolver -> btGjkEpaSolver2::Penetration,
 * the published expanding-polytope algorithm of van den Bergen); Bullet runs it on the shapes WITH their margins, and
 * depth(core + ball(mA), core + ball(mB)) = depth(cores) + mA + mB, so the signed distance of the rounded-core model
 * continues below -(mA + mB) as  -(mA + mB) - depth(cores).  PARITY UNPINNED like every collision distance here; pinned
 * instead by closed forms (point in box / cylinder, SAT for box pairs) and a brute-force minimisation of h_D (tests/).
 * Output: depth >= 0, `nrm` = unit normal from B (target) to A (subject) = the direction that moves A out of B,
 * witness points pa on core A / pb on core B with  pa - pb = -depth * nrm.
 * ------------------------------------------------------------------------------------------ */
#define EPA_MAXV 1000
#define EPA_MAXF 4096

typedef struct { int v[3]; double n[3]; double d; int alive; } epa_face;

static void support_diff(const oshape* A, const oshape* B, const double* dir, double* w, double* a) {
    const double nd[3] = {-dir[0], -dir[1], -dir[2]};
    double sb[3];
    support(A, dir, a);
    support(B, nd, sb);
    for (int c = 0; c < 3; ++c) w[c] = a[c] - sb[c];
}

static int epa_make_face(epa_face* f, double (*W)[3], int i0, int i1, int i2, const double* inside) {
    double e1[3], e2[3], n[3];
    for (int c = 0; c < 3; ++c) { e1[c] = W[i1][c] - W[i0][c]; e2[c] = W[i2][c] - W[i0][c]; }
    n[0] = e1[1] * e2[2] - e1[2] * e2[1]; n[1] = e1[2] * e2[0] - e1[0] * e2[2]; n[2] = e1[0] * e2[1] - e1[1] * e2[0];
    const double len = sqrt(dot3(n, n));
    if (!(len > 1e-300)) return 0;
    for (int c = 0; c < 3; ++c) n[c] /= len;
    const double rel[3] = {W[i0][0] - inside[0], W[i0][1] - inside[1], W[i0][2] - inside[2]};
    f->v[0] = i0; f->v[1] = i1; f->v[2] = i2;
    if (dot3(n, rel) < 0.0) { f->v[1] = i2; f->v[2] = i1; for (int c = 0; c < 3; ++c) n[c] = -n[c]; }
    for (int c = 0; c < 3; ++c) f->n[c] = n[c];
    f->d = dot3(n, W[i0]);
    f->alive = 1;
    return 1;
}

/* barycentric coordinates of the point of triangle (p0,p1,p2) closest to x (Ericson 5.1.5, float64) */
static void tri_bary(const double* p0, const double* p1, const double* p2, const double* x, double* l) {
    double ab[3], ac[3], ap[3];
    for (int c = 0; c < 3; ++c) { ab[c] = p1[c] - p0[c]; ac[c] = p2[c] - p0[c]; ap[c] = x[c] - p0[c]; }
    const double d1 = dot3(ab, ap), d2 = dot3(ac, ap);
    l[0] = 1; l[1] = 0; l[2] = 0;
    if (d1 <= 0 && d2 <= 0) return;
    double bp[3] = {x[0] - p1[0], x[1] - p1[1], x[2] - p1[2]};
    const double d3 = dot3(ab, bp), d4 = dot3(ac, bp);
    if (d3 >= 0 && d4 <= d3) { l[0] = 0; l[1] = 1; return; }
    const double vc = d1 * d4 - d3 * d2;
    if (vc <= 0 && d1 >= 0 && d3 <= 0) { const double t = d1 / (d1 - d3); l[0] = 1 - t; l[1] = t; return; }
    double cp[3] = {x[0] - p2[0], x[1] - p2[1], x[2] - p2[2]};
    const double d5 = dot3(ab, cp), d6 = dot3(ac, cp);
    if (d6 >= 0 && d5 <= d6) { l[0] = 0; l[2] = 1; return; }
    const double vb = d5 * d2 - d1 * d6;
    if (vb <= 0 && d2 >= 0 && d6 <= 0) { const double t = d2 / (d2 - d6); l[0] = 1 - t; l[2] = t; return; }
    const double va = d3 * d6 - d5 * d4;
    if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) { const double t = (d4 - d3) / ((d4 - d3) + (d5 - d6)); l[0] = 0; l[1] = 1 - t; l[2] = t; return; }
    const double den = 1.0 / (va + vb + vc);
    l[1] = vb * den; l[2] = vc * den; l[0] = 1.0 - l[1] - l[2];
}

/* returns 1 on success */
static int epa_from_simplex(const oshape* A, const oshape* B, double (*sw)[3], double (*sa)[3], double* depth, double* nrm,
                            double* pa, double* pb) {
    static const double kTol = 1e-11;
    double (*W)[3] = malloc(sizeof(double[3]) * EPA_MAXV);
    double (*PA)[3] = malloc(sizeof(double[3]) * EPA_MAXV);
    epa_face* F = malloc(sizeof(epa_face) * EPA_MAXF);
    int (*edges)[2] = malloc(sizeof(int[2]) * 3 * EPA_MAXF);
    int nv = 4, nf = 0, ok = 0;
    for (int i = 0; i < 4; ++i) for (int c = 0; c < 3; ++c) { W[i][c] = sw[i][c]; PA[i][c] = sa[i][c]; }
    double inside[3] = {0, 0, 0};
    for (int i = 0; i < 4; ++i) for (int c = 0; c < 3; ++c) inside[c] += 0.25 * W[i][c];
    static const int tet[4][3] = {{0, 1, 2}, {0, 1, 3}, {0, 2, 3}, {1, 2, 3}};
    for (int t = 0; t < 4; ++t)
        if (!epa_make_face(&F[nf++], W, tet[t][0], tet[t][1], tet[t][2], inside)) goto done;   /* flat seed */
    int best = -1;
    double best_gap = INFINITY, best_h = 0.0, lower = 0.0;
    int best_face_copy_valid = 0;
    epa_face best_face;
    for (int it = 0; it < 4 * EPA_MAXV; ++it) {
        int f = -1;
        for (int i = 0; i < nf; ++i)
            if (F[i].alive && (f < 0 || F[i].d < F[f].d)) f = i;
        if (f < 0) goto done;
        double w[3], a[3];
        support_diff(A, B, F[f].n, w, a);
        const double h = dot3(F[f].n, w);
        const double gap = h - F[f].d;
        lower = F[f].d;                  /* closest face of the inscribed polytope: a LOWER bound, non-decreasing */
        /* every h = h_D(n) is an UPPER bound of the depth (translating by h along n separates the cores); keep the smallest */
        if (!best_face_copy_valid || h < best_h) { best_gap = gap; best_h = h; best_face = F[f]; best_face_copy_valid = 1; best = f; }
        if (gap <= kTol * (1.0 + fabs(h)) || nv >= EPA_MAXV || nf + 64 >= EPA_MAXF) break;
        /* expand the polytope by w: drop the faces that see it, close the hole along the horizon */
        for (int c = 0; c < 3; ++c) { W[nv][c] = w[c]; PA[nv][c] = a[c]; }
        int ne = 0, nvis = 0;
        for (int i = 0; i < nf; ++i) {
            if (!F[i].alive) continue;
            const double rel[3] = {w[0] - W[F[i].v[0]][0], w[1] - W[F[i].v[0]][1], w[2] - W[F[i].v[0]][2]};
            if (dot3(F[i].n, rel) > 1e-14 * (1.0 + fabs(h))) {
                F[i].alive = 0; ++nvis;
                for (int k = 0; k < 3; ++k) { edges[ne][0] = F[i].v[k]; edges[ne][1] = F[i].v[(k + 1) % 3]; ++ne; }
            }
        }
        if (nvis == 0) break;                     /* w is on the hull already: converged to rounding */
        int added = 0;
        for (int i = 0; i < ne; ++i) {
            int twin = 0;
            for (int j = 0; j < ne; ++j)
                if (edges[j][0] == edges[i][1] && edges[j][1] == edges[i][0]) { twin = 1; break; }
            if (twin) continue;                   /* interior edge of the visible region */
            if (epa_make_face(&F[nf], W, edges[i][0], edges[i][1], nv, inside)) { ++nf; ++added; }
        }
        ++nv;
        if (added < 3) break;                     /* degenerate horizon: keep the best face found so far */
        (void)best;
    }
    if (best_face_copy_valid) {
        /* the depth lies in [closest face distance, smallest h]; the two agree to kTol at convergence, and the midpoint is
         * reported.  (Worst case: a minimum attained on a whole circle of normals -- coaxial cylinders -- where the bracket
         * closes only as 1 / vertices^2: ~5e-6 at the vertex cap.) */
        const epa_face* bf = &best_face;
        const double d = 0.5 * (fmax(fmin(lower, best_h), 0.0) + fmax(best_h, 0.0));
        double proj[3] = {bf->d * bf->n[0], bf->d * bf->n[1], bf->d * bf->n[2]}, l[3];
        tri_bary(W[bf->v[0]], W[bf->v[1]], W[bf->v[2]], proj, l);
        for (int c = 0; c < 3; ++c) {
            pa[c] = l[0] * PA[bf->v[0]][c] + l[1] * PA[bf->v[1]][c] + l[2] * PA[bf->v[2]][c];
            nrm[c] = -bf->n[c];
            pb[c] = pa[c] - d * bf->n[c];
        }
        *depth = d;
        ok = 1;
    }
done:
    free(W); free(PA); free(F); free(edges);
    return ok;
}

/* penetration depth of two overlapping cores (see above).  When the GJK simplex is degenerate -- the origin lies ON a
 * vertex / edge / face of it, which symmetric test scenes produce exactly -- A is shifted by a generic 1e-9 (then 1e-7)
 * vector and the search repeated: the depth changes by at most that much. */
static double core_penetration(const oshape* A, const oshape* B, double* nrm, double* pa, double* pb) {
    static const double gen[3] = {0.5377, -0.8311, 0.1418};
    static const double gen2[3] = {-0.3034, 0.2939, -0.9065};   /* first search direction of the retries: NOT along the shift,
                                                                    or the shifted problem stays symmetric about it */
    const double shifts[4] = {0.0, 1e-9, 1e-7, 1e-5};
    for (int s = 0; s < 4; ++s) {
        oshape As = *A;
        for (int c = 0; c < 3; ++c) As.c[c] += shifts[s] * gen[c];
        double sw[4][3], sa[4][3];
        int ov = 0, ns = 0;
        (void)core_distance_ex(&As, B, &ov, sw, sa, &ns, s == 0 ? NULL : gen2);
        if (!ov) {                       /* the shift separated a grazing contact: depth below the shift */
            nrm[0] = 0; nrm[1] = 0; nrm[2] = 1; for (int c = 0; c < 3; ++c) { pa[c] = A->c[c]; pb[c] = A->c[c]; }
            double d = As.c[0] - B->c[0], e = As.c[1] - B->c[1], f = As.c[2] - B->c[2], l = sqrt(d * d + e * e + f * f);
            if (l > 0) { nrm[0] = d / l; nrm[1] = e / l; nrm[2] = f / l; }
            return 0.0;
        }
        if (ns != 4) continue;
        double depth;
        if (epa_from_simplex(&As, B, sw, sa, &depth, nrm, pa, pb)) {
            for (int c = 0; c < 3; ++c) pa[c] -= shifts[s] * gen[c];
            return depth;
        }
    }
    nrm[0] = 0; nrm[1] = 0; nrm[2] = 1;
    for (int c = 0; c < 3; ++c) { pa[c] = A->c[c]; pb[c] = B->c[c]; }
    return 0.0;
}

static void make_shape(oshape* S, int type, const double* pose12, const double* dims_full, double margin, int sharp) {
    S->type = type;
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) S->R[r * 3 + c] = pose12[r * 4 + c];
        S->c[r] = pose12[r * 4 + 3];
    }
    const double m = (sharp || type == PRIM_SPHERE) ? 0.0 : margin;
    if (type == PRIM_SPHERE) { S->d[0] = S->d[1] = S->d[2] = 0.0; }
    else if (type == PRIM_BOX) { for (int k = 0; k < 3; ++k) S->d[k] = dims_full[k] - m; }
    else { S->d[0] = dims_full[0] - m; S->d[1] = dims_full[1] - m; S->d[2] = 0.0; }
}

/* distance of one pair under the rounded-core model; sharp => margins ignored except for spheres */
static double pair_distance(int ta, const double* pa, const double* da, double ma, int tb, const double* pb,
                            const double* db, double mb, int sharp) {
    oshape A, B;
    make_shape(&A, ta, pa, da, ma, sharp);
    make_shape(&B, tb, pb, db, mb, sharp);
    const double mA = ta == PRIM_SPHERE ? da[0] : (sharp ? 0.0 : ma);
    const double mB = tb == PRIM_SPHERE ? db[0] : (sharp ? 0.0 : mb);
    int ov;
    const double d = core_distance(&A, &B, &ov);
    if (ov) {
        /* overlapping cores: the signed distance continues below -(mA + mB) by the cores' penetration depth (EPA) */
        double nrm[3], pa2[3], pb2[3];
        return -(mA + mB) - core_penetration(&A, &B, nrm, pa2, pb2);
    }
    return d - mA - mB;
}

/* ------------------------------------------------------------------------------------------
 * exported: pair distances for given world poses (the checker path: FK comes from the NumPy
 * oracle, only the narrow phase runs here)
 *   poses  (n, n_link, 12) link-primitive world poses per configuration
 *   sposes (n_static, 12)  obstacle poses; shape index space = [0,n_link) U [n_link, n_link+n_static)
 *   out    (n, n_pairs)
 * ------------------------------------------------------------------------------------------ */
int oracle_pair_distances(int64_t n, int n_link, const double* poses, int n_static, const double* sposes,
                          const int32_t* types, const double* dims, const double* margins, int n_pairs,
                          const int32_t* pairs, const uint8_t* sharp, double* out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        for (int p = 0; p < n_pairs; ++p) {
            const int a = pairs[2 * p], b = pairs[2 * p + 1];
            const double* pa = a < n_link ? poses + ((size_t)i * n_link + a) * 12 : sposes + (size_t)(a - n_link) * 12;
            const double* pb = b < n_link ? poses + ((size_t)i * n_link + b) * 12 : sposes + (size_t)(b - n_link) * 12;
            out[(size_t)i * n_pairs + p] = pair_distance(types[a], pa, dims + 3 * a, margins[a], types[b], pb,
                                                         dims + 3 * b, margins[b], sharp ? sharp[p] : 0);
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * full pipeline on the flattened descriptor (same arrays as include/nbx.h); used for the CPU
 * baseline timing and cross-checked against the checker path in tests/.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int dof, n_mov, n_frames, n_prims;
    const int32_t *mov_parent, *mov_type, *mov_qidx;
    const double *mov_offset, *mov_axis;
    const int32_t* frame_mov;
    const double *frame_tail, *frame_com;
    const int32_t *prim_frame, *prim_type;
    const double *prim_dims, *prim_offset, *prim_margin;
} ochain;

typedef struct {
    int n_obs;
    const int32_t* obs_type;
    const double *obs_dims, *obs_pose, *obs_margin;
    int n_self;
    const int32_t* self_pairs;
    int n_env;
    const int32_t* env_pairs;
    const uint8_t *self_sharp, *env_sharp;
} oscene;

static void mul12(const double* a, const double* b, double* o) {
    double r[12];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) r[i * 4 + j] = a[i * 4] * b[j] + a[i * 4 + 1] * b[4 + j] + a[i * 4 + 2] * b[8 + j];
        r[i * 4 + 3] = a[i * 4] * b[3] + a[i * 4 + 1] * b[7] + a[i * 4 + 2] * b[11] + a[i * 4 + 3];
    }
    memcpy(o, r, sizeof(r));
}

/* X(q) of one joint: helpers.py:43-55 (revolute), intended form of :57-65 (prismatic) */
static void joint_x(const double* off, const double* a, int type, double q, double* X) {
    if (type == JOINT_REVOLUTE) {
        const double s = sin(q), c = cos(q);
        double M[9];
        const double K[9] = {a[0] * a[0], a[0] * a[1], a[0] * a[2], a[1] * a[0], a[1] * a[1], a[1] * a[2], a[2] * a[0], a[2] * a[1], a[2] * a[2]};
        const double S[9] = {0, -a[2], a[1], a[2], 0, -a[0], -a[1], a[0], 0};
        for (int i = 0; i < 9; ++i) M[i] = K[i] - c * (K[i] - ((i % 4) == 0 ? 1.0 : 0.0)) + s * S[i];
        for (int r = 0; r < 3; ++r) {
            for (int cc = 0; cc < 3; ++cc) X[r * 4 + cc] = off[r * 4] * M[cc] + off[r * 4 + 1] * M[3 + cc] + off[r * 4 + 2] * M[6 + cc];
            X[r * 4 + 3] = off[r * 4 + 3];
        }
    } else {
        for (int r = 0; r < 3; ++r) {
            for (int cc = 0; cc < 3; ++cc) X[r * 4 + cc] = off[r * 4 + cc];
            X[r * 4 + 3] = off[r * 4 + 3] + q * (off[r * 4] * a[0] + off[r * 4 + 1] * a[1] + off[r * 4 + 2] * a[2]);
        }
    }
}

/* FK + Jacobian of one frame, float64.  T_out (n,4,4), J_out (n,6,dof), both nullable. */
int oracle_fk_jac(const ochain* ch, int frame, const double* q, int64_t n, const double* base16, int use_com,
                  double* T_out, double* J_out) {
    int path[256], np = 0;
    for (int j = ch->frame_mov[frame]; j >= 0; j = ch->mov_parent[j]) path[np++] = j;
    const int dof = ch->dof;
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < n; ++b) {
        double T[12], X[12], frames[256][12];
        memcpy(T, base16, sizeof(T));
        for (int i = np - 1; i >= 0; --i) {
            const int j = path[i];
            joint_x(ch->mov_offset + 12 * j, ch->mov_axis + 3 * j, ch->mov_type[j], q[b * dof + ch->mov_qidx[j]], X);
            mul12(T, X, T);
            memcpy(frames[i], T, sizeof(T));
        }
        mul12(T, ch->frame_tail + 12 * frame, T);
        if (use_com) mul12(T, ch->frame_com + 12 * frame, T);
        if (T_out) {
            double* o = T_out + b * 16;
            memcpy(o, T, sizeof(T));
            o[12] = o[13] = o[14] = 0.0; o[15] = 1.0;
        }
        if (J_out) {
            double* J = J_out + b * 6 * dof;
            memset(J, 0, sizeof(double) * 6 * dof);
            for (int i = np - 1; i >= 0; --i) {
                const int j = path[i];
                const double* F = frames[i];
                const double* a = ch->mov_axis + 3 * j;
                const double w[3] = {F[0] * a[0] + F[1] * a[1] + F[2] * a[2], F[4] * a[0] + F[5] * a[1] + F[6] * a[2],
                                     F[8] * a[0] + F[9] * a[1] + F[10] * a[2]};
                const int col = ch->mov_qidx[j];
                if (ch->mov_type[j] == JOINT_REVOLUTE) {
                    const double r[3] = {T[3] - F[3], T[7] - F[7], T[11] - F[11]};
                    J[0 * dof + col] = w[1] * r[2] - w[2] * r[1];
                    J[1 * dof + col] = w[2] * r[0] - w[0] * r[2];
                    J[2 * dof + col] = w[0] * r[1] - w[1] * r[0];
                    J[3 * dof + col] = w[0]; J[4 * dof + col] = w[1]; J[5 * dof + col] = w[2];
                } else {
                    J[0 * dof + col] = w[0]; J[1 * dof + col] = w[1]; J[2 * dof + col] = w[2];
                }
            }
        }
    }
    return 0;
}

/* world poses of every link primitive for one configuration */
static void prim_poses(const ochain* ch, const double* q, const double* base16, double* out /* n_prims x 12 */) {
    double jf[256][12];
    for (int j = 0; j < ch->n_mov; ++j) {
        double X[12];
        joint_x(ch->mov_offset + 12 * j, ch->mov_axis + 3 * j, ch->mov_type[j], q[ch->mov_qidx[j]], X);
        mul12(ch->mov_parent[j] < 0 ? base16 : jf[ch->mov_parent[j]], X, jf[j]);
    }
    for (int p = 0; p < ch->n_prims; ++p) {
        const int f = ch->prim_frame[p];
        const int j = ch->frame_mov[f];
        double L[12];
        mul12(j < 0 ? base16 : jf[j], ch->frame_tail + 12 * f, L);
        mul12(L, ch->prim_offset + 12 * p, out + 12 * p);
    }
}

static double config_min_distance(const ochain* ch, const oscene* sc, const double* q, const double* base16, double stop_below) {
    double P[64 * 12];
    prim_poses(ch, q, base16, P);
    double best = INFINITY;
    for (int e = 0; e < sc->n_env; ++e) {
        const int a = sc->env_pairs[2 * e], o = sc->env_pairs[2 * e + 1];
        const double d = pair_distance(ch->prim_type[a], P + 12 * a, ch->prim_dims + 3 * a, ch->prim_margin[a], sc->obs_type[o],
                                       sc->obs_pose + 12 * o, sc->obs_dims + 3 * o, sc->obs_margin[o], sc->env_sharp ? sc->env_sharp[e] : 0);
        if (d < best) best = d;
        if (best < stop_below) return best;
    }
    for (int e = 0; e < sc->n_self; ++e) {
        const int a = sc->self_pairs[2 * e], b = sc->self_pairs[2 * e + 1];
        const double d = pair_distance(ch->prim_type[a], P + 12 * a, ch->prim_dims + 3 * a, ch->prim_margin[a], ch->prim_type[b],
                                       P + 12 * b, ch->prim_dims + 3 * b, ch->prim_margin[b], sc->self_sharp ? sc->self_sharp[e] : 0);
        if (d < best) best = d;
        if (best < stop_below) return best;
    }
    return best;
}

/* verdict[b] = min distance < threshold; min_dist nullable (then early exit on the first hit) */
int oracle_in_collision(const ochain* ch, const oscene* sc, const double* q, int64_t n, const double* base16,
                        double threshold, uint8_t* verdict, double* min_dist) {
    if (ch->n_prims > 64 || ch->n_mov > 256) return -1;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t b = 0; b < n; ++b) {
        const double d = config_min_distance(ch, sc, q + b * ch->dof, base16, min_dist ? -INFINITY : threshold);
        verdict[b] = d < threshold;
        if (min_dist) min_dist[b] = d;
    }
    return 0;
}

/* connectors.py:57-100.  valid[e] = 1 iff the reference would return a state.  n_samples / n_useful nullable:
 * samples of the edge, and samples the sequential reference evaluates (up to and including the first invalid).
 * min_d nullable: the smallest min-distance over ALL samples of the edge (no early exit then; checker use: an edge
 * is "borderline" iff |min_d - threshold| is inside the contact band). */
int oracle_edges_valid_ex(const ochain* ch, const oscene* sc, const double* starts, const double* goals, int64_t E,
                          const double* base16, double resolution, double max_distance, double threshold, uint8_t* valid,
                          int32_t* n_samples, int32_t* n_useful, double* min_d) {
    if (ch->n_prims > 64 || ch->n_mov > 256 || ch->dof > 256) return -1;
    const int dof = ch->dof;
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t e = 0; e < E; ++e) {
        const double* s = starts + e * dof;
        const double* g = goals + e * dof;
        double ss = 0.0;
        for (int k = 0; k < dof; ++k) ss += (s[k] - g[k]) * (s[k] - g[k]);
        const double dist = sqrt(ss);
        int ok = 1, cnt = 0, used = 0;
        double md = INFINITY;
        if (dist <= 1.1920928955078125e-07) {
            ok = 0;
        } else {
            const double step = resolution / dist;
            double tf = 1.0;
            if (max_distance > 0.0 && dist > max_distance) tf = max_distance / dist;
            const int len = (int)ceil(tf / step);   /* numpy.arange length rule */
            cnt = len + 1;
            double qv[256];
            for (int k = 0; k <= len && (ok || min_d); ++k) {
                const double t = k == len ? tf : k * step;
                for (int c = 0; c < dof; ++c) qv[c] = (1.0 - t) * s[c] + t * g[c];
                if (ok) ++used;
                const double d = config_min_distance(ch, sc, qv, base16, min_d ? -INFINITY : threshold);
                if (d < md) md = d;
                if (d < threshold) ok = 0;
            }
        }
        valid[e] = (uint8_t)ok;
        if (n_samples) n_samples[e] = cnt;
        if (n_useful) n_useful[e] = used;
        if (min_d) min_d[e] = md;
    }
    return 0;
}

int oracle_edges_valid(const ochain* ch, const oscene* sc, const double* starts, const double* goals, int64_t E,
                       const double* base16, double resolution, double max_distance, double threshold, uint8_t* valid,
                       int32_t* n_samples, int32_t* n_useful) {
    return oracle_edges_valid_ex(ch, sc, starts, goals, E, base16, resolution, max_distance, threshold, valid, n_samples,
                                 n_useful, NULL);
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
