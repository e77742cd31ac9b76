// Host build of the float32 device header as a tiny C library (tests only): one pair at a time.
#include <cstdio>
#include "nbx_gjk.cuh"
using namespace nbxg;
static void mkshape(ShapeW& S, int type, const double* pose12, const double* core_dims) {
    S.type = type; S.c = mk((float)pose12[3], (float)pose12[7], (float)pose12[11]);
    for (int k = 0; k < 9; ++k) S.ax[k] = 0.f;
    for (int k = 0; k < 3; ++k) S.d[k] = (float)core_dims[k];
    if (type == NBX_PRIM_CYLINDER) { S.ax[0] = (float)pose12[2]; S.ax[1] = (float)pose12[6]; S.ax[2] = (float)pose12[10]; }
    else if (type == NBX_PRIM_BOX) for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) S.ax[r * 3 + c] = (float)pose12[r * 4 + c];
}
// signed core distance: > 0 separated (witness GJK), < 0 penetration depth (EPA); out = {d, n[3], pa[3], pb[3]}
extern "C" void host_core_signed_distance(int ta, const double* pa12, const double* da, int tb, const double* pb12, const double* db, double* out7) {
    ShapeW A, B; mkshape(A, ta, pa12, da); mkshape(B, tb, pb12, db);
    const GjkWit g = gjk_witness<96>(A, B);
    if (!g.overlap) {
        out7[0] = g.d;
        const V3 n = (1.f / g.d) * (g.pa - g.pb);
        out7[1] = n.x; out7[2] = n.y; out7[3] = n.z;
        out7[4] = g.pa.x; out7[5] = g.pa.y; out7[6] = g.pa.z; out7[7] = g.pb.x; out7[8] = g.pb.y; out7[9] = g.pb.z;
        return;
    }
    const EpaOut e = core_penetration(A, B);
    out7[0] = -e.depth; out7[1] = e.n.x; out7[2] = e.n.y; out7[3] = e.n.z;
    out7[4] = e.pa.x; out7[5] = e.pa.y; out7[6] = e.pa.z; out7[7] = e.pb.x; out7[8] = e.pb.y; out7[9] = e.pb.z;
}
