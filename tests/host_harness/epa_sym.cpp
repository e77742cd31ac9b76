// Exactly symmetric poses (centres on a common line, axis-aligned shapes): the cases that degenerate the GJK seed and put the
// minimum of the support function on a whole circle of normals.  Same host build of the device header as epa_host.cpp.
#include <cstdio>
#include "nbx_gjk.cuh"
#include <cstdlib>
using namespace nbxg;
extern "C" int oracle_pair_distances(int64_t n, int n_link, const double* poses, int n_static, const double* sposes,
                          const int32_t* types, const double* dims, const double* margins, int n_pairs,
                          const int32_t* pairs, const uint8_t* sharp, double* out);
static void mkshape(ShapeW& S, int type, const double* c, const double* d) {
    S.type = type; S.c = mk((float)c[0],(float)c[1],(float)c[2]);
    for (int k=0;k<9;++k) S.ax[k]=0.f;
    S.d[0]=(float)d[0]; S.d[1]=(float)d[1]; S.d[2]=(float)d[2];
    if (type==NBX_PRIM_CYLINDER) { S.ax[2]=1.f; }
    else if (type==NBX_PRIM_BOX) { S.ax[0]=S.ax[4]=S.ax[8]=1.f; }
}
int main() {
    double dimsA[3][3] = {{0,0,0},{0.06,0.09,0.12},{0.05,0.15,0}};
    double dimsB[3][3] = {{0,0,0},{0.3,0.2,0.25},{0.25,0.25,0}};
    double worst = 0;
    for (int ta = 0; ta < 3; ++ta) for (int tb = 1; tb < 3; ++tb) for (int off = 0; off < 2; ++off) {
        double w2 = 0;
        for (int i = 0; i <= 240; ++i) {
            double x = -0.6 + 0.005 * i;
            double ca[3] = {x, 0, 0}, cb[3] = {0, off ? 0.013 : 0.0, off ? -0.021 : 0.0};
            ShapeW A, B; mkshape(A, ta, ca, dimsA[ta]); mkshape(B, tb, cb, dimsB[tb]);
            double sp[24] = {1,0,0,(double)(float)ca[0], 0,1,0,(double)(float)ca[1], 0,0,1,(double)(float)ca[2], 1,0,0,(double)(float)cb[0], 0,1,0,(double)(float)cb[1], 0,0,1,(double)(float)cb[2]};
            int32_t types[2]={ta,tb}; double dims[6]; for (int k=0;k<3;++k){dims[k]=(double)(float)dimsA[ta][k]; dims[3+k]=(double)(float)dimsB[tb][k];}
            double margins[2]={0,0}; int32_t pairs[2]={0,1}; uint8_t sharp=1; double out, dummy[12]={0};
            oracle_pair_distances(1,0,dummy,2,sp,types,dims,margins,1,pairs,&sharp,&out);
            if (out > -2e-5) continue;
            EpaOut e = core_penetration(A, B);
            double err = fabs(-e.depth - out);
            if (err > w2) w2 = err;
            if (err > 2e-4) printf("ta %d tb %d off %d x %.3f oracle %.6f epa %.6f\n", ta, tb, off, x, out, -e.depth);
        }
        printf("ta %d tb %d off %d worst %.3e\n", ta, tb, off, w2);
        if (w2 > worst) worst = w2;
    }
    printf("worst %.3e\n", worst);
}
