// Host build of numbotics_b200/csrc/nbx_gjk.cuh (the float32 device code: support functions, witness GJK, EPA) against the
// float64 oracle (oracle/nbx_oracle.c, linked as liboracle.so).  TEST INFRASTRUCTURE (tests/test_gjk_host_harness.py):
// the device header is compiled by g++ through tests/host_harness/fake/cuda_runtime.h, which maps the CUDA qualifiers away,
// so that the float32 penetration path can be exercised on random and on exactly symmetric pairs without a GPU.
// usage: epa_host <seed> <pairs>   -> "near-contact separated: N bad K worst E" (witness GJK, 0 < d < 3 mm) and
//                                    "tested N bad K worst E" (penetrating pairs; bad: |f32 - f64| > 1e-4 m)
#include <cstdio>
#include "nbx_gjk.cuh"
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
using namespace nbxg;
extern "C" int oracle_pair_distances(int64_t n, int n_link, const double* poses, int n_static, const double* sposes,
                          const int32_t* types, const double* dims, const double* margins, int n_pairs,
                          const int32_t* pairs, const uint8_t* sharp, double* out);
static void rot(std::mt19937& g, double* R) {
    std::normal_distribution<double> N(0, 1);
    double q[4]; double l = 0; for (int i = 0; i < 4; ++i) { q[i] = N(g); l += q[i]*q[i]; } l = sqrt(l); for (int i = 0; i < 4; ++i) q[i] /= l;
    double w=q[0],x=q[1],y=q[2],z=q[3];
    double M[9]={1-2*(y*y+z*z),2*(x*y-z*w),2*(x*z+y*w),2*(x*y+z*w),1-2*(x*x+z*z),2*(y*z-x*w),2*(x*z-y*w),2*(y*z+x*w),1-2*(x*x+y*y)};
    for (int i=0;i<9;++i) R[i]=M[i];
}
static void mkshape(ShapeW& S, int type, const double* R, const double* c, const double* d) {
    S.type = type; S.c = mk((float)c[0],(float)c[1],(float)c[2]);
    for (int k=0;k<9;++k) S.ax[k]=0.f;
    S.d[0]=(float)d[0]; S.d[1]=(float)d[1]; S.d[2]=(float)d[2];
    if (type==NBX_PRIM_CYLINDER) { S.ax[0]=(float)R[2]; S.ax[1]=(float)R[5]; S.ax[2]=(float)R[8]; }
    else if (type==NBX_PRIM_BOX) for (int k=0;k<9;++k) S.ax[k]=(float)R[k];
}
int main(int argc, char** argv) {
    int seed = argc > 1 ? atoi(argv[1]) : 1, N = argc > 2 ? atoi(argv[2]) : 20000;
    std::mt19937 g(seed);
    std::uniform_real_distribution<double> U(0, 1);
    int bad = 0, tested = 0; double worst = 0; int near_n = 0, near_bad = 0; double near_worst = 0;
    for (int it = 0; it < N; ++it) {
        int ta = (it % 3 == 0) ? NBX_PRIM_CYLINDER : (it % 3 == 1 ? NBX_PRIM_CYLINDER : NBX_PRIM_BOX);
        int tb = (it % 2 == 0) ? NBX_PRIM_BOX : NBX_PRIM_CYLINDER;
        if (it % 7 == 0) ta = NBX_PRIM_SPHERE;
        double Ra[9], Rb[9]; rot(g, Ra); rot(g, Rb);
        if (it % 5 == 0) { for (int i=0;i<9;++i) Rb[i] = (i%4==0); }   // axis-aligned obstacle
        double da[3], db[3];
        if (ta == NBX_PRIM_CYLINDER) { da[0]=0.03+0.03*U(g); da[1]=0.05+0.1*U(g); da[2]=0; } else if (ta==NBX_PRIM_BOX) { for(int k=0;k<3;++k) da[k]=0.03+0.1*U(g);} else {da[0]=da[1]=da[2]=0;}
        if (tb == NBX_PRIM_CYLINDER) { db[0]=0.03+0.3*U(g); db[1]=0.05+0.3*U(g); db[2]=0; } else { for(int k=0;k<3;++k) db[k]=0.05+0.45*U(g);}
        double ca[3], cb[3] = {1.35, 0.2, 0.5};
        double ext = std::max(db[0], std::max(db[1], db[2])) + 0.1;
        for (int k=0;k<3;++k) ca[k] = cb[k] + (2*U(g)-1) * ext;
        ShapeW A, B; mkshape(A, ta, Ra, ca, da); mkshape(B, tb, Rb, cb, db);
        // oracle: sharp=1 so that dims are the cores and margins 0 (spheres: radius 0 -> point)
        double sp[24]; 
        for (int r=0;r<3;++r){ for(int c=0;c<3;++c){ sp[r*4+c]=(double)(float)Ra[r*3+c]; sp[12+r*4+c]=(double)(float)Rb[r*3+c]; } sp[r*4+3]=(double)(float)ca[r]; sp[12+r*4+3]=(double)(float)cb[r]; }
        int32_t types[2]={ta,tb}; double dims[6]={(double)(float)da[0],(double)(float)da[1],(double)(float)da[2],(double)(float)db[0],(double)(float)db[1],(double)(float)db[2]};
        double margins[2]={0,0}; int32_t pairs[2]={0,1}; uint8_t sharp=1; double out; double dummy[12]={0};
        oracle_pair_distances(1,0,dummy,2,sp,types,dims,margins,1,pairs,&sharp,&out);
        if (out > -2e-5) {
            if (out < 3e-3) {   // near-contact separated pairs: the witness GJK itself
                const GjkWit gw = gjk_witness<96>(A, B);
                const double dg = gw.overlap ? 0.0 : gw.d;
                const double e2 = fabs(dg - out);
                ++near_n; if (e2 > near_worst) near_worst = e2; if (e2 > 1e-4) { ++near_bad; if (near_bad <= 5) printf("near: it %d types %d %d oracle %.6f gjk %.6f overlap %d\n", it, ta, tb, out, dg, (int)gw.overlap); }
            }
            continue;
        }
        ++tested;
        EpaOut e = core_penetration(A, B);
        double err = fabs(-e.depth - out);
        if (err > worst) worst = err;
        if (err > 1e-4) { ++bad; if (bad <= 10) printf("it %d types %d %d oracle %.6f epa %.6f ok %d\n", it, ta, tb, out, -e.depth, (int)e.ok); }
    }
    printf("near-contact separated: %d bad %d worst %.3e\n", near_n, near_bad, near_worst);
    printf("tested %d bad %d worst %.3e\n", tested, bad, worst);
}
