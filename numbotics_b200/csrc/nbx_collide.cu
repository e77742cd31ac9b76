// nbx_collide.cu -- batched Arm.in_collision and DiscreteConnector edge validation for sm_100a.
//
// Replaces reference numbotics/robots/arm.py:555-604 (collisions / closest_to / in_collision:
// one PyBullet state reset + getClosestPoints sweep per configuration) and
// numbotics/planning/sampling_based/connectors.py:57-100 (connect / steer: a Python loop of
// validity_checker calls along a linear trajectory).
//
// One thread evaluates one configuration: it walks the whole kinematic tree once (all link
// frames, not just one path), turns each link primitive into a world-space core as soon as its
// frame is known, and immediately tests it against the obstacles and the earlier primitives it
// is paired with.  Primitive poses that later pairs need are parked in a per-thread shared-memory
// scratch laid out [value][thread] (bank-conflict free).  The scene (tree, primitives,
// obstacles, pair lists; a few KB) is staged in shared memory once per CTA and read with
// warp-uniform addresses.  FP32-FMA bound; HBM traffic is 4*dof bytes in + 1 byte out.
//
// The edge kernel flattens the (edge, sample) space of a CTA's edges with a block prefix sum so
// every lane always has an interpolation sample to evaluate (no idle lanes on short edges) and
// edges that are already known invalid are skipped.
#include "nbx_common.cuh"
#include "nbx_gjk.cuh"
#include "nbx_joint.cuh"
#include <algorithm>
#include <mutex>

using namespace nbxg;

namespace {

constexpr int kCTile = 128;       // configurations (threads) per CTA, in_collision
constexpr int kETile = 128;       // edges per CTA, edge validation
constexpr int kMaxSave = 8;
constexpr int kGjkIters = 48;        // verdict mode: only near-threshold pairs ever get this far
constexpr int kGjkItersDist = 96;    // distance mode: run to convergence

struct SceneDev {
    const TreeJointK* tree;
    const PrimK* prims;
    const ObsK* obs;
    const PairK* self_pairs;
    const PairK* env_pairs;
    const int* self_begin;
    const int* env_begin;
    int n_mov, n_prims, n_base_prims, n_obs, n_self, n_env, scratch_floats;
    float base[12];
};

// shared-memory image of the scene
struct SceneSm {
    const TreeJointK* tree;
    const PrimK* prims;
    const ObsK* obs;
    const PairK* self_pairs;
    const PairK* env_pairs;
    const int* self_begin;
    const int* env_begin;
};

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

__host__ __device__ inline size_t scene_smem_bytes(int n_mov, int n_prims, int n_obs, int n_self, int n_env) {
    return align16(sizeof(TreeJointK) * n_mov) + align16(sizeof(PrimK) * n_prims) + align16(sizeof(ObsK) * n_obs) +
           align16(sizeof(PairK) * n_self) + align16(sizeof(PairK) * n_env) + 2 * align16(sizeof(int) * (n_prims + 1));
}

__device__ __forceinline__ void copy_words(void* dst, const void* src, size_t bytes) {
    const int n = (int)(bytes / 4);
    const int* s = static_cast<const int*>(src);
    int* d = static_cast<int*>(dst);
    for (int i = threadIdx.x; i < n; i += blockDim.x) d[i] = s[i];
}

__device__ __forceinline__ unsigned char* stage_scene(const SceneDev& S, unsigned char* sm, SceneSm& V) {
    unsigned char* p = sm;
    V.tree = reinterpret_cast<const TreeJointK*>(p);  copy_words(p, S.tree, sizeof(TreeJointK) * S.n_mov);  p += align16(sizeof(TreeJointK) * S.n_mov);
    V.prims = reinterpret_cast<const PrimK*>(p);      copy_words(p, S.prims, sizeof(PrimK) * S.n_prims);    p += align16(sizeof(PrimK) * S.n_prims);
    V.obs = reinterpret_cast<const ObsK*>(p);         copy_words(p, S.obs, sizeof(ObsK) * S.n_obs);         p += align16(sizeof(ObsK) * S.n_obs);
    V.self_pairs = reinterpret_cast<const PairK*>(p); copy_words(p, S.self_pairs, sizeof(PairK) * S.n_self); p += align16(sizeof(PairK) * S.n_self);
    V.env_pairs = reinterpret_cast<const PairK*>(p);  copy_words(p, S.env_pairs, sizeof(PairK) * S.n_env);   p += align16(sizeof(PairK) * S.n_env);
    V.self_begin = reinterpret_cast<const int*>(p);   copy_words(p, S.self_begin, sizeof(int) * (S.n_prims + 1)); p += align16(sizeof(int) * (S.n_prims + 1));
    V.env_begin = reinterpret_cast<const int*>(p);    copy_words(p, S.env_begin, sizeof(int) * (S.n_prims + 1));  p += align16(sizeof(int) * (S.n_prims + 1));
    return p;
}

using PoseF = nbxj::Pose<float>;

// world-space core of link primitive `pk` whose owning joint frame is X
__device__ __forceinline__ void prim_world(const PoseF& X, const PrimK& pk, ShapeW& S) {
    S.type = pk.type;
    S.c = mk(fmaf(X.R[0], pk.rel[3], fmaf(X.R[1], pk.rel[7], fmaf(X.R[2], pk.rel[11], X.p[0]))),
             fmaf(X.R[3], pk.rel[3], fmaf(X.R[4], pk.rel[7], fmaf(X.R[5], pk.rel[11], X.p[1]))),
             fmaf(X.R[6], pk.rel[3], fmaf(X.R[7], pk.rel[7], fmaf(X.R[8], pk.rel[11], X.p[2]))));
    S.d[0] = pk.dims[0]; S.d[1] = pk.dims[1]; S.d[2] = pk.dims[2];
    if (pk.type == NBX_PRIM_CYLINDER) {
        // only the local z axis matters
#pragma unroll
        for (int r = 0; r < 3; ++r)
            S.ax[r] = fmaf(X.R[r * 3 + 0], pk.rel[2], fmaf(X.R[r * 3 + 1], pk.rel[6], X.R[r * 3 + 2] * pk.rel[10]));
    } else if (pk.type == NBX_PRIM_BOX) {
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                S.ax[r * 3 + c] = fmaf(X.R[r * 3 + 0], pk.rel[c], fmaf(X.R[r * 3 + 1], pk.rel[4 + c], X.R[r * 3 + 2] * pk.rel[8 + c]));
    }
}

__device__ __forceinline__ void obs_world(const ObsK& o, ShapeW& S) {
    S.type = o.type;
    S.c = mk(o.c[0], o.c[1], o.c[2]);
    S.d[0] = o.dims[0]; S.d[1] = o.dims[1]; S.d[2] = o.dims[2];
    if (o.type == NBX_PRIM_CYLINDER) {
        S.ax[0] = o.R[2]; S.ax[1] = o.R[5]; S.ax[2] = o.R[8];
    } else if (o.type == NBX_PRIM_BOX) {
#pragma unroll
        for (int k = 0; k < 9; ++k) S.ax[k] = o.R[k];
    }
}

// sharp pairs (box-box under the Bullet margin model) put the margin back into the core
__device__ __forceinline__ void unround(ShapeW& S, float margin) {
    if (S.type == NBX_PRIM_BOX) { S.d[0] += margin; S.d[1] += margin; S.d[2] += margin; }
    else if (S.type == NBX_PRIM_CYLINDER) { S.d[0] += margin; S.d[1] += margin; }
}

__device__ __forceinline__ void scratch_store(float* scr, int stride, const ShapeW& S) {
    scr[0 * stride] = S.c.x; scr[1 * stride] = S.c.y; scr[2 * stride] = S.c.z;
    if (S.type == NBX_PRIM_CYLINDER) {
        scr[3 * stride] = S.ax[0]; scr[4 * stride] = S.ax[1]; scr[5 * stride] = S.ax[2];
    } else if (S.type == NBX_PRIM_BOX) {
#pragma unroll
        for (int k = 0; k < 9; ++k) scr[(3 + k) * stride] = S.ax[k];
    }
}

__device__ __forceinline__ void scratch_load(const float* scr, int stride, const PrimK& pk, ShapeW& S) {
    S.type = pk.type;
    S.c = mk(scr[0 * stride], scr[1 * stride], scr[2 * stride]);
    S.d[0] = pk.dims[0]; S.d[1] = pk.dims[1]; S.d[2] = pk.dims[2];
    if (pk.type == NBX_PRIM_CYLINDER) {
        S.ax[0] = scr[3 * stride]; S.ax[1] = scr[4 * stride]; S.ax[2] = scr[5 * stride];
    } else if (pk.type == NBX_PRIM_BOX) {
#pragma unroll
        for (int k = 0; k < 9; ++k) S.ax[k] = scr[(3 + k) * stride];
    }
}

// Distance of one pair minus margins, with verdict-style early outs:
//   returns a value d with  (d < thr) == (true pair distance < thr)  whenever the true distance is
//   outside the convergence band; when WANT_DIST the value itself is the pair distance
//   (lower-bounded by `best` pruning).
template <bool WANT_DIST>
__device__ __forceinline__ float pair_distance(ShapeW& A, float mA, ShapeW& B, float mB, bool sharp, float thr, float best) {
    float m = mA + mB;
    const float bd0 = B.d[0], bd1 = B.d[1], bd2 = B.d[2];
    if (sharp) {
        unround(A, mA);
        unround(B, mB);
        // spheres keep their radius (their core is a point)
        m = (A.type == NBX_PRIM_SPHERE ? mA : 0.f) + (B.type == NBX_PRIM_SPHERE ? mB : 0.f);
    }
    GjkOut g;
    if (WANT_DIST) g = gjk_distance<kGjkItersDist>(A, B, best + m, -FLT_MAX);
    else g = gjk_distance<kGjkIters>(A, B, thr + m, thr + m);
    B.d[0] = bd0; B.d[1] = bd1; B.d[2] = bd2;   // B is reused by the caller's next pair
    if (g.overlap) return -m - 1e-3f;   // cores overlap: depth beyond the margins is not computed (clamp)
    return g.d - m;
}

// Evaluate one configuration.  QF: float operator()(int col) returns q[col].
// Returns the minimum distance (WANT_DIST) or any value < thr as soon as a pair is in contact.
template <bool WANT_DIST, typename QF>
__device__ float eval_config(const SceneDev& S, const SceneSm& V, QF qf, float thr, float* scratch, int stride) {
    float best = FLT_MAX;
    PoseF X;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        X.R[r * 3 + 0] = S.base[r * 4 + 0]; X.R[r * 3 + 1] = S.base[r * 4 + 1]; X.R[r * 3 + 2] = S.base[r * 4 + 2];
        X.p[r] = S.base[r * 4 + 3];
    }
    PoseF saved[kMaxSave];

    auto do_prims = [&](int p0, int p1) -> bool {
        for (int p = p0; p < p1; ++p) {
            const PrimK& pk = V.prims[p];
            ShapeW Bs;
            prim_world(X, pk, Bs);
            for (int e = V.env_begin[p]; e < V.env_begin[p + 1]; ++e) {
                const PairK pr = V.env_pairs[e];
                const ObsK& o = V.obs[pr.b];
                // conservative cull: primitive bounding sphere vs the obstacle (exact point-box distance
                // for boxes, bounding sphere otherwise); (distance lower bound) > cut can neither hit nor
                // lower the running minimum
                {
                    const V3 dc = Bs.c - mk(o.c[0], o.c[1], o.c[2]);
                    float lbnd;
                    if (o.type == NBX_PRIM_BOX) {
                        const float hx = o.dims[0] + o.margin, hy = o.dims[1] + o.margin, hz = o.dims[2] + o.margin;
                        const float lx = fmaxf(fabsf(fmaf(o.R[0], dc.x, fmaf(o.R[3], dc.y, o.R[6] * dc.z))) - hx, 0.f);
                        const float ly = fmaxf(fabsf(fmaf(o.R[1], dc.x, fmaf(o.R[4], dc.y, o.R[7] * dc.z))) - hy, 0.f);
                        const float lz = fmaxf(fabsf(fmaf(o.R[2], dc.x, fmaf(o.R[5], dc.y, o.R[8] * dc.z))) - hz, 0.f);
                        lbnd = sqrtf(fmaf(lx, lx, fmaf(ly, ly, lz * lz))) - pk.brad;
                    } else {
                        lbnd = sqrtf(dot(dc, dc)) - pk.brad - o.brad;
                    }
                    if (lbnd > (WANT_DIST ? best : thr) + 1e-6f) continue;
                }
                ShapeW As;
                obs_world(o, As);
                const float d = pair_distance<WANT_DIST>(As, o.margin, Bs, pk.margin, pr.sharp != 0, thr, best);
                best = fminf(best, d);
                if (!WANT_DIST && d < thr) return true;
            }
            for (int e = V.self_begin[p]; e < V.self_begin[p + 1]; ++e) {
                const PairK pr = V.self_pairs[e];
                const PrimK& pa = V.prims[pr.a];
                {
                    const float* sc = scratch + pa.slot * stride;
                    const V3 dc = Bs.c - mk(sc[0], sc[stride], sc[2 * stride]);
                    if (sqrtf(dot(dc, dc)) - pk.brad - pa.brad > (WANT_DIST ? best : thr) + 1e-6f) continue;
                }
                ShapeW As;
                scratch_load(scratch + pa.slot * stride, stride, pa, As);
                const float d = pair_distance<WANT_DIST>(As, pa.margin, Bs, pk.margin, pr.sharp != 0, thr, best);
                best = fminf(best, d);
                if (!WANT_DIST && d < thr) return true;
            }
            if (pk.slot >= 0) scratch_store(scratch + pk.slot * stride, stride, Bs);
        }
        return false;
    };

    if (do_prims(0, S.n_base_prims)) return best;
#pragma unroll 1
    for (int j = 0; j < S.n_mov; ++j) {
        const TreeJointK& jk = V.tree[j];
        if (jk.load_slot == -2) {
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                X.R[r * 3 + 0] = S.base[r * 4 + 0]; X.R[r * 3 + 1] = S.base[r * 4 + 1]; X.R[r * 3 + 2] = S.base[r * 4 + 2];
                X.p[r] = S.base[r * 4 + 3];
            }
        } else if (jk.load_slot >= 0) {
            X = saved[jk.load_slot];
        }
        nbxj::joint_step<float, TreeJointK>(X, jk, qf(jk.qidx));
        if (jk.save_slot >= 0) saved[jk.save_slot] = X;
        if (do_prims(jk.prim_begin, jk.prim_end)) return best;
    }
    return best;
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCTile)
in_collision_kernel(const __grid_constant__ SceneDev S, const float* __restrict__ q, int64_t B, int dof, float thr,
                    uint8_t* __restrict__ verdict, float* __restrict__ min_dist) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SceneSm V;
    unsigned char* p = stage_scene(S, smem_raw, V);
    float* sq = reinterpret_cast<float*>(p);
    float* scratch = sq + (size_t)kCTile * dof;

    const int64_t tile0 = (int64_t)blockIdx.x * kCTile;
    const int nvalid = (int)min((int64_t)kCTile, B - tile0);
    const int tid = threadIdx.x;
    {
        const float* src = q + tile0 * dof;
        const int n = nvalid * dof;
        if (nvalid == kCTile && ((reinterpret_cast<uintptr_t>(q) & 15u) == 0)) {
            const int n16 = n / 4;
            const float4* s16 = reinterpret_cast<const float4*>(src);
            float4* d16 = reinterpret_cast<float4*>(sq);
            for (int i = tid; i < n16; i += kCTile) d16[i] = __ldg(s16 + i);
        } else {
            for (int i = tid; i < n; i += kCTile) sq[i] = src[i];
        }
    }
    __syncthreads();
    if (tid >= nvalid) return;
    const float* myq = sq + tid * dof;
    auto qf = [myq](int col) { return myq[col]; };
    float d;
    if (min_dist != nullptr) d = eval_config<true>(S, V, qf, thr, scratch + tid, kCTile);
    else d = eval_config<false>(S, V, qf, thr, scratch + tid, kCTile);
    verdict[tile0 + tid] = d < thr ? 1 : 0;
    if (min_dist != nullptr) min_dist[tile0 + tid] = d;
}

// ---------------------------------------------------------------------------------------------
// Edge validation.  Per edge (all float64, bit-for-bit the reference's NumPy arithmetic,
// connectors.py:62-69 / :84-93):
//   distance = ||goal - start||_2 ; distance <= float32 eps -> invalid
//   step = resolution / distance ; T_f = 1 (connect) or min(1, max_distance / distance) (steer)
//   samples: t_k = k * step for k < ceil(T_f / step)  (numpy.arange length rule), then t = T_f.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kETile)
edges_valid_kernel(const __grid_constant__ SceneDev S, const float* __restrict__ starts, const float* __restrict__ goals,
                   int64_t E, int dof, double resolution, double max_distance, float thr,
                   uint8_t* __restrict__ valid, int32_t* __restrict__ n_samples) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SceneSm V;
    unsigned char* p = stage_scene(S, smem_raw, V);
    float* s_start = reinterpret_cast<float*>(p);
    float* s_goal = s_start + (size_t)kETile * dof;
    double* s_step = reinterpret_cast<double*>(s_goal + (size_t)kETile * dof + ((kETile * dof) & 1));
    double* s_tf = s_step + kETile;
    int* s_off = reinterpret_cast<int*>(s_tf + kETile);      // [kETile + 1] exclusive prefix of sample counts
    volatile int* s_ok = s_off + kETile + 1;                 // [kETile]
    float* scratch = reinterpret_cast<float*>(s_off + 2 * kETile + 1);

    const int64_t tile0 = (int64_t)blockIdx.x * kETile;
    const int nvalid = (int)min((int64_t)kETile, E - tile0);
    const int tid = threadIdx.x;
    for (int i = tid; i < nvalid * dof; i += kETile) {
        s_start[i] = starts[tile0 * dof + i];
        s_goal[i] = goals[tile0 * dof + i];
    }
    __syncthreads();

    int cnt = 0;
    if (tid < nvalid) {
        double ss = 0.0;
        for (int k = 0; k < dof; ++k) {
            const double dlt = (double)s_start[tid * dof + k] - (double)s_goal[tid * dof + k];
            ss = fma(dlt, dlt, ss);
        }
        const double dist = sqrt(ss);
        int ok = 1;
        if (dist <= (double)1.1920928955078125e-07f) {
            ok = 0;
        } else {
            const double step = resolution / dist;
            double tf = 1.0;
            if (max_distance > 0.0 && dist > max_distance) tf = max_distance / dist;
            const double len = ceil(tf / step);
            cnt = (int)fmin(len, 1.0e9) + 1;
            s_step[tid] = step;
            s_tf[tid] = tf;
        }
        s_ok[tid] = ok;
        if (n_samples != nullptr) n_samples[tile0 + tid] = cnt;
    }
    // block exclusive scan of cnt (kETile = 128: warp scans + 4 partials)
    {
        __shared__ int warp_tot[kETile / 32];
        int x = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if ((tid & 31) >= o) x += y;
        }
        if ((tid & 31) == 31) warp_tot[tid >> 5] = x;
        __syncthreads();
        int base = 0;
        for (int w = 0; w < (tid >> 5); ++w) base += warp_tot[w];
        s_off[tid] = base + x - cnt;
        if (tid == kETile - 1) s_off[kETile] = base + x;
    }
    __syncthreads();
    const int total = s_off[kETile];

    for (int s = tid; s < total; s += kETile) {
        // owning edge: last e with s_off[e] <= s
        int lo = 0, hi = kETile - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (s_off[mid] <= s) lo = mid; else hi = mid - 1;
        }
        const int e = lo;
        if (s_ok[e] == 0) continue;   // already invalid (the sequential reference stops there too)
        const int k = s - s_off[e];
        const int last = s_off[e + 1] - s_off[e] - 1;
        const double td = (k == last) ? s_tf[e] : (double)k * s_step[e];
        const float t = (float)td;
        const float* a = s_start + e * dof;
        const float* g = s_goal + e * dof;
        auto qf = [a, g, t](int col) { return fmaf(t, g[col], fmaf(-t, a[col], a[col])); };
        const float d = eval_config<false>(S, V, qf, thr, scratch + tid, kETile);
        if (d < thr) s_ok[e] = 0;
    }
    __syncthreads();
    if (tid < nvalid) valid[tile0 + tid] = (uint8_t)s_ok[tid];
}

// ---------------------------------------------------------------------------------------------
int fill_scene_dev(const nbx_chain* ch, const nbx_scene* sc, const double* base_pose, SceneDev& S) {
    NBX_REQUIRE(ch != nullptr && sc != nullptr, "NULL handle");
    NBX_REQUIRE(sc->chain == ch, "scene was created for a different chain");
    NBX_REQUIRE(base_pose != nullptr, "base_pose is NULL");
    S.tree = ch->d_tree;
    S.prims = sc->d_prims;
    S.obs = sc->d_obs;
    S.self_pairs = sc->d_self;
    S.env_pairs = sc->d_env;
    S.self_begin = sc->d_self_begin;
    S.env_begin = sc->d_env_begin;
    S.n_mov = ch->n_mov;
    S.n_prims = ch->n_prims;
    S.n_base_prims = ch->n_base_prims;
    S.n_obs = sc->n_obs;
    S.n_self = sc->n_self;
    S.n_env = sc->n_env;
    S.scratch_floats = sc->scratch_floats;
    for (int k = 0; k < 12; ++k) S.base[k] = (float)base_pose[k];
    return NBX_OK;
}

int launch_in_collision(const nbx_chain* ch, const nbx_scene* sc, const float* q, int64_t B, const double* base_pose,
                        float thr, uint8_t* verdict, float* min_dist, cudaStream_t stream) {
    SceneDev S;
    int rc = fill_scene_dev(ch, sc, base_pose, S);
    if (rc != NBX_OK) return rc;
    NBX_REQUIRE(B >= 0, "negative batch size");
    if (B == 0) return NBX_OK;
    NBX_REQUIRE(q != nullptr && verdict != nullptr, "q / verdict is NULL");
    NBX_REQUIRE(thr >= 0.f, "negative thresholds need penetration depth, which this path does not compute");
    NBX_CUDA_OK(cudaSetDevice(ch->device));
    const size_t smem = scene_smem_bytes(S.n_mov, S.n_prims, S.n_obs, S.n_self, S.n_env) +
                        sizeof(float) * (size_t)kCTile * ch->dof + sizeof(float) * (size_t)kCTile * std::max(1, S.scratch_floats);
    NBX_REQUIRE(smem <= 220 * 1024, "scene too large for shared memory (%zu bytes)", smem);
    NBX_CUDA_OK(cudaFuncSetAttribute(in_collision_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t grid = (B + kCTile - 1) / kCTile;
    NBX_REQUIRE(grid <= 0x7fffffffLL, "batch too large for one launch");
    in_collision_kernel<<<(unsigned)grid, kCTile, smem, stream>>>(S, q, B, ch->dof, thr, verdict, min_dist);
    nbx_count_launch();
    NBX_CUDA_OK(cudaGetLastError());
    return NBX_OK;
}

int launch_edges(const nbx_chain* ch, const nbx_scene* sc, const float* starts, const float* goals, int64_t E,
                 const double* base_pose, double resolution, double max_distance, float thr, uint8_t* valid,
                 int32_t* n_samples, cudaStream_t stream) {
    SceneDev S;
    int rc = fill_scene_dev(ch, sc, base_pose, S);
    if (rc != NBX_OK) return rc;
    NBX_REQUIRE(E >= 0, "negative edge count");
    if (E == 0) return NBX_OK;
    NBX_REQUIRE(starts != nullptr && goals != nullptr && valid != nullptr, "starts / goals / valid is NULL");
    NBX_REQUIRE(resolution > 0.0, "Resolution must be positive");
    NBX_REQUIRE(thr >= 0.f, "negative thresholds need penetration depth, which this path does not compute");
    NBX_CUDA_OK(cudaSetDevice(ch->device));
    const int dof = ch->dof;
    const size_t smem = scene_smem_bytes(S.n_mov, S.n_prims, S.n_obs, S.n_self, S.n_env) +
                        sizeof(float) * (2 * (size_t)kETile * dof + ((kETile * dof) & 1)) + sizeof(double) * 2 * kETile +
                        sizeof(int) * (2 * kETile + 1) + 16 + sizeof(float) * (size_t)kETile * std::max(1, S.scratch_floats);
    NBX_REQUIRE(smem <= 220 * 1024, "scene too large for shared memory (%zu bytes)", smem);
    NBX_CUDA_OK(cudaFuncSetAttribute(edges_valid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t grid = (E + kETile - 1) / kETile;
    NBX_REQUIRE(grid <= 0x7fffffffLL, "too many edges for one launch");
    edges_valid_kernel<<<(unsigned)grid, kETile, smem, stream>>>(S, starts, goals, E, dof, resolution, max_distance, thr, valid, n_samples);
    nbx_count_launch();
    NBX_CUDA_OK(cudaGetLastError());
    return NBX_OK;
}

// host-buffer variants: one staging allocation per call is acceptable here (verdict bytes are tiny,
// the copies are the dominant term); stream-ordered allocation keeps it asynchronous.
struct DevTmp {
    void* p = nullptr;
    ~DevTmp() { if (p) cudaFree(p); }
};

}  // namespace

extern "C" {

int nbx_in_collision_f32(const nbx_chain* chain, const nbx_scene* scene, const float* q, int64_t B, const double* base_pose,
                         float threshold, uint8_t* verdict, float* min_dist, void* stream) {
    return launch_in_collision(chain, scene, q, B, base_pose, threshold, verdict, min_dist, static_cast<cudaStream_t>(stream));
}

int nbx_in_collision_host_f32(const nbx_chain* chain, const nbx_scene* scene, const float* q, int64_t B, const double* base_pose,
                              float threshold, uint8_t* verdict) {
    NBX_REQUIRE(chain != nullptr, "chain handle is NULL");
    if (B <= 0) return NBX_OK;
    NBX_CUDA_OK(cudaSetDevice(chain->device));
    DevTmp dq, dv;
    const size_t qb = sizeof(float) * (size_t)B * chain->dof;
    NBX_CUDA_OK(cudaMalloc(&dq.p, qb));
    NBX_CUDA_OK(cudaMalloc(&dv.p, (size_t)B));
    cudaStream_t st = nullptr;
    NBX_CUDA_OK(cudaMemcpyAsync(dq.p, q, qb, cudaMemcpyHostToDevice, st));
    int rc = launch_in_collision(chain, scene, static_cast<const float*>(dq.p), B, base_pose, threshold,
                                 static_cast<uint8_t*>(dv.p), nullptr, st);
    if (rc != NBX_OK) return rc;
    NBX_CUDA_OK(cudaMemcpyAsync(verdict, dv.p, (size_t)B, cudaMemcpyDeviceToHost, st));
    NBX_CUDA_OK(cudaStreamSynchronize(st));
    return NBX_OK;
}

int nbx_edges_valid_f32(const nbx_chain* chain, const nbx_scene* scene, const float* starts, const float* goals, int64_t E,
                        const double* base_pose, double resolution, double max_distance, float threshold, uint8_t* valid,
                        int32_t* n_samples, void* stream) {
    return launch_edges(chain, scene, starts, goals, E, base_pose, resolution, max_distance, threshold, valid, n_samples,
                        static_cast<cudaStream_t>(stream));
}

int nbx_edges_valid_host_f32(const nbx_chain* chain, const nbx_scene* scene, const float* starts, const float* goals, int64_t E,
                             const double* base_pose, double resolution, double max_distance, float threshold, uint8_t* valid) {
    NBX_REQUIRE(chain != nullptr, "chain handle is NULL");
    if (E <= 0) return NBX_OK;
    NBX_CUDA_OK(cudaSetDevice(chain->device));
    DevTmp ds, dg, dv;
    const size_t qb = sizeof(float) * (size_t)E * chain->dof;
    NBX_CUDA_OK(cudaMalloc(&ds.p, qb));
    NBX_CUDA_OK(cudaMalloc(&dg.p, qb));
    NBX_CUDA_OK(cudaMalloc(&dv.p, (size_t)E));
    cudaStream_t st = nullptr;
    NBX_CUDA_OK(cudaMemcpyAsync(ds.p, starts, qb, cudaMemcpyHostToDevice, st));
    NBX_CUDA_OK(cudaMemcpyAsync(dg.p, goals, qb, cudaMemcpyHostToDevice, st));
    int rc = launch_edges(chain, scene, static_cast<const float*>(ds.p), static_cast<const float*>(dg.p), E, base_pose,
                          resolution, max_distance, threshold, static_cast<uint8_t*>(dv.p), nullptr, st);
    if (rc != NBX_OK) return rc;
    NBX_CUDA_OK(cudaMemcpyAsync(valid, dv.p, (size_t)E, cudaMemcpyDeviceToHost, st));
    NBX_CUDA_OK(cudaStreamSynchronize(st));
    return NBX_OK;
}

}  // extern "C"
