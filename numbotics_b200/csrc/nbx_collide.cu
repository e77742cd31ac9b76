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
           align16(sizeof(PairK) * n_self) + align16(sizeof(PairK) * n_env) + align16(sizeof(int) * (n_obs + 1));
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
    V.env_begin = reinterpret_cast<const int*>(p);    copy_words(p, S.env_begin, sizeof(int) * (S.n_obs + 1));    p += align16(sizeof(int) * (S.n_obs + 1));
    V.self_begin = nullptr;
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

// ---------------------------------------------------------------------------------------------
// Block-cooperative evaluation of NT items (configurations or interpolation samples), two phases:
//
//   phase 1  thread = item.  Walk the whole tree, park EVERY primitive's world core in shared memory
//            ([value][thread], conflict free), and run the cheap conservative culls on every allowed
//            pair.  All lanes walk the same scene in lockstep, so survivors are appended to a block
//            queue with one warp-aggregated shared-memory atomic per (pair, warp).
//   phase 2  thread = queue entry.  Surviving (pair, item) candidates -- typically a handful per
//            item out of ~60 pairs -- go through the GJK narrow phase.  Entries are ordered by pair,
//            so a warp works on one pair type at a time and lanes only diverge in iteration count.
//
// The per-thread version of this (cull and GJK interleaved in one loop) ran at 72 % SIMT efficiency
// and 229 KB of inlined code; see profiles/.
// ---------------------------------------------------------------------------------------------
constexpr int kQueuePerThread = 20;
constexpr uint32_t kEnvFlag = 0x80000000u;

struct Ctx {
    SceneSm V;
    const float* base;        // 12 floats (kernel parameter)
    float* poses;             // [pose_floats][NT]
    uint32_t* queue;          // [qcap]
    int* qcount;
    int* flag;                // verdict mode: [n_flags] 1 = hit ; distance mode: ordered-int keys of the running minimum
    const int* item_flag;     // [NT] item -> flag index (nullptr = identity)
    int qcap;
    int n_mov, n_prims, n_base_prims;
    float thr;
};

__device__ __forceinline__ int float_key(float f) {            // order-preserving float -> int
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float key_float(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

__device__ __forceinline__ void pose_store(float* scr, int stride, const ShapeW& S) {
    scr[0 * stride] = S.c.x; scr[1 * stride] = S.c.y; scr[2 * stride] = S.c.z;
    if (S.type == NBX_PRIM_CYLINDER) {
        scr[3 * stride] = S.ax[0]; scr[4 * stride] = S.ax[1]; scr[5 * stride] = S.ax[2];
    } else if (S.type == NBX_PRIM_BOX) {
#pragma unroll
        for (int k = 0; k < 9; ++k) scr[(3 + k) * stride] = S.ax[k];
    }
}

__device__ __forceinline__ void pose_load(const float* scr, int stride, const PrimK& pk, ShapeW& S) {
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

// Narrow phase of one queued candidate.  The ONLY place the GJK is instantiated (per mode).
template <bool WANT_DIST, int NT>
__device__ __forceinline__ void eval_candidate(const Ctx& C, uint32_t entry) {
    const int item = (int)(entry & 0xffu);
    const int e = (int)((entry & ~kEnvFlag) >> 8);
    const int fi = C.item_flag ? C.item_flag[item] : item;
    float best = FLT_MAX;
    if (WANT_DIST) best = key_float(C.flag[fi]);
    else if (((volatile int*)C.flag)[fi]) return;          // this item is already decided
    ShapeW A, Bs;
    float mA, mB;
    bool sharp;
    if (entry & kEnvFlag) {
        const PairK pr = C.V.env_pairs[e];
        const PrimK& pb = C.V.prims[pr.a];
        const ObsK& o = C.V.obs[pr.b];
        obs_world(o, A);
        pose_load(C.poses + (size_t)pb.slot * NT + item, NT, pb, Bs);
        mA = o.margin; mB = pb.margin; sharp = pr.sharp != 0;
    } else {
        const PairK pr = C.V.self_pairs[e];
        const PrimK& pa = C.V.prims[pr.a];
        const PrimK& pb = C.V.prims[pr.b];
        pose_load(C.poses + (size_t)pa.slot * NT + item, NT, pa, A);
        pose_load(C.poses + (size_t)pb.slot * NT + item, NT, pb, Bs);
        mA = pa.margin; mB = pb.margin; sharp = pr.sharp != 0;
    }
    float m = mA + mB;
    if (sharp) {
        // box-box under the Bullet margin model: sharp boxes (spheres keep their radius: point cores)
        unround(A, mA);
        unround(Bs, mB);
        m = (A.type == NBX_PRIM_SPHERE ? mA : 0.f) + (Bs.type == NBX_PRIM_SPHERE ? mB : 0.f);
    }
    GjkOut g;
    if (WANT_DIST) g = gjk_distance<kGjkItersDist>(A, Bs, best + m, -FLT_MAX);
    else g = gjk_distance<kGjkIters>(A, Bs, C.thr + m, C.thr + m);
    // overlapping cores: depth beyond the margins is not computed (clamp)
    const float d = g.overlap ? (-m - 1e-3f) : (g.contact ? (-m - 1e-6f) : (g.d - m));
    if (WANT_DIST) {
        if (!g.pruned) atomicMin(&C.flag[fi], float_key(d));
    } else if (d < C.thr) {
        ((volatile int*)C.flag)[fi] = 1;
    }
}

// cold copy for the queue-overflow path, kept out of line so phase 1 stays small
template <bool WANT_DIST, int NT>
__device__ __noinline__ void eval_candidate_cold(const Ctx& C, uint32_t entry) { eval_candidate<WANT_DIST, NT>(C, entry); }

// distance between segments {ca + s ua, |s| <= ha} and {cb + t ub, |t| <= hb} (unit directions)
__device__ __forceinline__ float seg_seg_distance(V3 ca, V3 ua, float ha, V3 cb, V3 ub, float hb) {
    const V3 r = ca - cb;
    const float b = dot(ua, ub), c = dot(ua, r), f = dot(ub, r);
    const float den = 1.f - b * b;
    float s = den > 1e-9f ? fminf(fmaxf((b * f - c) / den, -ha), ha) : 0.f;
    float t = fminf(fmaxf(fmaf(b, s, f), -hb), hb);
    s = fminf(fmaxf(fmaf(b, t, -c), -ha), ha);
    t = fminf(fmaxf(fmaf(b, s, f), -hb), hb);      // one more alternation: exact for clamped cases
    const V3 d = madd(madd(r, s, ua), -t, ub);
    return sqrtf(dot(d, d));
}

// warp-aggregated append of the surviving lanes' candidates; lanes that do not fit evaluate inline
template <bool WANT_DIST, int NT>
__device__ __forceinline__ void push_candidates(const Ctx& C, bool surv, uint32_t code) {
    const unsigned mask = __ballot_sync(0xffffffffu, surv);
    if (mask == 0) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    int pos = 0;
    if (lane == leader) pos = atomicAdd(C.qcount, __popc(mask));
    pos = __shfl_sync(0xffffffffu, pos, leader) + __popc(mask & ((1u << lane) - 1u));
    if (surv) {
        const uint32_t entry = code | (uint32_t)threadIdx.x;
        if (pos < C.qcap) C.queue[pos] = entry;
        else eval_candidate_cold<WANT_DIST, NT>(C, entry);     // queue full (pathological scenes): no loss, just slower
    }
}

// phase 1 for this thread's item.  `valid` = the thread has an item; invalid lanes still take part in the ballots.
//   1a  walk the tree, park every primitive's world core in shared memory;
//   1b  broad phase.  Obstacle-major for the environment pairs (one obstacle's frame stays in registers
//       while the primitives paired with it stream past: exact point-to-box distance minus the primitive's
//       bounding radius), then the self pairs (bounding spheres, refined by the capsule bound for
//       cylinder-cylinder pairs that survive).  Every pair record is one 16-byte shared-memory load.
template <bool WANT_DIST, int NT, typename QF>
__device__ __forceinline__ void phase1(const Ctx& C, bool valid, QF qf, int n_obs, int n_self) {
    const int tid = threadIdx.x;
    float* my = C.poses + tid;
    {
        PoseF X;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            X.R[r * 3 + 0] = C.base[r * 4 + 0]; X.R[r * 3 + 1] = C.base[r * 4 + 1]; X.R[r * 3 + 2] = C.base[r * 4 + 2];
            X.p[r] = C.base[r * 4 + 3];
        }
        PoseF saved[kMaxSave];
        auto do_prims = [&](int p0, int p1) {
            for (int p = p0; p < p1; ++p) {
                const PrimK& pk = C.V.prims[p];
                ShapeW Bs;
                prim_world(X, pk, Bs);
                pose_store(my + (size_t)pk.slot * NT, NT, Bs);
            }
        };
        do_prims(0, C.n_base_prims);
#pragma unroll 1
        for (int j = 0; j < C.n_mov; ++j) {
            const TreeJointK& jk = C.V.tree[j];
            if (jk.load_slot == -2) {
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    X.R[r * 3 + 0] = C.base[r * 4 + 0]; X.R[r * 3 + 1] = C.base[r * 4 + 1]; X.R[r * 3 + 2] = C.base[r * 4 + 2];
                    X.p[r] = C.base[r * 4 + 3];
                }
            } else if (jk.load_slot >= 0) {
                X = saved[jk.load_slot];
            }
            nbxj::joint_step<float, TreeJointK>(X, jk, valid ? qf(jk.qidx) : 0.f);
            if (jk.save_slot >= 0) saved[jk.save_slot] = X;
            do_prims(jk.prim_begin, jk.prim_end);
        }
    }

    float run_ub = FLT_MAX;      // distance mode: no pair can beat the smallest centre distance seen so far
    const float4* env4 = reinterpret_cast<const float4*>(C.V.env_pairs);
    const float4* self4 = reinterpret_cast<const float4*>(C.V.self_pairs);
#pragma unroll 1
    for (int o = 0; o < n_obs; ++o) {
        const ObsK& ob = C.V.obs[o];
        const V3 oc = mk(ob.c[0], ob.c[1], ob.c[2]);
        const bool is_box = ob.type == NBX_PRIM_BOX;
        float R[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) R[k] = ob.R[k];
        const float hx = ob.dims[0] + ob.margin, hy = ob.dims[1] + ob.margin, hz = ob.dims[2] + ob.margin;
        const float orad = ob.brad;
        const int e1 = C.V.env_begin[o + 1];
#pragma unroll 2
        for (int e = C.V.env_begin[o]; e < e1; ++e) {
            const float4 rec = env4[e];                          // {a|b, sharp|flags|slot_a, slot_b|pad, rsum}
            const int slot_a = (int)(short)(__float_as_uint(rec.y) >> 16);
            const float* pc = my + (size_t)slot_a * NT;
            const V3 dc = mk(pc[0], pc[NT], pc[2 * NT]) - oc;
            float lbnd;
            if (is_box) {
                const float lx = fmaxf(fabsf(fmaf(R[0], dc.x, fmaf(R[3], dc.y, R[6] * dc.z))) - hx, 0.f);
                const float ly = fmaxf(fabsf(fmaf(R[1], dc.x, fmaf(R[4], dc.y, R[7] * dc.z))) - hy, 0.f);
                const float lz = fmaxf(fabsf(fmaf(R[2], dc.x, fmaf(R[5], dc.y, R[8] * dc.z))) - hz, 0.f);
                lbnd = sqrtf(fmaf(lx, lx, fmaf(ly, ly, lz * lz))) - rec.w;
            } else {
                lbnd = sqrtf(dot(dc, dc)) - rec.w - orad;
            }
            if (WANT_DIST) run_ub = fminf(run_ub, sqrtf(dot(dc, dc)));
            const bool surv = valid && lbnd <= (WANT_DIST ? run_ub : C.thr) + 1e-6f;
            push_candidates<WANT_DIST, NT>(C, surv, kEnvFlag | ((uint32_t)e << 8));
        }
    }
#pragma unroll 2
    for (int e = 0; e < n_self; ++e) {
        const float4 rec = self4[e];
        const uint32_t w1 = __float_as_uint(rec.y), w2 = __float_as_uint(rec.z);
        const int slot_a = (int)(short)(w1 >> 16), slot_b = (int)(short)(w2 & 0xffffu);
        const float* pa = my + (size_t)slot_a * NT;
        const float* pb = my + (size_t)slot_b * NT;
        const V3 ca = mk(pa[0], pa[NT], pa[2 * NT]);
        const V3 cb = mk(pb[0], pb[NT], pb[2 * NT]);
        const V3 dc = cb - ca;
        const float cd = sqrtf(dot(dc, dc));
        if (WANT_DIST) run_ub = fminf(run_ub, cd);
        const float cut = (WANT_DIST ? run_ub : C.thr) + 1e-6f;
        bool surv = valid && cd - rec.w <= cut;
        if (surv && ((w1 >> 8) & 0xffu)) {
            // thin links: capsule bound  d(cyl, cyl) >= d(axis segment, axis segment) - rA - rB
            const PairK pr = C.V.self_pairs[e];
            const PrimK& ka = C.V.prims[pr.a];
            const PrimK& kb = C.V.prims[pr.b];
            const V3 ua = mk(pa[3 * NT], pa[4 * NT], pa[5 * NT]);
            const V3 ub = mk(pb[3 * NT], pb[4 * NT], pb[5 * NT]);
            const float d = seg_seg_distance(ca, ua, ka.dims[1] + ka.margin, cb, ub, kb.dims[1] + kb.margin) -
                            (ka.dims[0] + ka.margin) - (kb.dims[0] + kb.margin);
            surv = d <= cut;
        }
        push_candidates<WANT_DIST, NT>(C, surv, (uint32_t)e << 8);
    }
}

template <bool WANT_DIST, int NT>
__device__ __forceinline__ void phase2(const Ctx& C) {
    const int count = min(*C.qcount, C.qcap);
    for (int k = threadIdx.x; k < count; k += NT) eval_candidate<WANT_DIST, NT>(C, C.queue[k]);
}

__host__ __device__ inline size_t work_smem_bytes(int nt, int pose_floats, int dof_floats_per_item) {
    return align16(sizeof(float) * (size_t)nt * (pose_floats > 0 ? pose_floats : 1)) + align16(sizeof(uint32_t) * (size_t)nt * kQueuePerThread) +
           align16(sizeof(float) * (size_t)nt * dof_floats_per_item) + align16(sizeof(int) * (size_t)(2 * nt + 4));
}

// ---------------------------------------------------------------------------------------------
template <bool WANT_DIST>
__global__ void __launch_bounds__(kCTile)
in_collision_kernel(const __grid_constant__ SceneDev S, const float* __restrict__ q, int64_t B, int dof, float thr,
                    uint8_t* __restrict__ verdict, float* __restrict__ min_dist) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Ctx C;
    unsigned char* p = stage_scene(S, smem_raw, C.V);
    C.poses = reinterpret_cast<float*>(p);              p += align16(sizeof(float) * (size_t)kCTile * max(S.scratch_floats, 1));
    C.queue = reinterpret_cast<uint32_t*>(p);           p += align16(sizeof(uint32_t) * (size_t)kCTile * kQueuePerThread);
    float* sq = reinterpret_cast<float*>(p);            p += align16(sizeof(float) * (size_t)kCTile * dof);
    C.flag = reinterpret_cast<int*>(p);
    C.qcount = C.flag + kCTile;
    C.item_flag = nullptr;
    C.qcap = kCTile * kQueuePerThread;
    C.base = S.base;
    C.n_mov = S.n_mov; C.n_prims = S.n_prims; C.n_base_prims = S.n_base_prims;
    C.thr = thr;

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
    C.flag[tid] = WANT_DIST ? float_key(FLT_MAX) : 0;
    if (tid == 0) *C.qcount = 0;
    __syncthreads();

    const float* myq = sq + tid * dof;
    phase1<WANT_DIST, kCTile>(C, tid < nvalid, [myq](int col) { return myq[col]; }, S.n_obs, S.n_self);
    __syncthreads();
    phase2<WANT_DIST, kCTile>(C);
    __syncthreads();
    if (tid < nvalid) {
        if (WANT_DIST) {
            const float d = key_float(C.flag[tid]);
            min_dist[tile0 + tid] = d;
            verdict[tile0 + tid] = d < thr ? 1 : 0;
        } else {
            verdict[tile0 + tid] = (uint8_t)C.flag[tid];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Edge validation.  Per edge (all float64, bit-for-bit the reference's NumPy arithmetic,
// connectors.py:62-69 / :84-93):
//   distance = ||goal - start||_2 ; distance <= float32 eps -> invalid
//   step = resolution / distance ; T_f = 1 (connect) or min(1, max_distance / distance) (steer)
//   samples: t_k = k * step for k < ceil(T_f / step)  (numpy.arange length rule), then t = T_f.
// A block prefix sum flattens the (edge, sample) space of the CTA's edges; rounds of kETile samples go
// through phase 1 / phase 2.  Samples of an edge that is already invalid are skipped, as the sequential
// reference does by returning None.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kETile)
edges_valid_kernel(const __grid_constant__ SceneDev S, const float* __restrict__ starts, const float* __restrict__ goals,
                   int64_t E, int dof, double resolution, double max_distance, float thr,
                   uint8_t* __restrict__ valid, int32_t* __restrict__ n_samples) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Ctx C;
    unsigned char* p = stage_scene(S, smem_raw, C.V);
    C.poses = reinterpret_cast<float*>(p);              p += align16(sizeof(float) * (size_t)kETile * max(S.scratch_floats, 1));
    C.queue = reinterpret_cast<uint32_t*>(p);           p += align16(sizeof(uint32_t) * (size_t)kETile * kQueuePerThread);
    float* s_start = reinterpret_cast<float*>(p);       p += align16(sizeof(float) * (size_t)kETile * 2 * dof);
    float* s_goal = s_start + (size_t)kETile * dof;
    int* s_hit = reinterpret_cast<int*>(p);             // [kETile] per EDGE: 1 = some sample collides
    int* s_item_edge = s_hit + kETile;                  // [kETile] per ITEM of the current round: its edge
    C.qcount = s_item_edge + kETile;                    p += align16(sizeof(int) * (size_t)(2 * kETile + 4));
    double* s_step = reinterpret_cast<double*>(p);      p += sizeof(double) * kETile;
    double* s_tf = reinterpret_cast<double*>(p);        p += sizeof(double) * kETile;
    int* s_off = reinterpret_cast<int*>(p);             // [kETile + 1] exclusive prefix of sample counts
    C.flag = s_hit;
    C.item_flag = s_item_edge;
    C.qcap = kETile * kQueuePerThread;
    C.base = S.base;
    C.n_mov = S.n_mov; C.n_prims = S.n_prims; C.n_base_prims = S.n_base_prims;
    C.thr = thr;

    const int64_t tile0 = (int64_t)blockIdx.x * kETile;
    const int nvalid = (int)min((int64_t)kETile, E - tile0);
    const int tid = threadIdx.x;
    for (int i = tid; i < nvalid * dof; i += kETile) {
        s_start[i] = starts[tile0 * dof + i];
        s_goal[i] = goals[tile0 * dof + i];
    }
    __syncthreads();

    int cnt = 0;
    int dead = 1;            // edges beyond nvalid, and edges shorter than float32 eps, have no samples and are invalid
    if (tid < nvalid) {
        double ss = 0.0;
        for (int k = 0; k < dof; ++k) {
            const double dlt = (double)s_start[tid * dof + k] - (double)s_goal[tid * dof + k];
            ss = fma(dlt, dlt, ss);
        }
        const double dist = sqrt(ss);
        if (dist > (double)1.1920928955078125e-07f) {
            const double step = resolution / dist;
            double tf = 1.0;
            if (max_distance > 0.0 && dist > max_distance) tf = max_distance / dist;
            const double len = ceil(tf / step);
            cnt = (int)fmin(len, 1.0e9) + 1;
            s_step[tid] = step;
            s_tf[tid] = tf;
            dead = 0;
        }
        if (n_samples != nullptr) n_samples[tile0 + tid] = cnt;
    }
    s_hit[tid] = dead;
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

    for (int round0 = 0; round0 < total; round0 += kETile) {
        const int s = round0 + tid;
        bool have = s < total;
        int e = 0;
        float t = 0.f;
        if (have) {
            // owning edge: last e with s_off[e] <= s
            int lo = 0, hi = kETile - 1;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (s_off[mid] <= s) lo = mid; else hi = mid - 1;
            }
            e = lo;
            const int k = s - s_off[e];
            const int last = s_off[e + 1] - s_off[e] - 1;
            t = (float)((k == last) ? s_tf[e] : (double)k * s_step[e]);
            have = ((volatile int*)s_hit)[e] == 0;      // already invalid: skip (the sequential reference stops there too)
        }
        s_item_edge[tid] = e;
        if (tid == 0) *C.qcount = 0;
        __syncthreads();
        const float* a = s_start + e * dof;
        const float* g = s_goal + e * dof;
        phase1<false, kETile>(C, have, [a, g, t](int col) { return fmaf(t, g[col], fmaf(-t, a[col], a[col])); }, S.n_obs, S.n_self);
        __syncthreads();
        phase2<false, kETile>(C);
        __syncthreads();
    }
    if (tid < nvalid) valid[tile0 + tid] = (uint8_t)(s_hit[tid] == 0);
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
    S.scratch_floats = ch->pose_floats;
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
    const size_t smem = scene_smem_bytes(S.n_mov, S.n_prims, S.n_obs, S.n_self, S.n_env) + work_smem_bytes(kCTile, S.scratch_floats, ch->dof);
    NBX_REQUIRE(smem <= 220 * 1024, "scene too large for shared memory (%zu bytes)", smem);
    const int64_t grid = (B + kCTile - 1) / kCTile;
    NBX_REQUIRE(grid <= 0x7fffffffLL, "batch too large for one launch");
    if (min_dist != nullptr) {
        NBX_CUDA_OK(cudaFuncSetAttribute(in_collision_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        in_collision_kernel<true><<<(unsigned)grid, kCTile, smem, stream>>>(S, q, B, ch->dof, thr, verdict, min_dist);
    } else {
        NBX_CUDA_OK(cudaFuncSetAttribute(in_collision_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        in_collision_kernel<false><<<(unsigned)grid, kCTile, smem, stream>>>(S, q, B, ch->dof, thr, verdict, min_dist);
    }
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
    const size_t smem = scene_smem_bytes(S.n_mov, S.n_prims, S.n_obs, S.n_self, S.n_env) + work_smem_bytes(kETile, S.scratch_floats, 2 * dof) +
                        sizeof(double) * 2 * kETile + sizeof(int) * (kETile + 4);
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
