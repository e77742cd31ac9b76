// nbx_collide.cu -- collision queries for sm_100a: batched Arm.in_collision, DiscreteConnector edge validation, minimum
// distance, closest pair with witness points and distance gradient, all-pairs proximities, IRIS bisection.
//
// Replaces reference numbotics/robots/arm.py:555-632 (collisions / closest_to / in_collision / distance_to /
// jacobian_proximity: one PyBullet state reset + getClosestPoints sweep per configuration),
// numbotics/planning/sampling_based/connectors.py:57-100 (connect / steer: a Python loop of validity_checker calls along a
// linear trajectory) and numbotics/planning/safe_sets.py:124-134 (bisection towards the contact boundary).
//
// Common structure: a warp owns a contiguous range of units and works through it in tiles of 32
// items, lane = item.  The lane walks the whole kinematic tree once (all link frames, not just one path), parks every link
// primitive's world-space core in the warp's shared-memory scratch [value][lane] (bank-conflict free) and runs the cheap
// conservative culls on every allowed pair; the scene (tree, primitives, obstacles, pair lists; a few KB) is staged in
// shared memory once per CTA and read with warp-uniform addresses.  Survivors go through the GJK narrow phase -- in a
// second kernel over a global queue for verdict queries, in the same warp for distance queries.  Issue-bound; HBM traffic
// is 4*dof bytes in + 1 byte out per configuration plus the candidate queue.
#include "nbx_common.cuh"
#include "nbx_gjk.cuh"
#include "nbx_joint.cuh"
#include <algorithm>
#include <cstdlib>
#include <mutex>

using namespace nbxg;

namespace {

constexpr int kMaxSave = 8;
constexpr int kGjkIters = 48;        // verdict mode: only near-threshold pairs ever get this far
constexpr int kGjkItersDist = 96;    // distance mode: run to convergence

struct SceneDev {
    const TreeJointK* tree;
    const PrimK* prims;
    const ObsK* obs;
    const PairK* self_pairs;
    const PairK* env_pairs;
    const int* pair_desc;
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
    const int* pair_desc;
    const int* env_begin;
};

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

__host__ __device__ inline size_t scene_smem_bytes(int n_mov, int n_prims, int n_obs, int n_self, int n_env) {
    return align16(sizeof(TreeJointK) * n_mov) + align16(sizeof(PrimK) * n_prims) + align16(sizeof(ObsK) * n_obs) +
           align16(sizeof(PairK) * n_self) + align16(sizeof(PairK) * n_env) + align16(sizeof(int) * (n_obs + 1));
}

// Asynchronous global -> shared copies (cp.async).  The scene arrays and the q tiles are staged with these so that all
// of a CTA's staging loads are in flight together (ncu, source page: six dependent load -> store round trips of the
// prologue and the per-tile q load were 11 % of the broad kernel's warp samples, all long_scoreboard).
__device__ __forceinline__ void cp_async4(void* sdst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(void* sdst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Programmatic dependent launch (sm_90+): the kernels of one verdict query -- and the 46 of a bisection -- follow each other
// on a stream.  Each one lets its successor be SCHEDULED as soon as all of its own CTAs have started
// (`griddepcontrol.launch_dependents`), and waits for its predecessor's completion and memory flush
// (`griddepcontrol.wait`) only after the part of its prologue that touches nothing a predecessor writes (staging the static
// scene, carving shared memory): launch latency, CTA start-up and scene staging overlap the predecessor's tail.  Both
// instructions are no-ops in a kernel launched without the attribute.  Rule kept by every kernel below: no global read of
// query data and no global write before pdl_wait().
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ void copy_words(void* dst, const void* src, size_t bytes) {
    const int n = (int)(bytes / 4);
    const int* s = static_cast<const int*>(src);
    int* d = static_cast<int*>(dst);
    for (int i = threadIdx.x; i < n; i += blockDim.x) cp_async4(d + i, s + i);
}

// Every thread waits for its own staging copies; the caller's __syncthreads() then publishes the scene to the CTA.
__device__ __forceinline__ void stage_wait() { cp_async_commit(); cp_async_wait_all(); }

__device__ __forceinline__ unsigned char* stage_scene(const SceneDev& S, unsigned char* sm, SceneSm& V) {
    unsigned char* p = sm;
    V.tree = reinterpret_cast<const TreeJointK*>(p);  copy_words(p, S.tree, sizeof(TreeJointK) * S.n_mov);  p += align16(sizeof(TreeJointK) * S.n_mov);
    V.prims = reinterpret_cast<const PrimK*>(p);      copy_words(p, S.prims, sizeof(PrimK) * S.n_prims);    p += align16(sizeof(PrimK) * S.n_prims);
    V.obs = reinterpret_cast<const ObsK*>(p);         copy_words(p, S.obs, sizeof(ObsK) * S.n_obs);         p += align16(sizeof(ObsK) * S.n_obs);
    V.self_pairs = reinterpret_cast<const PairK*>(p); copy_words(p, S.self_pairs, sizeof(PairK) * S.n_self); p += align16(sizeof(PairK) * S.n_self);
    V.env_pairs = reinterpret_cast<const PairK*>(p);  copy_words(p, S.env_pairs, sizeof(PairK) * S.n_env);   p += align16(sizeof(PairK) * S.n_env);
    V.env_begin = reinterpret_cast<const int*>(p);    copy_words(p, S.env_begin, sizeof(int) * (S.n_obs + 1));    p += align16(sizeof(int) * (S.n_obs + 1));
    V.pair_desc = nullptr;
    return p;                                            // copies still in flight: stage_wait() + __syncthreads() before use
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
// Warp-autonomous evaluation.  A warp owns a contiguous range of units (configurations, or edges) and works through
// it in tiles of 32 items (configurations, or interpolation samples) without ever meeting a block barrier:
//
//   phase 1      lane = item.  Walk the whole tree, park every primitive's world core in the warp's shared-memory
//                scratch ([value][lane], conflict free), run the cheap conservative culls on every allowed pair.
//                All lanes walk the same scene in lockstep; the rare survivors (~0.5 % of the pairs) are appended to
//                the warp's pending list with one ballot per pair and no atomics (the count is warp-uniform).
//   materialise  lane = pending survivor.  Its two world-space cores are copied from the scratch into a
//                self-contained queue entry, so the entry outlives the tile.
//   drain        lane = queue entry.  Whenever 32 entries have accumulated (typically every 3-4 tiles) they go
//                through the GJK narrow phase with every lane busy; the remainder is drained at the end of the range.
//
// Two arrangements of the narrow phase:
//   * verdict queries (in_collision, edge validation) are SPLIT over two kernels: the broad kernel appends its
//     entries to a global queue (env pairs in the first half, self pairs in the second, so that GJK warps see one
//     kind of pair) and a second, small kernel runs the GJK over the queue at full lane utilisation.  The broad
//     kernel then carries no GJK state: ~9 KB of shared memory per warp, 128 registers, 16 warps per SM (at 96 registers /
//     20 warps per SM the outer-loop state of a tile lived in local memory and a fifth of the warp samples waited on its
//     reloads: ncu long_scoreboard 1.57 per issue; 2^22 configurations 0.60 -> 0.56 ms with the spills gone).
//   * distance queries (min_distance) stay FUSED: the warp keeps a shared-memory queue and drains it itself whenever
//     32 entries have accumulated, because the running minimum of a configuration prunes its later candidates.
// The block-cooperative predecessor (phase 2 right after each tile's phase 1, between two __syncthreads) spent 26 % of
// its warp-cycles at the barrier behind the one or two warps that had candidates, and ran the GJK at ~30 % lane
// utilisation; see profiles/README.md.
// ---------------------------------------------------------------------------------------------
constexpr int kWarps = 4;            // autonomous warps per CTA (they only share the staged scene)
constexpr int kWT = NBX_WARP_TILE;   // items per tile
constexpr int kQCap = 40;            // materialised candidates per warp (kDist; < 32 wait between passes, a pass adds what fits)
constexpr int kPendCap = 256;        // survivors of one tile awaiting materialisation
constexpr int kEW = 34;              // words per queue entry
constexpr int kMaxRange = 256;       // units a warp owns
constexpr int kMaxRangeDist = 128;   // ... in a distance query, whose running minima stay in shared memory until the range is done
constexpr uint32_t kEnvFlag = 0x80000000u;

constexpr int kSplit = 0;            // verdict query, global queue + narrow kernel
constexpr int kSplitNeg = 2;         // the same for NEGATIVE thresholds: pairs whose margins the threshold exceeds need the cores'
                                     // penetration depth (EPA).  A separate instantiation so that the usual verdict kernels carry
                                     // neither its stack frame nor its registers (ptxas: 3.2 KB of stack and twice the spills when
                                     // the EPA call was reachable from them; in_collision 0.70 -> 0.97 ms)
constexpr int kDist = 1;             // distance query, fused

struct WCtx {
    SceneSm V;
    const float* base;        // 12 floats (kernel parameter)
    float* poses;             // [pose_floats][32]
    float* queue;             // kDist: [kEW][kQCap] in shared memory, word-major so that lanes = entries is conflict free
    int unit0;                // kDist: global index of the first unit of the warp's range (queue entries carry global unit indices)
    uint32_t* pend;           // [kPendCap] survivor codes: env flag | pair << 8 | item lane
    unsigned long long* flag; // kDist: [kMaxRangeDist] running minimum: (ordered key of the distance) << 32 | pair code
    int* item_fi;             // [32] item lane -> flag index (kDist: within the warp's range; kSplit: global unit index)
    int* item_aux;            // [32] edges only (else nullptr): the item's sample index on its edge, carried in the entry's pair-code word
    int* first_bad;           // edges only, nullable: per edge the smallest colliding sample index (atomicMin by the narrow phase)
    int n_mov, n_prims, n_base_prims;
    float thr;
    // kSplit
    float* gq;                // [kEW][2 * gq_half] global queue, word-major: env entries [0, gq_half), self entries [gq_half, 2 gq_half)
    int* gq_count;            // [2] entries requested so far (may exceed gq_half: the excess was evaluated in place)
    int gq_half;
    uint8_t* gflag;           // per unit: verdict (hit_value 1) or edge validity (hit_value 0)
    int hit_value;
    int gjk_iters;            // iteration cap of the verdict GJK (kGjkIters; NBX_GJK_ITERS overrides it for the tests)
};

__device__ __forceinline__ int float_key(float f) {            // order-preserving float -> int
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float key_float(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }
// running minimum of a distance query: distance key in the high word (unsigned order), pair code in the low word
__device__ __forceinline__ unsigned long long dist_entry(float d, uint32_t pc) {
    return ((unsigned long long)((uint32_t)float_key(d) ^ 0x80000000u) << 32) | pc;
}
__device__ __forceinline__ float entry_dist(unsigned long long e) { return key_float((int)((uint32_t)(e >> 32) ^ 0x80000000u)); }

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
#pragma unroll
    for (int k = 0; k < 9; ++k) S.ax[k] = 0.f;
    if (pk.type == NBX_PRIM_CYLINDER) {
        S.ax[0] = scr[3 * stride]; S.ax[1] = scr[4 * stride]; S.ax[2] = scr[5 * stride];
    } else if (pk.type == NBX_PRIM_BOX) {
#pragma unroll
        for (int k = 0; k < 9; ++k) S.ax[k] = scr[(3 + k) * stride];
    }
}

// One narrow-phase candidate: the two world-space cores, the margins to subtract, the flag it reports to.
struct Cand {
    ShapeW A, B;
    float m;
    int fi;
    uint32_t pc;    // pair code: kEnvFlag | index into the scene's env / self pair array
};

// survivor code -> candidate, from the warp's pose scratch (valid until the next tile's phase 1)
__device__ __forceinline__ void build_cand(const WCtx& C, uint32_t code, Cand& K) {
    const int item = (int)(code & 0xffu);
    const int e = (int)((code & ~kEnvFlag) >> 8);
    K.fi = C.item_fi[item];
    K.pc = C.item_aux != nullptr ? (uint32_t)C.item_aux[item] : ((code & kEnvFlag) | (uint32_t)e);
    float mA, mB;
    bool sharp;
    if (code & kEnvFlag) {
        const PairK pr = C.V.env_pairs[e];
        const PrimK& pb = C.V.prims[pr.a];
        const ObsK& o = C.V.obs[pr.b];
#pragma unroll
        for (int k = 0; k < 9; ++k) K.A.ax[k] = 0.f;
        obs_world(o, K.A);
        pose_load(C.poses + (size_t)pb.slot * kWT + item, kWT, pb, K.B);
        mA = o.margin; mB = pb.margin; sharp = pr.sharp != 0;
    } else {
        const PairK pr = C.V.self_pairs[e];
        const PrimK& pa = C.V.prims[pr.a];
        const PrimK& pb = C.V.prims[pr.b];
        pose_load(C.poses + (size_t)pa.slot * kWT + item, kWT, pa, K.A);
        pose_load(C.poses + (size_t)pb.slot * kWT + item, kWT, pb, K.B);
        mA = pa.margin; mB = pb.margin; sharp = pr.sharp != 0;
    }
    K.m = mA + mB;
    if (sharp) {
        // box-box under the Bullet margin model: sharp boxes (spheres keep their radius: point cores)
        unround(K.A, mA);
        unround(K.B, mB);
        K.m = (K.A.type == NBX_PRIM_SPHERE ? mA : 0.f) + (K.B.type == NBX_PRIM_SPHERE ? mB : 0.f);
    }
}

template <typename STRIDE>
__device__ __forceinline__ void store_cand(float* qd, int slot, const Cand& K, const STRIDE kQCap) {
    float* w = qd + slot;
    w[0 * kQCap] = __int_as_float(K.fi);
    w[1 * kQCap] = __int_as_float(K.A.type | (K.B.type << 8));
    w[2 * kQCap] = K.m;
    w[33 * kQCap] = __uint_as_float(K.pc);
    w[3 * kQCap] = K.A.c.x; w[4 * kQCap] = K.A.c.y; w[5 * kQCap] = K.A.c.z;
    w[18 * kQCap] = K.B.c.x; w[19 * kQCap] = K.B.c.y; w[20 * kQCap] = K.B.c.z;
#pragma unroll
    for (int k = 0; k < 9; ++k) { w[(6 + k) * kQCap] = K.A.ax[k]; w[(21 + k) * kQCap] = K.B.ax[k]; }
#pragma unroll
    for (int k = 0; k < 3; ++k) { w[(15 + k) * kQCap] = K.A.d[k]; w[(30 + k) * kQCap] = K.B.d[k]; }
}

template <typename STRIDE>
__device__ __forceinline__ void load_cand(const float* qd, int slot, Cand& K, const STRIDE kQCap) {
    const float* w = qd + slot;
    K.fi = __float_as_int(w[0 * kQCap]);
    const int ty = __float_as_int(w[1 * kQCap]);
    K.A.type = ty & 0xff; K.B.type = (ty >> 8) & 0xff;
    K.m = w[2 * kQCap];
    K.pc = __float_as_uint(w[33 * kQCap]);
    K.A.c = mk(w[3 * kQCap], w[4 * kQCap], w[5 * kQCap]);
    K.B.c = mk(w[18 * kQCap], w[19 * kQCap], w[20 * kQCap]);
#pragma unroll
    for (int k = 0; k < 9; ++k) { K.A.ax[k] = w[(6 + k) * kQCap]; K.B.ax[k] = w[(21 + k) * kQCap]; }
#pragma unroll
    for (int k = 0; k < 3; ++k) { K.A.d[k] = w[(15 + k) * kQCap]; K.B.d[k] = w[(30 + k) * kQCap]; }
}

// Narrow phase of one candidate.
template <int MODE>
__device__ __forceinline__ void eval_cand(const WCtx& C, const Cand& K) {
    if (MODE == kDist) {
        const float best = entry_dist(((volatile unsigned long long*)C.flag)[K.fi]);
        // distance values are stated to 1e-4 m: stop when the bracket [lb, ub] is below half of that;
        // curved cores (cylinders) converge only linearly, so the tolerance is a large part of the cost of this query
        const GjkOut g = gjk_distance<kGjkItersDist>(K.A, K.B, best + K.m, -FLT_MAX, 2e-5f, 4e-5f);
        if (g.pruned) return;
        // overlapping cores: the signed distance continues below -(mA + mB) by the cores' penetration depth (EPA, out of line)
        // ... and so does the contact regime: |v| below 5e-5 is grazing contact only if that part of the simplex lies on the
        // boundary; a chord through the interior of a DEEP overlap passes that close to the origin about once in 10^4 pairs
        // (and always in exactly symmetric poses), and core_penetration tells the two apart (depth 0 for a real graze)
        float d;
        if (g.overlap || g.contact) d = -K.m - core_penetration(K.A, K.B).depth;
        else d = g.d - K.m;
        atomicMin(&C.flag[K.fi], dist_entry(d, K.pc));
    } else {
        volatile uint8_t* f = C.gflag + K.fi;
        if (C.first_bad == nullptr) {
            if (*f == (uint8_t)C.hit_value) return;          // this unit is already decided
        } else if (((volatile int*)C.first_bad)[K.fi] <= (int)K.pc) {
            return;                                          // an earlier sample of this edge already collides
        }
        const float lim = C.thr + K.m;                       // the pair collides iff  core distance < lim
        bool hit;
        if (MODE != kSplitNeg || lim > 0.f) {
            // the bracket only has to close to a fraction of the stated 1e-4 m contact band: a pair whose [lb, ub] still straddles
            // the threshold at 2e-5 is inside the band whichever way it is called (round 1 iterated on to 1e-6: the narrow kernel
            // of a bisection round, where every candidate sits on the contact boundary, took 85 us instead of ~45)
            const GjkOut g = gjk_distance<kGjkIters>(K.A, K.B, lim, lim, 2e-5f, 1e-5f, C.gjk_iters);
            // Verdict: a pair is FREE only on proof -- a separating plane beyond the threshold (`pruned`).  Overlap, contact and
            // an upper bound below the threshold are hits; so is a bracket [lb, ub] that still straddles the threshold when the
            // iteration stops (converged to its tolerance, stalled, or out of iterations: rounded cylinders converge only
            // linearly).  The unsafe direction for in_collision / edges_valid / bisect would be to call such a pair free.
            hit = !g.pruned;
        } else if (MODE == kSplitNeg) {
            // negative threshold beyond this pair's margins: only cores that interpenetrate by more than -lim collide
            const GjkOut g = gjk_distance<kGjkIters>(K.A, K.B, 1e-6f, -1.f, 2e-5f, 1e-5f, C.gjk_iters);
            hit = (g.overlap || g.contact) && core_penetration(K.A, K.B).depth > -lim;
        }
        if (hit) {
            *f = (uint8_t)C.hit_value;
            if (C.first_bad != nullptr) atomicMin(&C.first_bad[K.fi], (int)K.pc);
        }
    }
}

// Evaluate a survivor on the spot, out of line so that phase 1 stays small: overflow of the pending list (> 64
// survivors in a tile of 32 items) or of the global queue.  Pathological scenes only; no loss, just slower.
template <int MODE>
__device__ __noinline__ void eval_now_cold(const WCtx& C, uint32_t code) {
    Cand K;
    build_cand(C, code, K);
    eval_cand<MODE>(C, K);
}

__device__ __forceinline__ float sq3(float x, float y, float z) { return fmaf(x, x, fmaf(y, y, z * z)); }

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

// append the surviving lanes' codes to the pending list: one ballot, no atomics (npend is warp-uniform).  A full list
// (> 256 survivors in a tile of 32 items: every lane deep inside several contacts) only raises `npend` past the capacity; the tile is then
// evaluated exhaustively after the culls (tile_exhaustive_cold) -- no call inside the cull loops, whose loop invariants
// would otherwise be spilled around it.
__device__ __forceinline__ void push_pending(const WCtx& C, int& npend, bool surv, uint32_t code) {
    const unsigned mask = __ballot_sync(0xffffffffu, surv);
    if (mask == 0) return;
    const int lane = threadIdx.x & 31;
    const int pos = npend + __popc(mask & ((1u << lane) - 1u));
    if (surv && pos < kPendCap) C.pend[pos] = code | (uint32_t)lane;
    npend += __popc(mask);
}

// every allowed pair of this lane's item through the narrow phase (the pose scratch of the tile is still valid)
template <int MODE>
__device__ __noinline__ void tile_exhaustive_cold(const WCtx& C, bool valid, int n_env, int n_self) {
    if (!valid) return;
    const uint32_t lane = threadIdx.x & 31;
    for (int e = 0; e < n_env; ++e) eval_now_cold<MODE>(C, kEnvFlag | ((uint32_t)e << 8) | lane);
    for (int e = 0; e < n_self; ++e) eval_now_cold<MODE>(C, ((uint32_t)e << 8) | lane);
}

template <int MODE>
__device__ __forceinline__ void materialise(const WCtx& C, int& npend, int& qn);

// phase 1 for this lane's item.  `valid` = the lane has an item; invalid lanes still take part in the ballots.
//   1a  walk the tree, park every primitive's world core in the warp's scratch;
//   1b  broad phase.  Obstacle-major for the environment pairs (one obstacle's frame stays in registers while the
//       primitives paired with it stream past: exact point-to-box distance against the primitive's bounding radius,
//       compared squared -- no square root -- and without the rotation for axis-aligned boxes), then the self pairs
//       (bounding spheres, refined by the capsule bound for cylinder-cylinder pairs that survive).
struct NoHook { __device__ __forceinline__ void operator()() const {} };

// `after_walk` runs once the whole warp has finished reading its configurations (between 1a and 1b): in_collision uses
// it to request the NEXT tile's q into the same staging tile, so that the load has the whole broad phase to land.
template <int MODE, bool FAST, typename QF, typename AW = NoHook>
__device__ __forceinline__ void phase1(const WCtx& C, bool valid, QF qf, int n_obs, int n_self, int& npend, int& qn, AW after_walk = AW()) {
    constexpr bool WANT_DIST = MODE == kDist;
    constexpr int NT = kWT;
    const int lane = threadIdx.x & 31;
    float* my = C.poses + lane;
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
            if (FAST) {
                // serial chain: the frame just reached is the parent of the next joint, nothing to save or restore
                // (a `saved[]` slot picked at run time would live in local memory)
            } else if (jk.load_slot == -2) {
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    X.R[r * 3 + 0] = C.base[r * 4 + 0]; X.R[r * 3 + 1] = C.base[r * 4 + 1]; X.R[r * 3 + 2] = C.base[r * 4 + 2];
                    X.p[r] = C.base[r * 4 + 3];
                }
            } else if (jk.load_slot >= 0) {
                X = saved[jk.load_slot];
            }
            if (FAST) {
                // serial DH-style arm (every joint a revolute about its own z, roll-only offsets: ch->fast_walk): no kind switch,
                // the offset rotation is a plane rotation of two columns
                float sn, cs;
                nbxj::sincos_t<float>(valid ? qf(jk.qidx) : 0.f, &sn, &cs);
                nbxj::joint_step_zx<float, TreeJointK>(X, jk, sn, cs);
            } else {
                nbxj::joint_step<float, TreeJointK>(X, jk, valid ? qf(jk.qidx) : 0.f);
            }
            if (!FAST && jk.save_slot >= 0) saved[jk.save_slot] = X;
            do_prims(jk.prim_begin, jk.prim_end);
        }
    }
    __syncwarp();
    after_walk();

    const float4* env4 = reinterpret_cast<const float4*>(C.V.env_pairs);
    const float4* self4 = reinterpret_cast<const float4*>(C.V.self_pairs);
    const char* myb = reinterpret_cast<const char*>(my);
    float cut0 = C.thr + 1e-6f;
    if (WANT_DIST) {
        // Distance mode, pass A: the tightest cheap UPPER bound of the minimum distance -- (exact distance from a
        // primitive's centre to the other shape, or centre to centre) minus the inscribed radii.  Pass B below then
        // keeps only the pairs whose lower bound does not exceed it: 2-4 GJK runs per configuration instead of ~15.
        float ub = FLT_MAX;
#pragma unroll 1
        for (int o = 0; o < n_obs; ++o) {
            const ObsK& ob = C.V.obs[o];
            const V3 oc = mk(ob.c[0], ob.c[1], ob.c[2]);
            const bool is_box = ob.type == NBX_PRIM_BOX;
            const float hx = ob.dims[0] + ob.margin, hy = ob.dims[1] + ob.margin, hz = ob.dims[2] + ob.margin;
            // box: l2 below is the distance to the SHARP full box, which contains the rounded one: the rounded surface is at
            // most (sqrt(3) - 1) * margin farther away (at a corner), so + margin keeps the bound an upper bound
            const float orin = is_box ? -ob.margin : __int_as_float(ob.pad[1]);
            for (int e = C.V.env_begin[o]; e < C.V.env_begin[o + 1]; ++e) {
                const float4 rec = env4[e];
                const float* pc = reinterpret_cast<const float*>(myb + __float_as_int(rec.y));
                const V3 dc = mk(pc[0], pc[NT], pc[2 * NT]) - oc;
                float l2 = dot(dc, dc);
                if (is_box)
                    l2 = sq3(fmaxf(fabsf(fmaf(ob.R[0], dc.x, fmaf(ob.R[3], dc.y, ob.R[6] * dc.z))) - hx, 0.f),
                             fmaxf(fabsf(fmaf(ob.R[1], dc.x, fmaf(ob.R[4], dc.y, ob.R[7] * dc.z))) - hy, 0.f),
                             fmaxf(fabsf(fmaf(ob.R[2], dc.x, fmaf(ob.R[5], dc.y, ob.R[8] * dc.z))) - hz, 0.f));
                ub = fminf(ub, sqrtf(l2) - orin - C.V.prims[__float_as_uint(rec.x) & 0xffu].rin);
            }
        }
        for (int e = 0; e < n_self; ++e) {
            const float4 rec = self4[e];
            const float* pa = reinterpret_cast<const float*>(myb + __float_as_int(rec.y));
            const float* pb = reinterpret_cast<const float*>(myb + __float_as_int(rec.z));
            const V3 dc = mk(pb[0], pb[NT], pb[2 * NT]) - mk(pa[0], pa[NT], pa[2 * NT]);
            const uint32_t ab = __float_as_uint(rec.x);
            ub = fminf(ub, sqrtf(dot(dc, dc)) - C.V.prims[ab & 0xffu].rin - C.V.prims[(ab >> 8) & 0xffu].rin);
        }
        cut0 = ub + 1e-5f;       // the bound is conservative up to rounding; equal lower bounds must survive
    }
    // one environment pair: squared distance l2 from the primitive's centre to the obstacle, against the reach
#define NBX_ENV_SURV(E, SURV, L2_EXPR, ORAD, INSIDE_EXPR)                                                           \
    bool SURV;                                                                                           \
    {                                                                                                    \
        const float4 rec = env4[E];        /* {a|b|sharp|both_cyl, off_a, off_b, rsum} */                \
        const float* pc = reinterpret_cast<const float*>(myb + __float_as_int(rec.y));                   \
        const V3 dc = mk(pc[0], pc[NT], pc[2 * NT]) - oc;                                                \
        const float l2 = (L2_EXPR);                                                                      \
        if (WANT_DIST) {                                                                                 \
            /* centre inside the obstacle (l2 == 0): the signed distance can be far below -radius, so the lower bound */ \
            /* continues inwards with the (negative) distance to the nearest face (INSIDE_EXPR) */        \
            const float cdist = l2 > 0.f ? sqrtf(l2) : fminf((INSIDE_EXPR), 0.f);                        \
            SURV = valid && cdist - rec.w - (ORAD) <= cut0;                                              \
        } else {                                                                                         \
            const float reach = rec.w + (ORAD) + cut0;     /* >= 0: thresholds are non-negative */       \
            SURV = valid && l2 <= reach * reach;                                                         \
        }                                                                                                \
    }
    // distance queries keep several pairs per item: empty the list into the queue before it can overflow
#define NBX_ENV_PUSH(E, SURV)                                                                            \
    {                                                                                                    \
        push_pending(C, npend, SURV, kEnvFlag | ((uint32_t)(E) << 8));                                   \
        if (WANT_DIST && npend > kPendCap - 32) materialise<MODE>(C, npend, qn);                         \
    }
    // The pairs of one obstacle in groups of four: the four culls are independent straight-line code and ONE vote decides
    // whether any lane kept any of them (survivors are rare: 0.28 per configuration out of 58 pairs), instead of a ballot
    // and a branch behind every pair (ncu: 4.6 % of the warp samples waited on those branches).
#define NBX_ENV_LOOP(L2_EXPR, ORAD, INSIDE_EXPR)                                                                    \
    {                                                                                                    \
        int e = e0;                                                                                      \
        _Pragma("unroll 1") for (; e + 4 <= e1; e += 4) {                                                \
            NBX_ENV_SURV(e, s0, L2_EXPR, ORAD, INSIDE_EXPR)                                                           \
            NBX_ENV_SURV(e + 1, s1, L2_EXPR, ORAD, INSIDE_EXPR)                                                       \
            NBX_ENV_SURV(e + 2, s2, L2_EXPR, ORAD, INSIDE_EXPR)                                                       \
            NBX_ENV_SURV(e + 3, s3, L2_EXPR, ORAD, INSIDE_EXPR)                                                       \
            if (__any_sync(0xffffffffu, s0 | s1 | s2 | s3)) {                                            \
                NBX_ENV_PUSH(e, s0) NBX_ENV_PUSH(e + 1, s1) NBX_ENV_PUSH(e + 2, s2) NBX_ENV_PUSH(e + 3, s3) \
            }                                                                                            \
        }                                                                                                \
        _Pragma("unroll 1") for (; e < e1; ++e) {                                                        \
            NBX_ENV_SURV(e, s0, L2_EXPR, ORAD, INSIDE_EXPR)                                                           \
            NBX_ENV_PUSH(e, s0)                                                                          \
        }                                                                                                \
    }
#pragma unroll 1
    for (int o = 0; o < n_obs; ++o) {
        const ObsK& ob = C.V.obs[o];
        const V3 oc = mk(ob.c[0], ob.c[1], ob.c[2]);
        const int e0 = C.V.env_begin[o], e1 = C.V.env_begin[o + 1];
        // the obstacle's kind is warp-uniform: one specialised inner loop each, no branch per pair
        if (ob.type == NBX_PRIM_BOX) {
            const float hx = ob.dims[0] + ob.margin, hy = ob.dims[1] + ob.margin, hz = ob.dims[2] + ob.margin;
            if (ob.pad[0] != 0) {              // axis-aligned box: exact point-to-box distance without the rotation
                NBX_ENV_LOOP(sq3(fmaxf(fabsf(dc.x) - hx, 0.f), fmaxf(fabsf(dc.y) - hy, 0.f), fmaxf(fabsf(dc.z) - hz, 0.f)), 0.f,
                             fmaxf(fabsf(dc.x) - hx, fmaxf(fabsf(dc.y) - hy, fabsf(dc.z) - hz)))
            } else {
                float R[9];
#pragma unroll
                for (int k = 0; k < 9; ++k) R[k] = ob.R[k];
                NBX_ENV_LOOP(sq3(fmaxf(fabsf(fmaf(R[0], dc.x, fmaf(R[3], dc.y, R[6] * dc.z))) - hx, 0.f),
                                 fmaxf(fabsf(fmaf(R[1], dc.x, fmaf(R[4], dc.y, R[7] * dc.z))) - hy, 0.f),
                                 fmaxf(fabsf(fmaf(R[2], dc.x, fmaf(R[5], dc.y, R[8] * dc.z))) - hz, 0.f)), 0.f,
                             fmaxf(fabsf(fmaf(R[0], dc.x, fmaf(R[3], dc.y, R[6] * dc.z))) - hx,
                                   fmaxf(fabsf(fmaf(R[1], dc.x, fmaf(R[4], dc.y, R[7] * dc.z))) - hy,
                                         fabsf(fmaf(R[2], dc.x, fmaf(R[5], dc.y, R[8] * dc.z))) - hz)))
            }
        } else {                               // sphere / cylinder obstacle: bounding spheres
            const float orad = ob.brad;
            NBX_ENV_LOOP(dot(dc, dc), orad, 0.f)
        }
    }
#undef NBX_ENV_LOOP
#undef NBX_ENV_PUSH
#undef NBX_ENV_SURV
#pragma unroll 4
    for (int e = 0; e < n_self; ++e) {
        const float4 rec = self4[e];
        const float* pa = reinterpret_cast<const float*>(myb + __float_as_int(rec.y));
        const float* pb = reinterpret_cast<const float*>(myb + __float_as_int(rec.z));
        const V3 ca = mk(pa[0], pa[NT], pa[2 * NT]);
        const V3 cb = mk(pb[0], pb[NT], pb[2 * NT]);
        const V3 dc = cb - ca;
        const float cd2 = dot(dc, dc);
        float cut;
        bool surv;
        if (WANT_DIST) {
            cut = cut0;
            surv = valid && sqrtf(cd2) - rec.w <= cut;
        } else {
            cut = cut0;
            const float reach = rec.w + cut0;
            surv = valid && cd2 <= reach * reach;
        }
        if (WANT_DIST && surv && (__float_as_uint(rec.x) >> 24)) {
            // thin links: capsule bound  d(cyl, cyl) >= d(axis segment, axis segment) - rA - rB.  Distance queries refine here
            // (the cut is the item's own upper bound); verdict queries refine in materialise, lane = survivor: here the
            // refinement ran behind a divergent branch with 1-2 lanes active for ~12 of the 28 pairs of a tile (ncu: a fifth of
            // the issue slots of the broad kernel)
            const PairK pr = C.V.self_pairs[e];
            const PrimK& ka = C.V.prims[pr.a];
            const PrimK& kb = C.V.prims[pr.b];
            const V3 ua = mk(pa[3 * NT], pa[4 * NT], pa[5 * NT]);
            const V3 ub = mk(pb[3 * NT], pb[4 * NT], pb[5 * NT]);
            const float d = seg_seg_distance(ca, ua, ka.dims[1] + ka.margin, cb, ub, kb.dims[1] + kb.margin) -
                            (ka.dims[0] + ka.margin) - (kb.dims[0] + kb.margin);
            surv = d <= cut;
        }
        push_pending(C, npend, surv, (uint32_t)e << 8);
        if (WANT_DIST && npend > kPendCap - 32) materialise<MODE>(C, npend, qn);
    }
    if (npend > kPendCap) {          // warp-uniform
        tile_exhaustive_cold<MODE>(C, valid, C.V.env_begin[n_obs], n_self);
        npend = 0;
    }
}

// kDist: GJK on the 32 newest queue entries (all lanes busy), or on what is left at the end of the range.  Candidates whose
// cores overlap (or sit in the contact regime) need the EPA, ~30 k instructions against a few hundred for a GJK and needed by
// 3 % of random configurations: run inline it had 1.2 of 32 lanes active and was two thirds of this kernel's time (ncu,
// profiles/r02_min_distance_*), most of it waiting for its 3.6 KB polytope in local memory -- this kernel's 190 KB of shared
// memory per SM leave ~30 KB of L1.  Such candidates go to a GLOBAL queue instead (global unit index in the entry); epa_kernel
// runs them, lane = candidate, with the whole L1 for its arrays, and lowers the per-unit minimum that this kernel leaves in
// global memory at the end of the range.  What the queue cannot take is evaluated here.
__device__ __forceinline__ void drain(const WCtx& C, int& qn, bool all) {
    const int lane = threadIdx.x & 31;
    while (qn >= 32 || (all && qn > 0)) {
        const int take = min(qn, 32);
        bool need_epa = false;
        Cand K;
        if (lane < take) {
            load_cand(C.queue, qn - take + lane, K, kQCap);
            const float best = entry_dist(((volatile unsigned long long*)C.flag)[K.fi]);
            const GjkOut g = gjk_distance<kGjkItersDist, true>(K.A, K.B, best + K.m, -FLT_MAX, 2e-5f, 4e-5f);     // as eval_cand<kDist>
            if (!g.pruned) {
                need_epa = g.overlap || g.contact;
                if (!need_epa) atomicMin(&C.flag[K.fi], dist_entry(g.d - K.m, K.pc));
            }
        }
        qn -= take;
        const unsigned m = __ballot_sync(0xffffffffu, need_epa);
        if (m != 0) {
            int base = 0;
            const int leader = __ffs(m) - 1;
            if (lane == leader) base = atomicAdd(&C.gq_count[0], __popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (need_epa) {
                const int pos = base + __popc(m & ((1u << lane) - 1u));
                if (pos < C.gq_half) {
                    const int local = K.fi;
                    K.fi = local + C.unit0;
                    store_cand(C.gq, pos, K, C.gq_half);
                } else {
                    atomicMin(&C.flag[K.fi], dist_entry(-K.m - core_penetration(K.A, K.B).depth, K.pc));
                }
            }
        }
        __syncwarp();
    }
}

// The EPA over the global queue of the distance query: persistent, few warps per SM (each warp's polytopes are ~115 KB of
// L1 lines), lane = candidate.  Lowers the running minimum (distance key << 32 | pair code) of the candidate's unit.
__global__ void __launch_bounds__(32)
epa_kernel(const float* __restrict__ gq, const int* __restrict__ gq_count, int gq_half, unsigned long long* __restrict__ entries) {
    const int n = min(gq_count[0], gq_half);
    for (int i = blockIdx.x * 32 + threadIdx.x; i < n; i += gridDim.x * 32) {
        Cand K;
        load_cand(gq, i, K, gq_half);
        const float d = -K.m - core_penetration(K.A, K.B).depth;
        atomicMin(&entries[K.fi], dist_entry(d, K.pc));
    }
}

// per-unit running minimum -> the caller's arrays
__global__ void dist_finalize_kernel(const unsigned long long* __restrict__ entries, int64_t B, float thr, uint8_t* __restrict__ verdict,
                                     float* __restrict__ min_dist, uint32_t* __restrict__ pair_code) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const unsigned long long e = entries[i];
    const float d = entry_dist(e);
    min_dist[i] = d;
    verdict[i] = d < thr ? 1 : 0;
    if (pair_code != nullptr) pair_code[i] = (uint32_t)e;
}

// Verdict queries: the capsule bound of a bounding-sphere survivor (self pair of two cylinders),
//   d(cyl, cyl) >= d(axis segment, axis segment) - rA - rB,
// evaluated by the lane that holds the survivor, from the tile's pose scratch.  false = the pair cannot be within the threshold.
__device__ __forceinline__ bool capsule_keep(const WCtx& C, uint32_t code) {
    const int item = (int)(code & 0xffu);
    const PairK pr = C.V.self_pairs[code >> 8];
    if (!pr.both_cyl) return true;
    const PrimK& ka = C.V.prims[pr.a];
    const PrimK& kb = C.V.prims[pr.b];
    const float* pa = reinterpret_cast<const float*>(reinterpret_cast<const char*>(C.poses + item) + pr.off_a);
    const float* pb = reinterpret_cast<const float*>(reinterpret_cast<const char*>(C.poses + item) + pr.off_b);
    constexpr int NT = kWT;
    const V3 ca = mk(pa[0], pa[NT], pa[2 * NT]), ua = mk(pa[3 * NT], pa[4 * NT], pa[5 * NT]);
    const V3 cb = mk(pb[0], pb[NT], pb[2 * NT]), ub = mk(pb[3 * NT], pb[4 * NT], pb[5 * NT]);
    const float d = seg_seg_distance(ca, ua, ka.dims[1] + ka.margin, cb, ub, kb.dims[1] + kb.margin) -
                    (ka.dims[0] + ka.margin) - (kb.dims[0] + kb.margin);
    return d <= C.thr + 1e-6f;
}

// pending survivors of the tile just walked -> self-contained queue entries (the scratch is overwritten by the next
// tile).  kDist: the warp's shared-memory queue, drained as soon as it holds 32 entries.  kSplit: the global queue;
// one atomic per (tile, pair kind) reserves the slots.
template <int MODE>
__device__ __forceinline__ void materialise(const WCtx& C, int& npend, int& qn) {
    const int lane = threadIdx.x & 31;
    __syncwarp();
    int done = 0;
    while (done < npend) {
        if (MODE == kDist) {
            const int chunk = min(min(32, npend - done), kQCap - qn);
            if (lane < chunk) {
                Cand K;
                build_cand(C, C.pend[done + lane], K);
                store_cand(C.queue, qn + lane, K, kQCap);
            }
            __syncwarp();
            qn += chunk;
            done += chunk;
            drain(C, qn, false);
        } else {
            const int chunk = min(32, npend - done);
            bool have = lane < chunk;
            const uint32_t code = have ? C.pend[done + lane] : 0u;
            const bool env = have && (code & kEnvFlag) != 0;
            if (have && !env) have = capsule_keep(C, code);      // self pairs of two cylinders: the capsule bound, lane = survivor
            const bool slf = have && !env;
            const unsigned menv = __ballot_sync(0xffffffffu, env), mslf = __ballot_sync(0xffffffffu, slf);
            int benv = 0, bslf = 0;
            if (lane == 0) {
                if (menv) benv = atomicAdd(&C.gq_count[0], __popc(menv));
                if (mslf) bslf = atomicAdd(&C.gq_count[1], __popc(mslf));
            }
            benv = __shfl_sync(0xffffffffu, benv, 0);
            bslf = __shfl_sync(0xffffffffu, bslf, 0);
            if (have) {
                const unsigned below = (1u << lane) - 1u;
                const int pos = env ? benv + __popc(menv & below) : bslf + __popc(mslf & below);
                if (pos < C.gq_half) {
                    Cand K;
                    build_cand(C, code, K);
                    store_cand(C.gq, env ? pos : C.gq_half + pos, K, 2 * C.gq_half);
                } else {
                    eval_now_cold<MODE>(C, code);
                }
            }
            done += chunk;
        }
    }
    npend = 0;
}

// per-warp work area: pose scratch, queue, pending list, flags, item -> flag map, + `extra_floats` for the caller
__host__ __device__ inline size_t warp_smem_bytes(int mode, int pose_floats, int extra_floats) {
    return align16(sizeof(float) * (size_t)kWT * (pose_floats > 0 ? pose_floats : 1)) +
           (mode == kDist ? align16(sizeof(float) * (size_t)kEW * kQCap) + align16(sizeof(unsigned long long) * kMaxRangeDist) : 0) +
           align16(sizeof(uint32_t) * kPendCap) + 2 * align16(sizeof(int) * kWT) + align16(sizeof(float) * (size_t)extra_floats);
}

template <int MODE>
__device__ __forceinline__ unsigned char* carve_warp(WCtx& C, unsigned char* p, int pose_floats) {
    C.poses = reinterpret_cast<float*>(p);      p += align16(sizeof(float) * (size_t)kWT * max(pose_floats, 1));
    C.queue = nullptr;
    C.flag = nullptr;
    if (MODE == kDist) {
        C.queue = reinterpret_cast<float*>(p);  p += align16(sizeof(float) * (size_t)kEW * kQCap);
        C.flag = reinterpret_cast<unsigned long long*>(p);   p += align16(sizeof(unsigned long long) * kMaxRangeDist);
    }
    C.pend = reinterpret_cast<uint32_t*>(p);    p += align16(sizeof(uint32_t) * kPendCap);
    C.item_fi = reinterpret_cast<int*>(p);      p += align16(sizeof(int) * kWT);
    C.item_aux = nullptr;                       // edges point it at the next kWT ints
    C.first_bad = nullptr;
    p += align16(sizeof(int) * kWT);
    return p;
}

struct SplitArgs {
    float* gq;
    int* gq_count;
    int gq_half;
    int gjk_iters;
    int* first_bad;           // edges only, nullable
    unsigned long long* entries;   // distance query only: per-unit running minimum in global memory
};

// ---------------------------------------------------------------------------------------------
// in_collision.  MODE kSplit: broad kernel of the verdict query (verdict[b] = 0 here, the narrow kernel raises it);
// MODE kDist: fused distance query (min_dist[b] and the verdict at the end of the warp's range).
template <int MODE, bool FAST>
__global__ void __launch_bounds__(kWarps * 32, MODE == kDist ? 3 : 4)
in_collision_kernel(const __grid_constant__ SceneDev S, const float* __restrict__ q, int64_t B, int dof, float thr,
                    uint8_t* __restrict__ verdict, float* __restrict__ min_dist, uint32_t* __restrict__ pair_code, int upw,
                    const SplitArgs G) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WCtx C;
    unsigned char* p = stage_scene(S, smem_raw, C.V);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const size_t wbytes = warp_smem_bytes(MODE, S.scratch_floats, kWT * dof);
    p = carve_warp<MODE>(C, p + (size_t)wid * wbytes, S.scratch_floats);
    float* sq = reinterpret_cast<float*>(p);             // [32][dof]
    C.base = S.base;
    C.n_mov = S.n_mov; C.n_prims = S.n_prims; C.n_base_prims = S.n_base_prims;
    C.thr = thr;
    C.gq = G.gq; C.gq_count = G.gq_count; C.gq_half = G.gq_half;
    C.gflag = verdict; C.hit_value = 1; C.gjk_iters = G.gjk_iters;

    const int64_t range0 = ((int64_t)blockIdx.x * kWarps + wid) * upw;
    const int rangeN = (int)max((int64_t)0, min((int64_t)upw, B - range0));
    const bool vec = (reinterpret_cast<uintptr_t>(q) & 15u) == 0;
    // request the q tile that starts at item t0 of the warp's range into the staging tile (asynchronous)
    auto request_q = [&](int t0) {
        const int nv = min(kWT, rangeN - t0);
        const float* src = q + (range0 + t0) * dof;
        const int n = nv * dof;
        if (nv == kWT && vec && (((range0 + t0) * dof) & 3) == 0) {
            const float4* s16 = reinterpret_cast<const float4*>(src);
            float4* d16 = reinterpret_cast<float4*>(sq);
            for (int i = lane; i < n / 4; i += 32) cp_async16(d16 + i, s16 + i);
        } else {
            for (int i = lane; i < n; i += 32) cp_async4(sq + i, src + i);
        }
        cp_async_commit();
    };
    pdl_launch_dependents();
    pdl_wait();                                          // q may be the predecessor's output (bisection midpoints)
    if (rangeN > 0) request_q(0);                        // travels together with the scene
    stage_wait();
    __syncthreads();                                     // the staged scene is the only thing the warps share
    if (rangeN == 0) return;
    for (int i = lane; i < rangeN; i += 32) {
        if (MODE == kDist) C.flag[i] = dist_entry(FLT_MAX, 0xffffffffu);
        else verdict[range0 + i] = 0;
    }
    C.unit0 = (int)range0;
    int npend = 0, qn = 0;
#pragma unroll 1
    for (int t0 = 0; t0 < rangeN; t0 += kWT) {
        const int nv = min(kWT, rangeN - t0);
        cp_async_wait_all();                             // this tile's q (requested during the previous tile's broad phase)
        __syncwarp();                                    // the previous tile's materialise has read item_fi
        // lanes beyond a ragged last tile repeat its first item (the same candidates twice: harmless for a verdict flag and
        // for the running minimum) -- so phase 1 carries no per-lane `valid` predicate through its joint and pair loops
        const int src = lane < nv ? lane : 0;
        C.item_fi[lane] = (MODE == kDist) ? t0 + src : (int)(range0 + t0 + src);
        __syncwarp();
        const float* myq = sq + src * dof;
        phase1<MODE, FAST>(C, true, [myq](int col) { return myq[col]; }, S.n_obs, S.n_self, npend, qn,
                           [&]() { if (t0 + kWT < rangeN) request_q(t0 + kWT); });
        materialise<MODE>(C, npend, qn);
    }
    if (MODE == kDist) {
        drain(C, qn, true);
        __syncwarp();
        // the running minima of the range, for epa_kernel to lower and dist_finalize_kernel to unpack (min_dist is the
        // caller's float array: the 64-bit entries live in the query's workspace, G.entries)
        for (int i = lane; i < rangeN; i += 32) G.entries[range0 + i] = C.flag[i];
    }
}

// ---------------------------------------------------------------------------------------------
// Edge validation (reference DiscreteConnector.connect / steer, connectors.py:57-100), flattened over the GRID.
//
// Per edge, all float64 and bit for bit the reference's NumPy arithmetic (connectors.py:62-69 / :84-93):
//   distance = ||goal - start||_2 ; distance <= float32 eps -> invalid
//   step = resolution / distance ; T_f = 1 (connect) or min(1, max_distance / distance) (steer)
//   samples: t_k = k * step for k < ceil(T_f / step)  (numpy.arange length rule), then t = T_f.
//
// Round 1 gave every warp 32 edges; at 125 k edges per GPU that was 1.65 waves of ragged CTAs (strong scaling 0.56 on 8
// GPUs).  Now:  (1) edge_prepare_kernel  -- one thread per edge: sample count, {step, T_f}, valid[e] = 1 (0 for an edge
// shorter than float32 eps) and per-block sums;  (2) edge_offsets_kernel turns the counts into int64 offsets of ONE flat
// sample space and writes the first edge of every draw;  (3) edges_flat_kernel -- a persistent grid whose warps draw runs
// of consecutive 32-sample tiles of that space from an atomic counter.  A lane finds
// its (edge, sample) in the 32 edges the warp keeps staged in shared memory (offsets, {step, T_f}, start and goal rows:
// restaged, coalesced, once per ~13 tiles -- no dependent global load per tile), interpolates its own configuration into
// the warp's q tile, and the tile goes through phase 1 like a tile of in_collision.  The narrow
// kernel clears valid[e] when a candidate of the edge collides (and keeps the smallest colliding sample index when the
// caller asks for it: that is how "useful" samples -- the ones the sequential reference evaluates -- are counted).
// ---------------------------------------------------------------------------------------------
struct EdgeMeta { double step, tf; };
constexpr int kPrepBlock = 256;

__global__ void __launch_bounds__(kPrepBlock)
edge_prepare_kernel(const float* __restrict__ starts, const float* __restrict__ goals, int64_t E, int dof, double resolution,
                    double max_distance, uint8_t* __restrict__ valid, int32_t* __restrict__ n_samples, int* __restrict__ first_bad,
                    EdgeMeta* __restrict__ meta, long long* __restrict__ offs /* [E + 1]: offs[e + 1] = count */,
                    long long* __restrict__ block_sums, int* __restrict__ gq_count, int* __restrict__ tile_counter) {
    pdl_launch_dependents();
    pdl_wait();                                            // the counters may still be read by the previous query's narrow kernel
    const int64_t e = (int64_t)blockIdx.x * kPrepBlock + threadIdx.x;
    if (e == 0) { gq_count[0] = 0; gq_count[1] = 0; *tile_counter = 0; }     // no memset node in front of the query
    int cnt = 0;
    if (e < E) {
        int dead = 1;            // edges shorter than float32 eps have no samples and are invalid
        double ss = 0.0;
        const float* a = starts + e * dof;
        const float* g = goals + e * dof;
        for (int k = 0; k < dof; ++k) {
            const double dlt = (double)a[k] - (double)g[k];
            ss = fma(dlt, dlt, ss);
        }
        const double dist = sqrt(ss);
        EdgeMeta m{0.0, 0.0};
        if (dist > (double)1.1920928955078125e-07f) {
            m.step = resolution / dist;
            m.tf = 1.0;
            if (max_distance > 0.0 && dist > max_distance) m.tf = max_distance / dist;
            const double len = ceil(m.tf / m.step);
            cnt = (int)fmin(len, 1.0e9) + 1;
            dead = 0;
        }
        meta[e] = m;
        if (n_samples != nullptr) n_samples[e] = cnt;
        if (first_bad != nullptr) first_bad[e] = 0x7fffffff;
        valid[e] = (uint8_t)(dead ? 0 : 1);
        offs[e + 1] = cnt;
    }
    // block sum of the counts (int64: 2^31 edges of up to 10^9 samples)
    __shared__ long long wsum[kPrepBlock / 32];
    long long x = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int i = 0; i < kPrepBlock / 32; ++i) t += wsum[i];
        block_sums[blockIdx.x] = t;
        if (blockIdx.x == 0) offs[0] = 0;
    }
}

// tiles per draw: >= 24 draws per warp so that the last draw (the tail) is a few per cent of a warp's work, <= 8 tiles so
// that a draw's fixed cost stays small next to them -- and never more than kMaxDraws draws, the size of the draw table
constexpr int kMaxDraws = 1 << 16;
__device__ __forceinline__ long long edge_tiles_per_draw(long long n_tiles, long long warps) {
    long long tpc = max(1LL, min(8LL, n_tiles / (warps * 24)));
    return max(tpc, (n_tiles + kMaxDraws - 1) / kMaxDraws);
}

// offs[e + 1] (count) -> inclusive prefix, i.e. offs[e] = first flat sample of edge e and offs[E] = total; and the DRAW
// TABLE: draw c of the persistent kernel starts at flat sample c * tpc * 32, draw_edge[c] = the edge that sample belongs to.
// Every block adds up the sums of ALL blocks itself (a few thousand numbers from L2) instead of waiting for a scan kernel:
// that gives it its own base and the grand total -- hence tiles-per-draw -- in one launch (round 2 started with four
// preparation launches in front of the edge kernel; at 125 k edges per GPU they were 10 % of the call).
__global__ void __launch_bounds__(kPrepBlock)
edge_offsets_kernel(long long* __restrict__ offs, int64_t E, const long long* __restrict__ block_sums, int n_blocks,
                    long long warps, int* __restrict__ draw_edge) {
    __shared__ long long red[2][kPrepBlock / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    pdl_launch_dependents();
    pdl_wait();
    long long below = 0, all = 0;
    for (int i = threadIdx.x; i < n_blocks; i += kPrepBlock) {
        const long long v = __ldg(block_sums + i);
        all += v;
        if (i < (int)blockIdx.x) below += v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        below += __shfl_down_sync(0xffffffffu, below, o);
        all += __shfl_down_sync(0xffffffffu, all, o);
    }
    if (lane == 0) { red[0][wid] = below; red[1][wid] = all; }
    __syncthreads();
    long long base = 0, total = 0;
    for (int i = 0; i < kPrepBlock / 32; ++i) { base += red[0][i]; total += red[1][i]; }
    __syncthreads();

    const int64_t e = (int64_t)blockIdx.x * kPrepBlock + threadIdx.x;
    const long long c = e < E ? offs[e + 1] : 0;
    long long x = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) red[0][wid] = x;
    __syncthreads();
    for (int i = 0; i < wid; ++i) base += red[0][i];
    if (e >= E) return;
    const long long hi = base + x, lo = hi - c;              // edge e owns the flat samples [lo, hi)
    offs[e + 1] = hi;
    // draws that start inside this edge
    const long long n_tiles = (total + kWT - 1) / kWT;
    const long long span = edge_tiles_per_draw(n_tiles, warps) * kWT;
    for (long long d = (lo + span - 1) / span; d * span < hi; ++d) draw_edge[d] = (int)e;
}

constexpr int kEdgeStage = 32;         // edges a warp keeps staged in shared memory (offsets, {step, T_f}, start and goal rows)

__host__ __device__ inline int edge_extra_floats(int dof) {
    // [32][dof] configurations of the tile + start / goal rows of the staged edges + {step, T_f} (double) + offsets (int64)
    return kWT * dof + 2 * kEdgeStage * dof + 4 * kEdgeStage + 2 * (kEdgeStage + 2);
}

__device__ __forceinline__ void cp_async8(void* sdst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(sdst)), "l"(gsrc) : "memory");
}

// The warps of the (resident, persistent) grid DRAW runs of consecutive 32-sample tiles of the flat sample space from an
// atomic counter: tiles cost the same only to first order (a tile with survivors also builds their queue entries), and with
// ~20 tiles per warp on an eighth of the edges a static split ends when the unluckiest of 2 368 warps does (measured:
// 0.38 ms against 0.34 ms for 125 k edges).  Nothing a tile needs is loaded when it starts: the 32 edges its samples can
// belong to -- offsets, {step, T_f}, start and goal rows -- are staged in shared memory by cp.async issued right after the
// PREVIOUS tile's configurations were interpolated (they land during that tile's walk and culls, the in_collision kernel's
// q-tile trick); the draw after next is taken from the counter, and the next draw's first edge read from the draw table, one
// tile ahead of their use.
template <bool FAST, int MODE>
__global__ void __launch_bounds__(kWarps * 32, 4)
edges_flat_kernel(const __grid_constant__ SceneDev S, const float* __restrict__ starts, const float* __restrict__ goals,
                  int64_t E, int dof, float thr, uint8_t* __restrict__ valid, const EdgeMeta* __restrict__ meta,
                  const long long* __restrict__ offs, int* __restrict__ tile_counter, const int* __restrict__ draw_edge,
                  const SplitArgs G) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WCtx C;
    unsigned char* p = stage_scene(S, smem_raw, C.V);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const size_t wbytes = warp_smem_bytes(kSplit, S.scratch_floats, edge_extra_floats(dof));
    p = carve_warp<MODE>(C, p + (size_t)wid * wbytes, S.scratch_floats);
    C.item_aux = C.item_fi + kWT;                           // carve_warp reserves 2 x kWT ints
    double* s_meta = reinterpret_cast<double*>(p);         // [32] {step, T_f}   (16-byte aligned: first in the extra area)
    long long* s_off = reinterpret_cast<long long*>(s_meta + 2 * kEdgeStage);   // [33] first flat sample of the staged edges
    float* sq = reinterpret_cast<float*>(s_off + kEdgeStage + 2);    // [32][dof] configurations of the tile
    float* s_start = sq + (size_t)kWT * dof;               // [32][dof] start rows of the staged edges
    float* s_goal = s_start + (size_t)kEdgeStage * dof;    // [32][dof]
    C.base = S.base;
    C.n_mov = S.n_mov; C.n_prims = S.n_prims; C.n_base_prims = S.n_base_prims;
    C.thr = thr;
    C.gq = G.gq; C.gq_count = G.gq_count; C.gq_half = G.gq_half;
    C.gflag = valid; C.hit_value = 0; C.gjk_iters = G.gjk_iters;
    C.first_bad = G.first_bad;
    pdl_launch_dependents();
    pdl_wait();                                            // offsets, draw table, counters: the preparation kernels' output
    stage_wait();
    __syncthreads();

    const long long total = __ldg(offs + E);
    const long long n_tiles = (total + kWT - 1) / kWT;
    const long long warps = (long long)gridDim.x * kWarps;
    const long long tpc = edge_tiles_per_draw(n_tiles, warps);
    const long long n_draws = (n_tiles + tpc - 1) / tpc;
    // request the window of 32 edges that starts at edge e0 into the staging buffers (asynchronous)
    auto request_window = [&](long long e0) {
        const int nE = (int)min((long long)kEdgeStage, E - e0);
        for (int i = lane; i <= kEdgeStage; i += 32) cp_async8(s_off + i, offs + min(e0 + i, (long long)E));
        if (lane < nE) cp_async16(s_meta + 2 * lane, meta + e0 + lane);
        const float* gs = starts + e0 * dof;
        const float* gg = goals + e0 * dof;
        for (int i = lane; i < nE * dof; i += 32) { cp_async4(s_start + i, gs + i); cp_async4(s_goal + i, gg + i); }
    };
    auto draw = [&]() { int c = 0; if (lane == 0) c = atomicAdd(tile_counter, 1); return c; };      // lane 0 holds the result
    auto first_edge = [&](long long c) { return c < n_draws ? (long long)__ldg(draw_edge + c) : 0LL; };
    // pipeline: cA = the draw being worked on (its window requested), cB = the next one (eB = its first edge), cC = the one after
    long long cA = __shfl_sync(0xffffffffu, draw(), 0);
    if (cA >= n_draws) return;
    long long cB = __shfl_sync(0xffffffffu, draw(), 0);
    int cC_raw = draw();
    long long st_e = first_edge(cA), eB = first_edge(cB);
    request_window(st_e);
    cp_async_commit();
    int npend = 0, qn = 0;
#pragma unroll 1
    for (;;) {
        const long long t_begin = cA * tpc, t_end = min(n_tiles, t_begin + tpc);
#pragma unroll 1
        for (long long tile = t_begin; tile < t_end; ++tile) {
        const long long s0 = tile * kWT;
        const long long s = min(s0 + lane, total - 1);      // lanes beyond the very last sample repeat it (harmless duplicates),
        constexpr bool have = true;                         // so phase 1 carries no per-lane predicate through its loops
        cp_async_wait_all();                                // the window of this tile (requested during the previous one)
        __syncwarp();                                       // ... visible to all lanes; previous materialise has read item_fi / item_aux
        long long e = st_e;
        int k = 0;
        if (have) {
            float t;
            const float *a, *g;
            if (s >= s_off[kEdgeStage] && st_e + kEdgeStage < E) {
                // beyond the staged edges although they start at the tile's first edge: more than 32 edges inside one tile,
                // only possible with edges that have no samples at all -- look the edge up the long way
                long long lo = st_e, hi = E - 1;
                while (lo < hi) {
                    const long long mid = (lo + hi + 1) >> 1;
                    if (__ldg(offs + mid) <= s) lo = mid; else hi = mid - 1;
                }
                e = lo;
                k = (int)(s - __ldg(offs + e));
                const int last = (int)(__ldg(offs + e + 1) - __ldg(offs + e)) - 1;
                const EdgeMeta m = meta[e];
                t = (float)((k == last) ? m.tf : (double)k * m.step);
                a = starts + e * dof; g = goals + e * dof;
            } else {
                int lo = 0, hi = kEdgeStage - 1;            // last j with s_off[j] <= s
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (s_off[mid] <= s) lo = mid; else hi = mid - 1;
                }
                e = st_e + lo;
                k = (int)(s - s_off[lo]);
                const int last = (int)(s_off[lo + 1] - s_off[lo]) - 1;
                t = (float)((k == last) ? s_meta[2 * lo + 1] : (double)k * s_meta[2 * lo]);
                a = s_start + lo * dof; g = s_goal + lo * dof;
            }
            float* mq = sq + lane * dof;
            for (int col = 0; col < dof; ++col) {
                const float av = a[col];
                mq[col] = fmaf(t, g[col], fmaf(-t, av, av));
            }
        }
        C.item_fi[lane] = (int)e;
        C.item_aux[lane] = k;
        // the next tile starts at, or after, the edge of this tile's last sample; is it still inside the staged edges?
        const long long e_next = __shfl_sync(0xffffffffu, e, 31);
        const long long s_last_next = min(s0 + 2 * kWT - 1, total - 1);
        const bool next_covered = e_next >= st_e && e_next < st_e + kEdgeStage && s_last_next < s_off[kEdgeStage];
        __syncwarp();                                       // every lane is done with the staged edges (and s_off above)
        if (tile + 1 < t_end) {
            if (!next_covered) { st_e = e_next; request_window(st_e); }      // lands while this tile is walked and culled
        } else if (cB < n_draws) {
            st_e = eB;                                      // last tile of the draw: the next draw's window
            request_window(st_e);
        }
        cp_async_commit();
        const float* myq = sq + lane * dof;
        phase1<MODE, FAST>(C, have, [myq](int col) { return myq[col]; }, S.n_obs, S.n_self, npend, qn);
        materialise<MODE>(C, npend, qn);
        }
        // rotate the pipeline: the atomic of the draw after next was issued a whole draw ago, its table entry is read now and
        // used one draw from now
        if (cB >= n_draws) break;
        cA = cB;
        cB = __shfl_sync(0xffffffffu, cC_raw, 0);
        eB = first_edge(cB);
        cC_raw = draw();
    }
}

// ---------------------------------------------------------------------------------------------
// Narrow kernel of the verdict queries: one thread per queue entry, grid-stride, first the environment pairs, then
// the self pairs (a warp works on one kind of pair).  The entry count lives on the device: no host round trip.
template <int MODE>
__global__ void __launch_bounds__(128)
narrow_kernel(const float* __restrict__ gq, const int* __restrict__ gq_count, int gq_half, float thr,
              uint8_t* __restrict__ gflag, int hit_value, int gjk_iters, int* __restrict__ first_bad) {
    WCtx C;
    C.first_bad = first_bad;
    C.item_aux = nullptr;
    C.thr = thr;
    C.gflag = gflag;
    C.hit_value = hit_value;
    C.gjk_iters = gjk_iters;
    pdl_launch_dependents();
    pdl_wait();
    const int n_env = min(gq_count[0], gq_half), n_self = min(gq_count[1], gq_half);
    const int stride = gridDim.x * blockDim.x;
    const int first = blockIdx.x * blockDim.x + threadIdx.x;
    for (int half = 0; half < 2; ++half) {
        const int n = half == 0 ? n_env : n_self;
        const int off = half == 0 ? 0 : gq_half;
        for (int i = first; i < n; i += stride) {
            Cand K;
            load_cand(gq, off + i, K, 2 * gq_half);
            eval_cand<MODE>(C, K);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Closest pair of a configuration with witness points, contact normal and the gradient of the distance
// (reference Arm.closest_to, robots/arm.py:599-600; Proximity of physics/chain.py:952-968; Arm.jacobian_proximity,
// robots/arm.py:621-632 for the closest pair).  One thread per configuration; `pair_code` comes from the distance
// query (kDist kernel).  Walk 1 rebuilds the two primitives of that pair and runs the witness GJK; walk 2 (only when
// the gradient is wanted) accumulates  n . (w_j x (p_S - p_j))  over the joints above the subject, minus the same
// over the joints above the target when the target is a link of the arm.
// Subject / target: link primitive / obstacle for environment pairs; the pair's two link primitives (device
// order) for self pairs.  The normal points from the target to the subject.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
closest_witness_kernel(const __grid_constant__ SceneDev S, const float* __restrict__ q, int64_t B, int dof,
                       const uint32_t* __restrict__ pair_code, float* __restrict__ dist, float* __restrict__ p_subject,
                       float* __restrict__ p_target, float* __restrict__ normal, int32_t* __restrict__ subject,
                       int32_t* __restrict__ target, uint8_t* __restrict__ valid, float* __restrict__ grad,
                       int* __restrict__ epa_list, int* __restrict__ epa_count, int phase) {
    // Two launches.  phase 0, thread = configuration: everything for a separated closest pair; a configuration whose closest
    // pair OVERLAPS (3 % of random ones) is only noted in `epa_list`.  phase 1, thread = list entry: the same configuration
    // again, now with the EPA -- every lane of a warp busy with one, instead of one lane of most warps running ~30 k
    // instructions while 31 wait (that was ~3 of the 3.7 ms this kernel took for 2^20 configurations).
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SceneSm V;
    stage_scene(S, smem_raw, V);
    stage_wait();
    __syncthreads();
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (phase == 1) {
        if (b >= (int64_t)*epa_count) return;
        b = epa_list[b];
    }
    if (b >= B) return;
    const uint32_t code = pair_code[b];
    const float* myq = q + b * dof;
    if (grad != nullptr)
        for (int k = 0; k < dof; ++k) grad[b * dof + k] = 0.f;
    if (code == 0xffffffffu) {              // no allowed pair produced a distance (every pair pruned): nothing to report
        dist[b] = FLT_MAX; valid[b] = 0; subject[b] = -1; target[b] = -1;
        for (int k = 0; k < 3; ++k) { p_subject[b * 3 + k] = 0.f; p_target[b * 3 + k] = 0.f; normal[b * 3 + k] = 0.f; }
        return;
    }
    const bool env = (code & kEnvFlag) != 0;
    const int e = (int)(code & ~kEnvFlag);
    const PairK pr = env ? V.env_pairs[e] : V.self_pairs[e];
    const int ia = pr.a;                        // link primitive (subject)
    const int ib = env ? -1 : pr.b;             // link primitive (target) or none
    ShapeW Ssub, Stgt;
#pragma unroll
    for (int k = 0; k < 9; ++k) { Ssub.ax[k] = 0.f; Stgt.ax[k] = 0.f; }
    float m_sub = V.prims[ia].margin, m_tgt;
    if (env) { obs_world(V.obs[pr.b], Stgt); m_tgt = V.obs[pr.b].margin; }
    else m_tgt = V.prims[ib].margin;

    auto walk = [&](auto&& on_prims, auto&& on_joint) {
        PoseF X;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            X.R[r * 3 + 0] = S.base[r * 4 + 0]; X.R[r * 3 + 1] = S.base[r * 4 + 1]; X.R[r * 3 + 2] = S.base[r * 4 + 2];
            X.p[r] = S.base[r * 4 + 3];
        }
        PoseF saved[kMaxSave];
        on_prims(X, 0, S.n_base_prims);
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
            nbxj::joint_step<float, TreeJointK>(X, jk, myq[jk.qidx]);
            if (jk.save_slot >= 0) saved[jk.save_slot] = X;
            on_joint(X, jk, j);
            on_prims(X, jk.prim_begin, jk.prim_end);
        }
    };
    walk([&](const PoseF& X, int p0, int p1) {
             for (int p = p0; p < p1; ++p) {
                 if (p == ia) prim_world(X, V.prims[p], Ssub);
                 if (p == ib) prim_world(X, V.prims[p], Stgt);
             }
         },
         [](const PoseF&, const TreeJointK&, int) {});
    if (pr.sharp) {
        unround(Ssub, m_sub);
        unround(Stgt, m_tgt);
        m_sub = Ssub.type == NBX_PRIM_SPHERE ? m_sub : 0.f;
        m_tgt = Stgt.type == NBX_PRIM_SPHERE ? m_tgt : 0.f;
    }
    const GjkWit g = gjk_witness<96>(Ssub, Stgt);
    subject[b] = V.prims[ia].desc;
    target[b] = env ? ((int)pr.b | 0x40000000) : V.prims[ib].desc;
    V3 n, ps, pt;
    if (g.overlap) {
        if (phase == 0) {
            epa_list[atomicAdd(epa_count, 1)] = (int)b;
            return;
        }
        // cores overlap: penetration depth, push-out normal and witness points from the EPA (out of line)
        const EpaOut ep = core_penetration(Ssub, Stgt);
        n = ep.n;
        ps = madd(ep.pa, -m_sub, n);
        pt = madd(ep.pb, m_tgt, n);
        dist[b] = -ep.depth - m_sub - m_tgt;
        valid[b] = 1;
    } else {
        n = (1.f / g.d) * (g.pa - g.pb);            // target -> subject
        ps = madd(g.pa, -m_sub, n);
        pt = madd(g.pb, m_tgt, n);
        dist[b] = g.d - m_sub - m_tgt;
        valid[b] = 1;
    }
    p_subject[b * 3 + 0] = ps.x; p_subject[b * 3 + 1] = ps.y; p_subject[b * 3 + 2] = ps.z;
    p_target[b * 3 + 0] = pt.x; p_target[b * 3 + 1] = pt.y; p_target[b * 3 + 2] = pt.z;
    normal[b * 3 + 0] = n.x; normal[b * 3 + 1] = n.y; normal[b * 3 + 2] = n.z;
    if (grad == nullptr) return;
    const unsigned long long anc_s = V.prims[ia].anc, anc_t = env ? 0ull : V.prims[ib].anc;
    walk([](const PoseF&, int, int) {},
         [&](const PoseF& X, const TreeJointK& jk, int j) {
             const bool on_s = (anc_s >> j) & 1ull, on_t = (anc_t >> j) & 1ull;
             if (on_s == on_t) return;             // above neither, or above both (the relative motion cancels)
             V3 w;
             if (jk.kind == NBX_KIND_X) w = mk(X.R[0], X.R[3], X.R[6]);
             else if (jk.kind == NBX_KIND_Y) w = mk(X.R[1], X.R[4], X.R[7]);
             else if (jk.kind == NBX_KIND_Z) w = mk(X.R[2], X.R[5], X.R[8]);
             else w = mk(fmaf(X.R[0], jk.ax[0], fmaf(X.R[1], jk.ax[1], X.R[2] * jk.ax[2])),
                         fmaf(X.R[3], jk.ax[0], fmaf(X.R[4], jk.ax[1], X.R[5] * jk.ax[2])),
                         fmaf(X.R[6], jk.ax[0], fmaf(X.R[7], jk.ax[1], X.R[8] * jk.ax[2])));
             if (jk.kind == NBX_KIND_X || jk.kind == NBX_KIND_Y || jk.kind == NBX_KIND_Z) w = jk.sgn * w;
             const V3 pj = mk(X.p[0], X.p[1], X.p[2]);
             float gj;
             if (jk.type == NBX_JOINT_REVOLUTE) gj = dot(n, cross(w, (on_s ? ps : pt) - pj));
             else gj = dot(n, w);
             grad[b * dof + jk.qidx] = on_s ? gj : -gj;
         });
}

// ---------------------------------------------------------------------------------------------
// Every allowed pair of every configuration: the batched form of Arm.collisions (reference robots/arm.py:555-580, one
// Proximity per allowed pair).  Warp-autonomous like phase 1: the warp walks a tile of 32 configurations, parks the world
// cores in its scratch, then loops over the pairs with every lane running the witness GJK of its own configuration on the
// same pair (no divergence in shape kinds).  Outputs are indexed [configuration][descriptor pair], self pairs first.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarps * 32)
pair_distances_kernel(const __grid_constant__ SceneDev S, const float* __restrict__ q, int64_t B, int dof,
                      float* __restrict__ dist, float* __restrict__ p_subject, float* __restrict__ p_target,
                      float* __restrict__ normal, uint8_t* __restrict__ valid_out, float* __restrict__ grad) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WCtx C;
    unsigned char* p = stage_scene(S, smem_raw, C.V);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // per warp: the primitives' world cores, and (gradient only) every joint's world axis and origin [6 n_mov][32]
    const size_t pose_bytes = align16(sizeof(float) * (size_t)kWT * max(S.scratch_floats, 1));
    const size_t jscr_bytes = grad != nullptr ? align16(sizeof(float) * (size_t)kWT * 6 * max(S.n_mov, 1)) : 0;
    p += (size_t)wid * (pose_bytes + jscr_bytes);
    C.poses = reinterpret_cast<float*>(p);
    float* jscr = reinterpret_cast<float*>(p + pose_bytes) + (threadIdx.x & 31);
    stage_wait();
    __syncthreads();
    const int64_t b = ((int64_t)blockIdx.x * kWarps + wid) * kWT + lane;
    const bool have = b < B;
    const float* myq = q + (have ? b : 0) * dof;
    float* my = C.poses + lane;
    {
        PoseF X;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            X.R[r * 3 + 0] = S.base[r * 4 + 0]; X.R[r * 3 + 1] = S.base[r * 4 + 1]; X.R[r * 3 + 2] = S.base[r * 4 + 2];
            X.p[r] = S.base[r * 4 + 3];
        }
        PoseF saved[kMaxSave];
        auto do_prims = [&](int p0, int p1) {
            for (int pi = p0; pi < p1; ++pi) {
                const PrimK& pk = C.V.prims[pi];
                ShapeW Bs;
                prim_world(X, pk, Bs);
                pose_store(my + (size_t)pk.slot * kWT, kWT, Bs);
            }
        };
        do_prims(0, S.n_base_prims);
#pragma unroll 1
        for (int j = 0; j < S.n_mov; ++j) {
            const TreeJointK& jk = C.V.tree[j];
            if (jk.load_slot == -2) {
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    X.R[r * 3 + 0] = S.base[r * 4 + 0]; X.R[r * 3 + 1] = S.base[r * 4 + 1]; X.R[r * 3 + 2] = S.base[r * 4 + 2];
                    X.p[r] = S.base[r * 4 + 3];
                }
            } else if (jk.load_slot >= 0) {
                X = saved[jk.load_slot];
            }
            nbxj::joint_step<float, TreeJointK>(X, jk, have ? myq[jk.qidx] : 0.f);
            if (jk.save_slot >= 0) saved[jk.save_slot] = X;
            if (grad != nullptr) {
                // world axis and origin of joint j: what d(point)/dq_j needs (w x (p - p_j) revolute, w prismatic)
                V3 w;
                if (jk.kind == NBX_KIND_X) w = jk.sgn * mk(X.R[0], X.R[3], X.R[6]);
                else if (jk.kind == NBX_KIND_Y) w = jk.sgn * mk(X.R[1], X.R[4], X.R[7]);
                else if (jk.kind == NBX_KIND_Z) w = jk.sgn * mk(X.R[2], X.R[5], X.R[8]);
                else w = mk(fmaf(X.R[0], jk.ax[0], fmaf(X.R[1], jk.ax[1], X.R[2] * jk.ax[2])),
                            fmaf(X.R[3], jk.ax[0], fmaf(X.R[4], jk.ax[1], X.R[5] * jk.ax[2])),
                            fmaf(X.R[6], jk.ax[0], fmaf(X.R[7], jk.ax[1], X.R[8] * jk.ax[2])));
                float* js = jscr + (size_t)j * 6 * kWT;
                js[0] = w.x; js[kWT] = w.y; js[2 * kWT] = w.z;
                js[3 * kWT] = X.p[0]; js[4 * kWT] = X.p[1]; js[5 * kWT] = X.p[2];
            }
            do_prims(jk.prim_begin, jk.prim_end);
        }
    }
    __syncwarp();
    const int n_env = C.V.env_begin[S.n_obs];
    const int n_pairs = S.n_self + n_env;
    const int* pair_desc = S.pair_desc;               // device pair -> descriptor pair (global memory, uniform reads)
#pragma unroll 1
    for (int k = 0; k < n_pairs; ++k) {
        const bool env = k >= S.n_self;
        const int e = env ? k - S.n_self : k;
        const PairK pr = env ? C.V.env_pairs[e] : C.V.self_pairs[e];
        const PrimK& pa = C.V.prims[pr.a];
        ShapeW Ssub, Stgt;
        float m_sub = pa.margin, m_tgt;
        pose_load(my + (size_t)pa.slot * kWT, kWT, pa, Ssub);
        if (env) {
#pragma unroll
            for (int i = 0; i < 9; ++i) Stgt.ax[i] = 0.f;
            obs_world(C.V.obs[pr.b], Stgt);
            m_tgt = C.V.obs[pr.b].margin;
        } else {
            const PrimK& pb = C.V.prims[pr.b];
            pose_load(my + (size_t)pb.slot * kWT, kWT, pb, Stgt);
            m_tgt = pb.margin;
        }
        if (pr.sharp) {
            unround(Ssub, m_sub);
            unround(Stgt, m_tgt);
            m_sub = Ssub.type == NBX_PRIM_SPHERE ? m_sub : 0.f;
            m_tgt = Stgt.type == NBX_PRIM_SPHERE ? m_tgt : 0.f;
        }
        const GjkWit g = gjk_witness<96>(Ssub, Stgt);
        if (!have) continue;
        const int64_t o = b * n_pairs + pair_desc[env ? S.n_self + e : e];
        V3 n, ps, pt;
        float d;
        if (g.overlap) {
            const EpaOut ep = core_penetration(Ssub, Stgt);
            n = ep.n;
            ps = madd(ep.pa, -m_sub, n);
            pt = madd(ep.pb, m_tgt, n);
            d = -ep.depth - m_sub - m_tgt;
        } else {
            n = (1.f / g.d) * (g.pa - g.pb);
            ps = madd(g.pa, -m_sub, n);
            pt = madd(g.pb, m_tgt, n);
            d = g.d - m_sub - m_tgt;
        }
        dist[o] = d;
        if (valid_out != nullptr) valid_out[o] = 1;
        if (p_subject != nullptr) { p_subject[o * 3 + 0] = ps.x; p_subject[o * 3 + 1] = ps.y; p_subject[o * 3 + 2] = ps.z; }
        if (p_target != nullptr) { p_target[o * 3 + 0] = pt.x; p_target[o * 3 + 1] = pt.y; p_target[o * 3 + 2] = pt.z; }
        if (normal != nullptr) { normal[o * 3 + 0] = n.x; normal[o * 3 + 1] = n.y; normal[o * 3 + 2] = n.z; }
        if (grad != nullptr) {
            // d distance / d q = n . J_v(subject, p_subject) - n . J_v(target, p_target)  (reference robots/arm.py:621-632):
            // joints above exactly one of the two bodies contribute, joints above both cancel
            const unsigned long long anc_s = pa.anc, anc_t = env ? 0ull : C.V.prims[pr.b].anc;
            float* go = grad + o * dof;
#pragma unroll 1
            for (int j = 0; j < S.n_mov; ++j) {
                const TreeJointK& jk = C.V.tree[j];
                const bool on_s = (anc_s >> j) & 1ull, on_t = (anc_t >> j) & 1ull;
                float gj = 0.f;
                if (on_s != on_t) {
                    const float* js = jscr + (size_t)j * 6 * kWT;
                    const V3 w = mk(js[0], js[kWT], js[2 * kWT]);
                    if (jk.type == NBX_JOINT_REVOLUTE) {
                        const V3 pj = mk(js[3 * kWT], js[4 * kWT], js[5 * kWT]);
                        gj = dot(n, cross(w, (on_s ? ps : pt) - pj));
                    } else {
                        gj = dot(n, w);
                    }
                    if (!on_s) gj = -gj;
                }
                go[jk.qidx] = gj;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Bisection towards the collision boundary (IRIS "zero-order hold" counter-example search, reference
// numbotics/planning/safe_sets.py:124-134): per sample an interval [lo, hi] in configuration space, lo free, hi
// colliding; each round tests the midpoint and keeps the half that still straddles the boundary.  Interval
// arithmetic in float64 exactly as the reference's NumPy ((lo + hi) / 2.0); the collision test sees the midpoint
// rounded to float32.  One elementwise kernel per round applies the previous verdict and emits the next midpoints.
template <typename T>
__global__ void bisect_step_kernel(double* __restrict__ lo, double* __restrict__ hi, const T* __restrict__ inside, int64_t inside_stride,
                                   const T* __restrict__ outside, const uint8_t* __restrict__ verdict, float* __restrict__ mid32,
                                   T* __restrict__ out, int64_t B, int dof, int round, int last, int* __restrict__ gq_count) {
    pdl_launch_dependents();
    pdl_wait();                                            // the previous round's narrow kernel is done and its verdicts visible
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { gq_count[0] = 0; gq_count[1] = 0; }
    if (i >= B * dof) return;
    const int64_t b = i / dof;
    double l, h;
    if (round == 0) {
        l = (double)inside[inside_stride ? i : i - b * dof];
        h = (double)outside[i];
    } else {
        l = lo[i]; h = hi[i];
        const double m = (l + h) / 2.0;
        if (verdict[b]) h = m; else l = m;
    }
    if (last) { out[i] = (T)h; return; }
    lo[i] = l; hi[i] = h;
    mid32[i] = (float)((l + h) / 2.0);
}

// Launch with programmatic stream serialization (see pdl_wait above): the kernel may be scheduled while its predecessor on
// the stream is still draining.  NBX_NO_PDL=1 turns the attribute off (A/B measurements).
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t stream, Args&&... args) {
    static const bool off = std::getenv("NBX_NO_PDL") != nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = off ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------------------------------------
int fill_scene_dev(const nbx_chain* ch, const nbx_scene* sc, const double* base_pose, SceneDev& S) {
    NBX_REQUIRE(ch != nullptr && sc != nullptr, "NULL handle");
    // a non-sticky CUDA error left behind by someone else's unchecked call (torch probes, a handle finalised out of order)
    // must not be blamed on the launch below; a real device fault is sticky and still shows up there
    (void)cudaGetLastError();
    NBX_REQUIRE(sc->chain == ch, "scene was created for a different chain");
    NBX_REQUIRE(base_pose != nullptr, "base_pose is NULL");
    S.tree = ch->d_tree;
    S.prims = sc->d_prims;
    S.obs = sc->d_obs;
    S.self_pairs = sc->d_self;
    S.env_pairs = sc->d_env;
    S.pair_desc = sc->d_pair_desc;
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

// units (configurations / edges) per warp: long ranges amortise the final partial drain, but the grid must keep
// every SM busy for several waves
int units_per_warp(int64_t n_units, int sm_count, int warps_per_sm, int waves = 6, int max_range = kMaxRange) {
    static const int env = std::getenv("NBX_COLLIDE_UPW") ? std::atoi(std::getenv("NBX_COLLIDE_UPW")) : 0;   // tuning only
    const int64_t tiles = (n_units + kWT - 1) / kWT, resident = (int64_t)sm_count * warps_per_sm;
    int64_t u;
    if (tiles <= resident * (max_range / kWT))
        u = (tiles + resident - 1) / resident;            // small batch (a bisection round, a planner's query): ONE wave, every
                                                          // warp resident from the start instead of a ragged second wave
    else
        u = n_units / (resident * waves * kWT);
    u = std::max<int64_t>(1, std::min<int64_t>(u, max_range / kWT));
    if (env > 0) u = std::min(env, max_range / kWT);
    return (int)u * kWT;
}

// stream-ordered workspace of a verdict query: the global queue and its two counters
struct SplitWs {
    void* p = nullptr;
    cudaStream_t st = nullptr;
    SplitArgs args{nullptr, nullptr, 0, kGjkIters, nullptr, nullptr};
    int alloc(int device, int64_t n_items, cudaStream_t stream, bool zero_counters = true) {
        st = stream;
        cudaMemPool_t pool = nullptr;
        int rc = nbx_workspace_pool(device, &pool);
        if (rc != NBX_OK) return rc;
        // ~0.3 candidates per item on random PRM configurations, ~3 on contact-seeking workloads (IRIS bisection rounds);
        // what does not fit is evaluated in place by the broad kernel (correct, but several times slower)
        int64_t half = std::max<int64_t>(4096, std::min<int64_t>(2 * n_items, 1 << 21));
        half = (half + 31) & ~(int64_t)31;
        const size_t bytes = 256 + sizeof(float) * (size_t)kEW * 2 * (size_t)half;
        NBX_CUDA_OK(cudaMallocFromPoolAsync(&p, bytes, pool, stream));
        if (zero_counters) NBX_CUDA_OK(cudaMemsetAsync(p, 0, 256, stream));
        args.gq_count = static_cast<int*>(p);
        args.gq = reinterpret_cast<float*>(static_cast<unsigned char*>(p) + 256);
        args.gq_half = (int)half;
        static const int iters_env = std::getenv("NBX_GJK_ITERS") ? std::atoi(std::getenv("NBX_GJK_ITERS")) : 0;   // tests only
        args.gjk_iters = iters_env > 0 ? iters_env : kGjkIters;
        return NBX_OK;
    }
    ~SplitWs() { if (p) cudaFreeAsync(p, st); }
};

int launch_narrow(const nbx_chain* ch, const SplitArgs& G, float thr, uint8_t* gflag, int hit_value, cudaStream_t stream) {
    static const int per_sm_env = std::getenv("NBX_NARROW_CTAS_PER_SM") ? std::atoi(std::getenv("NBX_NARROW_CTAS_PER_SM")) : 0;   // tuning only
    const unsigned grid = (unsigned)(ch->sm_count * (per_sm_env > 0 ? per_sm_env : 10));      // 5 CTAs of 89 registers are resident per SM: two even waves
    if (thr < 0.f)
        NBX_CUDA_OK(launch_pdl(narrow_kernel<kSplitNeg>, grid, 128, 0, stream, G.gq, G.gq_count, G.gq_half, thr, gflag, hit_value, G.gjk_iters, G.first_bad));
    else
        NBX_CUDA_OK(launch_pdl(narrow_kernel<kSplit>, grid, 128, 0, stream, G.gq, G.gq_count, G.gq_half, thr, gflag, hit_value, G.gjk_iters, G.first_bad));
    nbx_count_launch();
    NBX_CUDA_OK(cudaGetLastError());
    return NBX_OK;
}

// `shared`: workspace owned by the caller, counters already zero on `stream` (the bisection keeps ONE workspace for all of
// its rounds and resets the counters inside its own step kernel: no allocation, memset or free per round).
int launch_in_collision(const nbx_chain* ch, const nbx_scene* sc, const float* q, int64_t B, const double* base_pose,
                        float thr, uint8_t* verdict, float* min_dist, cudaStream_t stream, uint32_t* pair_code = nullptr,
                        const SplitArgs* shared = nullptr) {
    SceneDev S;
    int rc = fill_scene_dev(ch, sc, base_pose, S);
    if (rc != NBX_OK) return rc;
    NBX_REQUIRE(B >= 0, "negative batch size");
    if (B == 0) return NBX_OK;
    NBX_REQUIRE(q != nullptr && verdict != nullptr, "q / verdict is NULL");
    NBX_REQUIRE(B <= 0x7fffffffLL, "batch too large for one launch");
    NBX_CUDA_OK(cudaSetDevice(ch->device));
    const int mode = min_dist != nullptr ? kDist : kSplit;
    const size_t smem = scene_smem_bytes(S.n_mov, S.n_prims, S.n_obs, S.n_self, S.n_env) + kWarps * warp_smem_bytes(mode, S.scratch_floats, kWT * ch->dof);
    NBX_REQUIRE(smem <= 220 * 1024, "scene too large for shared memory (%zu bytes)", smem);
    const int upw = mode == kDist ? units_per_warp(B, ch->sm_count, 12, 6, kMaxRangeDist) : units_per_warp(B, ch->sm_count, 16);
    const int64_t grid = (B + (int64_t)kWarps * upw - 1) / ((int64_t)kWarps * upw);
    if (mode == kDist) {
        // workspace of the distance query: [counters 256 B][running minimum per unit, 8 B][EPA queue, kEW words per entry]
        cudaMemPool_t pool = nullptr;
        if ((rc = nbx_workspace_pool(ch->device, &pool)) != NBX_OK) return rc;
        const int64_t cap = std::max<int64_t>(4096, std::min<int64_t>((B / 4 + 31) & ~(int64_t)31, 1 << 20));   // ~4 % of random configurations collide
        const size_t entries_bytes = (sizeof(unsigned long long) * (size_t)B + 255) & ~(size_t)255;
        void* wsp = nullptr;
        NBX_CUDA_OK(cudaMallocFromPoolAsync(&wsp, 256 + entries_bytes + sizeof(float) * (size_t)kEW * (size_t)cap, pool, stream));
        NBX_CUDA_OK(cudaMemsetAsync(wsp, 0, 256, stream));
        SplitArgs G{nullptr, nullptr, 0, kGjkIters, nullptr, nullptr};
        G.gq_count = static_cast<int*>(wsp);
        G.entries = reinterpret_cast<unsigned long long*>(static_cast<unsigned char*>(wsp) + 256);
        G.gq = reinterpret_cast<float*>(static_cast<unsigned char*>(wsp) + 256 + entries_bytes);
        G.gq_half = (int)cap;
        NBX_SMEM_LIMIT((in_collision_kernel<kDist, false>), smem);
        in_collision_kernel<kDist, false><<<(unsigned)grid, kWarps * 32, smem, stream>>>(S, q, B, ch->dof, thr, verdict, min_dist, pair_code, upw, G);
        nbx_count_launch();
        static const int epa_cps_env = std::getenv("NBX_EPA_CTAS_PER_SM") ? std::atoi(std::getenv("NBX_EPA_CTAS_PER_SM")) : 0;   // tuning only
        epa_kernel<<<(unsigned)(ch->sm_count * (epa_cps_env > 0 ? epa_cps_env : 16)), 32, 0, stream>>>(G.gq, G.gq_count, G.gq_half, G.entries);
        nbx_count_launch();
        dist_finalize_kernel<<<(unsigned)((B + 255) / 256), 256, 0, stream>>>(G.entries, B, thr, verdict, min_dist, pair_code);
        nbx_count_launch();
        const cudaError_t e = cudaGetLastError();
        cudaFreeAsync(wsp, stream);
        NBX_CUDA_OK(e);
        return NBX_OK;
    }
    SplitWs ws;
    if (shared == nullptr && (rc = ws.alloc(ch->device, B, stream)) != NBX_OK) return rc;
    const SplitArgs& G = shared != nullptr ? *shared : ws.args;
    auto kern = thr < 0.f ? (ch->fast_walk ? in_collision_kernel<kSplitNeg, true> : in_collision_kernel<kSplitNeg, false>)
                          : (ch->fast_walk ? in_collision_kernel<kSplit, true> : in_collision_kernel<kSplit, false>);
    NBX_SMEM_LIMIT(kern, smem);
    NBX_CUDA_OK(launch_pdl(kern, (unsigned)grid, kWarps * 32, smem, stream, S, q, B, ch->dof, thr, verdict, (float*)nullptr, (uint32_t*)nullptr, upw, G));
    nbx_count_launch();
    NBX_CUDA_OK(cudaGetLastError());
    return launch_narrow(ch, G, thr, verdict, 1, stream);
}

// persistent grid of the flattened edge kernel: what the launch bounds make resident
int edges_grid(const nbx_chain* ch) {
    static const int env = std::getenv("NBX_EDGE_CTAS_PER_SM") ? std::atoi(std::getenv("NBX_EDGE_CTAS_PER_SM")) : 0;   // tuning only
    return ch->sm_count * (env > 0 ? env : 4);
}

int launch_edges(const nbx_chain* ch, const nbx_scene* sc, const float* starts, const float* goals, int64_t E,
                 const double* base_pose, double resolution, double max_distance, float thr, uint8_t* valid,
                 int32_t* n_samples, int32_t* first_invalid, cudaStream_t stream) {
    SceneDev S;
    int rc = fill_scene_dev(ch, sc, base_pose, S);
    if (rc != NBX_OK) return rc;
    NBX_REQUIRE(E >= 0, "negative edge count");
    if (E == 0) return NBX_OK;
    NBX_REQUIRE(starts != nullptr && goals != nullptr && valid != nullptr, "starts / goals / valid is NULL");
    NBX_REQUIRE(resolution > 0.0, "Resolution must be positive");
    NBX_REQUIRE(E <= 0x7fffffffLL, "too many edges for one launch");
    NBX_CUDA_OK(cudaSetDevice(ch->device));
    const int dof = ch->dof;
    const size_t smem = scene_smem_bytes(S.n_mov, S.n_prims, S.n_obs, S.n_self, S.n_env) +
                        kWarps * warp_smem_bytes(kSplit, S.scratch_floats, edge_extra_floats(dof));
    NBX_REQUIRE(smem <= 220 * 1024, "scene too large for shared memory (%zu bytes)", smem);
    SplitWs ws;
    if ((rc = ws.alloc(ch->device, E * 8, stream, /*zero_counters=*/false)) != NBX_OK) return rc;      // edge_prepare_kernel zeroes them
    ws.args.first_bad = first_invalid;
    // per-call workspace: offsets (int64 [E + 1]), {step, T_f} per edge, block sums, the tile counter
    const int n_blocks = (int)((E + kPrepBlock - 1) / kPrepBlock);
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t b_offs = up(sizeof(long long) * (size_t)(E + 1)), b_meta = up(sizeof(EdgeMeta) * (size_t)E),
                 b_sums = up(sizeof(long long) * (size_t)n_blocks), b_draws = up(sizeof(int) * (size_t)kMaxDraws);
    cudaMemPool_t pool = nullptr;
    if ((rc = nbx_workspace_pool(ch->device, &pool)) != NBX_OK) return rc;
    struct Tmp { void* p = nullptr; cudaStream_t st = nullptr; ~Tmp() { if (p) cudaFreeAsync(p, st); } } tmp;
    tmp.st = stream;
    NBX_CUDA_OK(cudaMallocFromPoolAsync(&tmp.p, b_offs + b_meta + b_sums + b_draws + 256, pool, stream));
    unsigned char* w = static_cast<unsigned char*>(tmp.p);
    long long* offs = reinterpret_cast<long long*>(w);
    EdgeMeta* meta = reinterpret_cast<EdgeMeta*>(w + b_offs);
    long long* sums = reinterpret_cast<long long*>(w + b_offs + b_meta);
    int* draws = reinterpret_cast<int*>(w + b_offs + b_meta + b_sums);
    int* counter = reinterpret_cast<int*>(w + b_offs + b_meta + b_sums + b_draws);
    const int grid_main = edges_grid(ch);
    NBX_CUDA_OK(launch_pdl(edge_prepare_kernel, (unsigned)n_blocks, kPrepBlock, 0, stream, starts, goals, E, dof, resolution, max_distance,
                           valid, n_samples, first_invalid, meta, offs, sums, ws.args.gq_count, counter));
    NBX_CUDA_OK(launch_pdl(edge_offsets_kernel, (unsigned)n_blocks, kPrepBlock, 0, stream, offs, E, (const long long*)sums, n_blocks,
                           (long long)grid_main * kWarps, draws));
    nbx_count_launch(2);
    NBX_CUDA_OK(cudaGetLastError());
    auto kern = thr < 0.f ? (ch->fast_walk ? edges_flat_kernel<true, kSplitNeg> : edges_flat_kernel<false, kSplitNeg>)
                          : (ch->fast_walk ? edges_flat_kernel<true, kSplit> : edges_flat_kernel<false, kSplit>);
    NBX_SMEM_LIMIT(kern, smem);
    NBX_CUDA_OK(launch_pdl(kern, (unsigned)grid_main, kWarps * 32, smem, stream, S, starts, goals, E, dof, thr, valid, (const EdgeMeta*)meta,
                           (const long long*)offs, counter, (const int*)draws, ws.args));
    nbx_count_launch();
    NBX_CUDA_OK(cudaGetLastError());
    return launch_narrow(ch, ws.args, thr, valid, 0, stream);
}

int launch_closest(const nbx_chain* ch, const nbx_scene* sc, const float* q, int64_t B, const double* base_pose,
                   float* dist, float* p_subject, float* p_target, float* normal, int32_t* subject, int32_t* target,
                   uint8_t* valid, float* grad, cudaStream_t stream) {
    NBX_REQUIRE(ch != nullptr && sc != nullptr, "NULL handle");
    NBX_REQUIRE(B >= 0, "negative batch size");
    if (B == 0) return NBX_OK;
    NBX_REQUIRE(q != nullptr && dist != nullptr && p_subject != nullptr && p_target != nullptr && normal != nullptr &&
                subject != nullptr && target != nullptr && valid != nullptr, "NULL output buffer");
    NBX_REQUIRE(grad == nullptr || ch->n_mov <= 64, "distance gradients support chains of at most 64 movable joints");
    NBX_CUDA_OK(cudaSetDevice(ch->device));
    cudaMemPool_t pool = nullptr;
    int rc = nbx_workspace_pool(ch->device, &pool);
    if (rc != NBX_OK) return rc;
    struct Tmp { void* p = nullptr; cudaStream_t st = nullptr; ~Tmp() { if (p) cudaFreeAsync(p, st); } } tmp;
    tmp.st = stream;
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    NBX_CUDA_OK(cudaMallocFromPoolAsync(&tmp.p, up(sizeof(uint32_t) * (size_t)B) + up((size_t)B), pool, stream));
    uint32_t* code = static_cast<uint32_t*>(tmp.p);
    uint8_t* verdict = static_cast<unsigned char*>(tmp.p) + up(sizeof(uint32_t) * (size_t)B);
    // distance pass (fused kDist kernel): minimum distance and the pair that attains it; `dist` is rewritten below
    rc = launch_in_collision(ch, sc, q, B, base_pose, 0.f, verdict, dist, stream, code);
    if (rc != NBX_OK) return rc;
    SceneDev S;
    rc = fill_scene_dev(ch, sc, base_pose, S);
    if (rc != NBX_OK) return rc;
    const size_t smem = scene_smem_bytes(S.n_mov, S.n_prims, S.n_obs, S.n_self, S.n_env);
    NBX_REQUIRE(smem <= 220 * 1024, "scene too large for shared memory (%zu bytes)", smem);
    NBX_SMEM_LIMIT(closest_witness_kernel, smem);
    // [count][list of the configurations whose closest pair overlaps]
    int* wl = nullptr;
    NBX_CUDA_OK(cudaMallocFromPoolAsync(reinterpret_cast<void**>(&wl), sizeof(int) * ((size_t)B + 64), pool, stream));
    NBX_CUDA_OK(cudaMemsetAsync(wl, 0, sizeof(int), stream));
    for (int phase = 0; phase < 2; ++phase) {
        closest_witness_kernel<<<(unsigned)((B + 127) / 128), 128, smem, stream>>>(S, q, B, ch->dof, code, dist, p_subject, p_target, normal,
                                                                                  subject, target, valid, grad, wl + 64, wl, phase);
        nbx_count_launch();
    }
    const cudaError_t e = cudaGetLastError();
    cudaFreeAsync(wl, stream);
    NBX_CUDA_OK(e);
    return NBX_OK;
}

int launch_pair_distances(const nbx_chain* ch, const nbx_scene* sc, const float* q, int64_t B, const double* base_pose,
                          float* dist, float* p_subject, float* p_target, float* normal, uint8_t* valid, float* grad,
                          cudaStream_t stream) {
    SceneDev S;
    int rc = fill_scene_dev(ch, sc, base_pose, S);
    if (rc != NBX_OK) return rc;
    NBX_REQUIRE(B >= 0, "negative batch size");
    if (B == 0 || sc->n_self + sc->n_env == 0) return NBX_OK;
    NBX_REQUIRE(q != nullptr && dist != nullptr, "q / distance is NULL");
    NBX_CUDA_OK(cudaSetDevice(ch->device));
    NBX_REQUIRE(grad == nullptr || ch->n_mov <= 64, "distance gradients support chains of at most 64 movable joints");
    const size_t smem = scene_smem_bytes(S.n_mov, S.n_prims, S.n_obs, S.n_self, S.n_env) +
                        kWarps * (align16(sizeof(float) * (size_t)kWT * std::max(S.scratch_floats, 1)) +
                                  (grad != nullptr ? align16(sizeof(float) * (size_t)kWT * 6 * std::max(S.n_mov, 1)) : 0));
    NBX_REQUIRE(smem <= 220 * 1024, "scene too large for shared memory (%zu bytes)", smem);
    NBX_SMEM_LIMIT(pair_distances_kernel, smem);
    const int64_t grid = (B + kWarps * kWT - 1) / (kWarps * kWT);
    NBX_REQUIRE(grid <= 0x7fffffffLL, "batch too large for one launch");
    pair_distances_kernel<<<(unsigned)grid, kWarps * 32, smem, stream>>>(S, q, B, ch->dof, dist, p_subject, p_target, normal, valid, grad);
    nbx_count_launch();
    NBX_CUDA_OK(cudaGetLastError());
    return NBX_OK;
}

template <typename T>
int launch_bisect(const nbx_chain* ch, const nbx_scene* sc, const T* inside, int64_t inside_stride, const T* outside, int64_t B,
                  const double* base_pose, int n_bisections, float thr, T* out, cudaStream_t stream) {
    NBX_REQUIRE(ch != nullptr && sc != nullptr, "NULL handle");
    NBX_REQUIRE(B >= 0, "negative batch size");
    NBX_REQUIRE(n_bisections >= 1, "num_bisections must be greater than 0");
    if (B == 0) return NBX_OK;
    NBX_REQUIRE(inside != nullptr && outside != nullptr && out != nullptr, "inside / outside / out is NULL");
    NBX_REQUIRE(inside_stride == 0 || inside_stride == ch->dof, "inside_stride must be 0 (one shared point) or dof");
    NBX_CUDA_OK(cudaSetDevice(ch->device));
    const int dof = ch->dof;
    const size_t n = (size_t)B * dof;
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t bytes = 2 * up(sizeof(double) * n) + up(sizeof(float) * n) + up((size_t)B);
    cudaMemPool_t pool = nullptr;
    int rc = nbx_workspace_pool(ch->device, &pool);
    if (rc != NBX_OK) return rc;
    struct Tmp { void* p = nullptr; cudaStream_t st = nullptr; ~Tmp() { if (p) cudaFreeAsync(p, st); } } tmp;
    tmp.st = stream;
    NBX_CUDA_OK(cudaMallocFromPoolAsync(&tmp.p, bytes, pool, stream));
    unsigned char* w = static_cast<unsigned char*>(tmp.p);
    double* lo = reinterpret_cast<double*>(w);
    double* hi = reinterpret_cast<double*>(w + up(sizeof(double) * n));
    float* mid = reinterpret_cast<float*>(w + 2 * up(sizeof(double) * n));
    uint8_t* verdict = w + 2 * up(sizeof(double) * n) + up(sizeof(float) * n);
    const unsigned grid = (unsigned)((n + 255) / 256);
    SplitWs ws;                          // one candidate queue for all rounds; the step kernel resets its counters
    // late rounds sit on the contact boundary: ~3 candidates per sample, nearly all of one kind -> 8 entries per sample and kind
    if ((rc = ws.alloc(ch->device, 4 * B, stream)) != NBX_OK) return rc;
    for (int r = 0; r <= n_bisections; ++r) {
        const int last = r == n_bisections;
        NBX_CUDA_OK(launch_pdl(bisect_step_kernel<T>, grid, 256, 0, stream, lo, hi, inside, inside_stride, outside, (const uint8_t*)verdict, mid, out,
                               B, dof, r, last, ws.args.gq_count));
        nbx_count_launch();
        NBX_CUDA_OK(cudaGetLastError());
        if (last) break;
        rc = launch_in_collision(ch, sc, mid, B, base_pose, thr, verdict, nullptr, stream, nullptr, &ws.args);
        if (rc != NBX_OK) return rc;
    }
    return NBX_OK;
}

// host-buffer variants: the batch is cut into chunks that alternate between two streams borrowed from the library's
// stream pool (nothing is created or destroyed per call), each chunk = stream-ordered staging buffer from the library's
// memory pool -> H2D -> kernels -> D2H of the verdict bytes -> free; chunk k+1's upload overlaps chunk k's kernels.
// Concurrent host threads borrow different streams and overlap.
struct HostLanes {
    static constexpr int kLanes = 2;
    int device;
    cudaStream_t st[kLanes] = {nullptr, nullptr};
    cudaMemPool_t pool = nullptr;
    explicit HostLanes(int dev) : device(dev) {}
    int open() {
        int rc = nbx_workspace_pool(device, &pool);
        if (rc != NBX_OK) return rc;
        for (int i = 0; i < kLanes; ++i)
            if ((rc = nbx_stream_acquire(device, &st[i])) != NBX_OK) return rc;
        return NBX_OK;
    }
    int finish(int rc) {
        for (int i = 0; i < kLanes; ++i) {
            if (st[i] == nullptr) continue;
            const cudaError_t e = cudaStreamSynchronize(st[i]);
            if (e != cudaSuccess && rc == NBX_OK) { nbx_set_error("cudaStreamSynchronize failed: %s", cudaGetErrorString(e)); rc = NBX_ERR_CUDA; }
            nbx_stream_release(device, st[i]);
            st[i] = nullptr;
        }
        return rc;
    }
    ~HostLanes() { finish(NBX_OK); }
};

int64_t host_chunk(int64_t n, int64_t dflt) {
    static const int64_t env = std::getenv("NBX_HOST_CHUNK") ? std::atoll(std::getenv("NBX_HOST_CHUNK")) : 0;   // tuning only
    int64_t chunk = env > 0 ? env : dflt;
    if (n < 2 * chunk) chunk = std::max<int64_t>(4096, ((n + 1) / 2 + 127) & ~(int64_t)127);
    return chunk;
}

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
    NBX_REQUIRE(q != nullptr && verdict != nullptr, "q / verdict is NULL");
    NBX_CUDA_OK(cudaSetDevice(chain->device));
    const int dof = chain->dof;
    const int64_t chunk = host_chunk(B, (int64_t)1 << 19);
    const size_t qb = (sizeof(float) * (size_t)chunk * dof + 255) & ~(size_t)255;
    HostLanes hl(chain->device);
    int rc = hl.open();
    if (rc != NBX_OK) return rc;
    int lane = 0;
    for (int64_t off = 0; off < B && rc == NBX_OK; off += chunk, lane ^= 1) {
        const int64_t n = std::min(chunk, B - off);
        cudaStream_t st = hl.st[lane];
        void* p = nullptr;
        NBX_CUDA_OK(cudaMallocFromPoolAsync(&p, qb + (size_t)chunk, hl.pool, st));
        float* dq = static_cast<float*>(p);
        uint8_t* dv = static_cast<unsigned char*>(p) + qb;
        cudaError_t e = cudaMemcpyAsync(dq, q + off * dof, sizeof(float) * (size_t)n * dof, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) {
            rc = launch_in_collision(chain, scene, dq, n, base_pose, threshold, dv, nullptr, st);
            if (rc == NBX_OK) e = cudaMemcpyAsync(verdict + off, dv, (size_t)n, cudaMemcpyDeviceToHost, st);
        }
        cudaFreeAsync(p, st);
        if (e != cudaSuccess) { nbx_set_error("cudaMemcpyAsync failed: %s", cudaGetErrorString(e)); rc = NBX_ERR_CUDA; }
    }
    return hl.finish(rc);
}

int nbx_closest_f32(const nbx_chain* chain, const nbx_scene* scene, const float* q, int64_t B, const double* base_pose,
                    float* distance, float* position_on_subject, float* position_on_target, float* normal_target_to_subject,
                    int32_t* subject_prim, int32_t* target, uint8_t* valid, float* gradient, void* stream) {
    return launch_closest(chain, scene, q, B, base_pose, distance, position_on_subject, position_on_target, normal_target_to_subject,
                          subject_prim, target, valid, gradient, static_cast<cudaStream_t>(stream));
}

int nbx_pair_distances_f32(const nbx_chain* chain, const nbx_scene* scene, const float* q, int64_t B, const double* base_pose,
                           float* distance, float* position_on_subject, float* position_on_target, float* normal_target_to_subject,
                           uint8_t* valid, void* stream) {
    return launch_pair_distances(chain, scene, q, B, base_pose, distance, position_on_subject, position_on_target,
                                 normal_target_to_subject, valid, nullptr, static_cast<cudaStream_t>(stream));
}

int nbx_pair_distances_grad_f32(const nbx_chain* chain, const nbx_scene* scene, const float* q, int64_t B, const double* base_pose,
                                float* distance, float* position_on_subject, float* position_on_target,
                                float* normal_target_to_subject, uint8_t* valid, float* gradient, void* stream) {
    return launch_pair_distances(chain, scene, q, B, base_pose, distance, position_on_subject, position_on_target,
                                 normal_target_to_subject, valid, gradient, static_cast<cudaStream_t>(stream));
}

int nbx_bisect_f32(const nbx_chain* chain, const nbx_scene* scene, const float* inside, int64_t inside_stride, const float* outside,
                   int64_t B, const double* base_pose, int n_bisections, float threshold, float* out, void* stream) {
    return launch_bisect<float>(chain, scene, inside, inside_stride, outside, B, base_pose, n_bisections, threshold, out,
                                static_cast<cudaStream_t>(stream));
}

int nbx_bisect_f64(const nbx_chain* chain, const nbx_scene* scene, const double* inside, int64_t inside_stride, const double* outside,
                   int64_t B, const double* base_pose, int n_bisections, float threshold, double* out, void* stream) {
    return launch_bisect<double>(chain, scene, inside, inside_stride, outside, B, base_pose, n_bisections, threshold, out,
                                 static_cast<cudaStream_t>(stream));
}

int nbx_edges_valid_f32(const nbx_chain* chain, const nbx_scene* scene, const float* starts, const float* goals, int64_t E,
                        const double* base_pose, double resolution, double max_distance, float threshold, uint8_t* valid,
                        int32_t* n_samples, void* stream) {
    return launch_edges(chain, scene, starts, goals, E, base_pose, resolution, max_distance, threshold, valid, n_samples, nullptr,
                        static_cast<cudaStream_t>(stream));
}

int nbx_edges_valid_detail_f32(const nbx_chain* chain, const nbx_scene* scene, const float* starts, const float* goals, int64_t E,
                               const double* base_pose, double resolution, double max_distance, float threshold, uint8_t* valid,
                               int32_t* n_samples, int32_t* first_invalid, void* stream) {
    return launch_edges(chain, scene, starts, goals, E, base_pose, resolution, max_distance, threshold, valid, n_samples, first_invalid,
                        static_cast<cudaStream_t>(stream));
}

int nbx_edges_valid_host_f32(const nbx_chain* chain, const nbx_scene* scene, const float* starts, const float* goals, int64_t E,
                             const double* base_pose, double resolution, double max_distance, float threshold, uint8_t* valid) {
    NBX_REQUIRE(chain != nullptr, "chain handle is NULL");
    if (E <= 0) return NBX_OK;
    NBX_REQUIRE(starts != nullptr && goals != nullptr && valid != nullptr, "starts / goals / valid is NULL");
    NBX_CUDA_OK(cudaSetDevice(chain->device));
    const int dof = chain->dof;
    // Chunks GROW: an edge costs the kernels ~1.5 ns and the link ~1.0 ns (56 bytes at ~55 GB/s), so only the first upload is
    // exposed and every further chunk adds the query's fixed cost (~0.1 ms: preparation launches, narrow phase, launch gaps).
    // First chunk small, each next one 2.5 x the last -- what uploads while its predecessor is being validated; 10^6 edges
    // are 4 chunks (62 k, 156 k, 390 k, the rest) instead of 8 equal ones.  NBX_HOST_CHUNK forces equal chunks (tuning).
    static const int64_t chunk_env = std::getenv("NBX_HOST_CHUNK") ? std::atoll(std::getenv("NBX_HOST_CHUNK")) : 0;
    HostLanes hl(chain->device);
    int rc = hl.open();
    if (rc != NBX_OK) return rc;
    int lane = 0;
    int64_t chunk = chunk_env > 0 ? chunk_env : std::min<int64_t>(std::max<int64_t>((int64_t)1 << 15, ((E / 16) + 127) & ~(int64_t)127), (int64_t)1 << 21);
    for (int64_t off = 0; off < E && rc == NBX_OK; lane ^= 1) {
        int64_t n = std::min(chunk, E - off);
        if (E - off - n < n / 2) n = E - off;                 // no stub at the end: fold it into this chunk
        cudaStream_t st = hl.st[lane];
        const size_t qb = (sizeof(float) * (size_t)n * dof + 255) & ~(size_t)255;
        void* p = nullptr;
        NBX_CUDA_OK(cudaMallocFromPoolAsync(&p, 2 * qb + (size_t)n, hl.pool, st));
        float* ds = static_cast<float*>(p);
        float* dg = reinterpret_cast<float*>(static_cast<unsigned char*>(p) + qb);
        uint8_t* dv = static_cast<unsigned char*>(p) + 2 * qb;
        const size_t raw = sizeof(float) * (size_t)n * dof;
        cudaError_t e = cudaMemcpyAsync(ds, starts + off * dof, raw, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(dg, goals + off * dof, raw, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) {
            rc = launch_edges(chain, scene, ds, dg, n, base_pose, resolution, max_distance, threshold, dv, nullptr, nullptr, st);
            if (rc == NBX_OK) e = cudaMemcpyAsync(valid + off, dv, (size_t)n, cudaMemcpyDeviceToHost, st);
        }
        cudaFreeAsync(p, st);
        if (e != cudaSuccess) { nbx_set_error("cudaMemcpyAsync failed: %s", cudaGetErrorString(e)); rc = NBX_ERR_CUDA; }
        off += n;
        if (chunk_env <= 0) chunk = std::min<int64_t>((chunk * 5 / 2 + 127) & ~(int64_t)127, (int64_t)1 << 21);   // staging stays bounded
    }
    return hl.finish(rc);
}

}  // extern "C"
