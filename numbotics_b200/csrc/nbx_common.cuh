// nbx_common.cuh -- internal types shared by the sm_100a kernels and the C-ABI glue.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "nbx.h"

// ---------------------------------------------------------------------------------------------
// error plumbing (no C++ exceptions cross the C-ABI)
// ---------------------------------------------------------------------------------------------
void nbx_set_error(const char* fmt, ...);
void nbx_count_launch(int n = 1);
// library-owned stream-ordered memory pool of a device (unbounded release threshold); see nbx_api.cu
int nbx_workspace_pool(int device, cudaMemPool_t* out);
// raise a kernel's dynamic shared-memory limit once per (kernel, device); see nbx_api.cu
int nbx_func_smem(const void* func, int device, size_t bytes);
#define NBX_SMEM_LIMIT(kern, bytes)                                                       \
    do {                                                                                  \
        const int _rc = nbx_func_smem(reinterpret_cast<const void*>(kern), -1, (bytes)); \
        if (_rc != NBX_OK) return _rc;                                                    \
    } while (0)
// borrow / return one of the library's non-blocking streams (the *_host_* entry points)
int nbx_stream_acquire(int device, cudaStream_t* out);
void nbx_stream_release(int device, cudaStream_t st);

#define NBX_CUDA_OK(expr)                                                                    \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            nbx_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                          __LINE__);                                                         \
            return NBX_ERR_CUDA;                                                             \
        }                                                                                    \
    } while (0)

#define NBX_REQUIRE(cond, ...)              \
    do {                                    \
        if (!(cond)) {                      \
            nbx_set_error(__VA_ARGS__);     \
            return NBX_ERR_INVALID;         \
        }                                   \
    } while (0)

// ---------------------------------------------------------------------------------------------
// kernel-parameter descriptors (passed by value as __grid_constant__ -> constant bank 0,
// so joint constants are FFMA operands, never loads)
// ---------------------------------------------------------------------------------------------
// Per movable joint, reference numbotics/robots/helpers.py:43-55 pre-folded:
//   revolute : X(q) = [ A - cos q * Bm + sin q * Cm | t ],  A = R K, Bm = R (K - I), Cm = R [a]x
//   prismatic: X(q) = [ A | t + q * d ],                     A = R,   d = R a
// with R | t the folded offset and K = a a^T.
//
// Fast path (`kind` 1..3): the axis is exactly +-e_x / +-e_y / +-e_z (every joint of a typical URDF).
// Then Rodrigues is a plane rotation of two columns:  R <- (R . R_off) with columns (k+1, k+2) rotated
// by sgn * q, and the world axis is sgn * column k of the result -- 39 instead of 54 FMAs and no
// R . a product.  `A` holds R_off in that case.  Same mathematics as helpers.py:43-55 for a unit axis.
#define NBX_KIND_GENERIC 0
#define NBX_KIND_X 1
#define NBX_KIND_Y 2
#define NBX_KIND_Z 3
#define NBX_KIND_PRISMATIC 4

template <typename T>
struct JointK {
    T A[9];
    T Bm[9];
    T Cm[9];
    T t[3];
    T ax[3];    // joint axis in the joint frame (world axis = R_after * ax)
    T d[3];     // prismatic direction R a (zero for revolute)
    T sgn;      // +-1 for the canonical-axis kinds
    int type;   // NBX_JOINT_*
    int qidx;   // column of q
    int kind;   // NBX_KIND_*
    int okind;  // structure of the offset rotation: 0 general, 1 rotation about x only (DH-style link offsets, URDF
                // origins with rpy = (r, 0, 0)), 2 identity.  Then R <- R * R_off is a plane rotation of two columns.
    T oc, os;   // cos / sin of that roll angle
};

template <typename T, int MAXJ>
struct PathK {
    int n;          // movable joints on the path
    int dof;        // q columns
    int dense;      // 1 iff n == dof and qidx[i] == i
    int use_com;
    T base[12];     // base pose
    T tail[12];     // trailing fixed offset (already multiplied by com offset when use_com)
    JointK<T> j[MAXJ];
};

struct PrimK {
    float rel[12];  // pose relative to the owning movable joint frame (tail . prim offset)
    float dims[3];
    float margin;
    int type;
    int mov;        // owning movable joint (-1 = base)
    int slot;       // float offset of the stored world pose in the per-thread scratch, -1 = not stored
    float brad;     // bounding-sphere radius of the FULL primitive (core + margin) about its centre
    float rin;      // inscribed-sphere radius of the FULL primitive about its centre (upper bounds of distance queries)
    int desc;       // index of this primitive in the chain descriptor (device order is sorted by owning joint)
    unsigned long long anc;   // bit j: movable joint j is on the path root -> this primitive (chains of <= 64 movable joints)
};

struct ObsK {
    float R[9];     // world rotation
    float c[3];     // world centre
    float dims[3];
    float margin;
    int type;
    float brad;     // bounding-sphere radius of the full obstacle
    int pad[2];     // pad[0]: 1 = axis-aligned (rotation exactly the identity); pad[1]: inscribed radius (float bits)
};

// One allowed pair, 16 bytes so the broad phase reads it with a single LDS.128.
#define NBX_WARP_TILE 32   // items per warp tile of the collision kernels: the pose scratch is [value][NBX_WARP_TILE]
struct PairK {
    uint8_t a;        // link primitive (self pairs: the one with the smaller device index)
    uint8_t b;        // link primitive (self) or obstacle (env)
    uint8_t sharp;    // 1 = ignore margins (box-box under the Bullet model)
    uint8_t both_cyl; // self pairs: both primitives are cylinders -> capsule bound available
    int32_t off_a;    // BYTE offset of a's world core in the warp's pose scratch (slot * NBX_WARP_TILE * 4)
    int32_t off_b;    // same for b (self pairs only)
    float rsum;       // self: bounding radius a + bounding radius b ; env: bounding radius of the link primitive
};
static_assert(sizeof(PairK) == 16, "PairK must stay one 16-byte record");

// Whole-tree walk used by the collision kernels.  Lives in global memory; blocks stage it in
// shared memory once (it is read with uniform addresses -> LDS broadcast).
struct TreeJointK {
    float A[9], Bm[9], Cm[9], t[3], d[3];
    int type;
    int qidx;
    int load_slot;   // -1: continue from the running transform, -2: restart from base, >=0: saved slot
    int save_slot;   // -1: do not save, >=0: save the transform after this joint
    int prim_begin, prim_end;   // primitives owned by this joint (sorted)
    int kind;        // NBX_KIND_*
    float sgn;
    float ax[3];     // joint axis in the joint frame (world axis = R_after * ax)
    int okind;       // structure of the offset rotation (JointK::okind): 0 general, 1 roll about x, 2 identity
    float oc, os;    // cos / sin of the roll
};

struct nbx_chain {
    int device;
    int sm_count;
    int dof, n_mov, n_frames, n_prims;
    // host copies (float64) used to build per-call kernel parameters
    std::vector<int> mov_parent, mov_type, mov_qidx;
    std::vector<double> mov_offset, mov_axis;
    std::vector<int> frame_mov;
    std::vector<double> frame_tail, frame_com;
    std::vector<int> prim_frame, prim_type;
    std::vector<double> prim_dims, prim_offset, prim_margin;
    std::vector<std::vector<int>> frame_path;   // movable joints root->frame
    // collision walk
    int fast_walk;                   // 1: serial chain, every movable joint a revolute about its own z with a roll-only offset
    int n_save_slots;
    int n_base_prims;                // primitives attached to the base (sorted first)
    int pose_floats;                 // per-item floats needed to park every primitive's world core
    std::vector<int> prim_order;     // device primitive index -> descriptor primitive index
    std::vector<int> prim_rank;      // descriptor primitive index -> device primitive index
    std::vector<TreeJointK> h_tree;
    std::vector<PrimK> h_prims;
    TreeJointK* d_tree;
    PrimK* d_prims;
};

struct nbx_scene {
    const nbx_chain* chain;          // identity check only (the chain may be destroyed before the scene: never dereferenced)
    int device;
    int n_obs, n_self, n_env;
    int scratch_floats;              // per-thread scratch for stored primitive poses
    std::vector<PrimK> h_prims;      // chain prims with `slot` resolved for this pair set
    std::vector<ObsK> h_obs;
    std::vector<PairK> h_self, h_env;   // self: sorted by (b, a); env: sorted by (obstacle, primitive)
    std::vector<int> h_pair_desc, h_env_begin;   // h_env_begin: [n_obs+1] ranges per obstacle; h_pair_desc: [n_self + n_env]
                                                  // descriptor pair index of each device (sorted) pair, self pairs first
    PrimK* d_prims;
    ObsK* d_obs;
    PairK* d_self;
    PairK* d_env;
    int* d_pair_desc;
    int* d_env_begin;
};

// small host helpers -------------------------------------------------------------------------
static inline void pose12_mul(const double* a, const double* b, double* out) {
    double r[12];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j)
            r[i * 4 + j] = a[i * 4 + 0] * b[0 * 4 + j] + a[i * 4 + 1] * b[1 * 4 + j] + a[i * 4 + 2] * b[2 * 4 + j];
        r[i * 4 + 3] = a[i * 4 + 0] * b[3] + a[i * 4 + 1] * b[7] + a[i * 4 + 2] * b[11] + a[i * 4 + 3];
    }
    for (int i = 0; i < 12; ++i) out[i] = r[i];
}

static inline void pose12_identity(double* o) {
    for (int i = 0; i < 12; ++i) o[i] = 0.0;
    o[0] = o[5] = o[10] = 1.0;
}

template <typename T>
void nbx_fill_joint(JointK<T>& jk, const double* off12, const double* axis, int type, int qidx);

// float paths of 9..32 joints: register-resident dense kernels (nbx_fkjac_wide_a.cu: 9..22, _b.cu: 23..32)
#define NBX_WIDE_CAP 32
#define NBX_WIDE_SPLIT 22
int nbx_launch_wide_a_f32(const PathK<float, NBX_WIDE_CAP>& P, int allz, const float* q, int64_t B, const float* lp, int64_t lps,
                          const float* gp, int64_t gps, float* T_out, float* J_out, int vec_ok, int sm_count, cudaStream_t stream);
int nbx_launch_wide_b_f32(const PathK<float, NBX_WIDE_CAP>& P, int allz, const float* q, int64_t B, const float* lp, int64_t lps,
                          const float* gp, int64_t gps, float* T_out, float* J_out, int vec_ok, int sm_count, cudaStream_t stream);

int nbx_launch_fkjac_f32(const nbx_chain* chain, int frame, const float* q, int64_t B, const double* base_pose,
                         int use_com, const float* lp, int64_t lps, const float* gp, int64_t gps, float* T_out,
                         float* J_out, cudaStream_t stream);
int nbx_launch_fkjac_f64(const nbx_chain* chain, int frame, const double* q, int64_t B, const double* base_pose,
                         int use_com, const double* lp, int64_t lps, const double* gp, int64_t gps, double* T_out,
                         double* J_out, cudaStream_t stream);
