/*
 * nbx.h -- C-ABI of the B200-native batched kinematics / collision-validity path.
 *
 * This is the drop-in boundary.  The reference
 * (landonclark97/numbotics) has no FFI layer of its own: its hot path is made of
 * plain Python callables that borrow native arithmetic from numba and PyBullet.
 * Each entry point below names the reference callable(s) it replaces
 * (paths relative to the reference checkout).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++ / torch types.
 *   - All matrices are row-major.  "pose12" = the top three rows of a 4x4 (3x4, 12 values).
 *   - Descriptors are given in float64; the library narrows them for the f32 kernels.
 *   - `*_f32` / `*_f64` entry points take DEVICE pointers (borrowed, contiguous) and
 *     launch asynchronously on `stream` (a cudaStream_t passed as void*, NULL = default
 *     stream).  No hidden synchronisation, no per-call host allocation.
 *   - `*_host_*` entry points take HOST pointers; the library pipelines
 *     H2D copy / kernel / D2H copy over internal streams and returns when the results
 *     are in the output buffers.
 *   - Every function returns 0 on success, a negative nbx_status otherwise;
 *     nbx_last_error() returns a thread-local message for the last failure.
 *   - Handles are immutable after creation: any number of host threads may launch
 *     concurrently on distinct streams (mirrors the IRIS thread-pool caller,
 *     numbotics/utils/threading.py:33-37).
 *   - There is no CPU fallback anywhere in this library.
 */
#ifndef NBX_H_
#define NBX_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NBX_VERSION 1

#if defined(__GNUC__)
#define NBX_API __attribute__((visibility("default")))
#else
#define NBX_API
#endif

typedef enum {
    NBX_OK = 0,
    NBX_ERR_INVALID = -1,   /* bad argument / descriptor */
    NBX_ERR_CUDA = -2,      /* CUDA runtime error (message has the cudaError string) */
    NBX_ERR_UNSUPPORTED = -3,
    NBX_ERR_NOMEM = -4
} nbx_status;

/* Joint types = PyBullet JOINT_* integers = numbotics/physics/constraint.py:11-15 */
#define NBX_JOINT_REVOLUTE  0
#define NBX_JOINT_PRISMATIC 1
#define NBX_JOINT_FIXED     4

/* Collision primitive types (numbotics/utils/shape.py:17-25, physics/helpers.py:231-257) */
#define NBX_PRIM_SPHERE   0   /* dims = {radius, 0, 0}                                  */
#define NBX_PRIM_BOX      1   /* dims = {hx, hy, hz} half extents                       */
#define NBX_PRIM_CYLINDER 2   /* dims = {radius, half_height, 0}, axis = local z        */

#define NBX_MAX_PATH   64     /* max movable joints between the root and a frame        */
#define NBX_MAX_PRIMS  64
#define NBX_MAX_OBS    64
#define NBX_MAX_PAIRS  1024

/*
 * Flattened kinematic tree.  Built once per Arm from the cache that
 * numbotics/robots/arm.py:20-62 builds (Arm.__init__, `_link_joint_sequence`):
 * runs of FIXED joints are folded into the offset of the next movable joint, the
 * trailing FIXED run of a frame becomes that frame's `tail`.
 *
 * Movable joints are listed in topological order (a parent always precedes its
 * children); `mov_qidx` is the q column (numbotics/physics/chain.py:516-534).
 */
typedef struct {
    int32_t dof;                 /* number of q columns                                      */
    int32_t n_mov;               /* movable (1-dof) joints                                   */
    const int32_t* mov_parent;   /* [n_mov] index of the parent movable joint, -1 = base     */
    const int32_t* mov_type;     /* [n_mov] NBX_JOINT_REVOLUTE | NBX_JOINT_PRISMATIC         */
    const int32_t* mov_qidx;     /* [n_mov] column of q                                      */
    const double*  mov_offset;   /* [n_mov][12] folded (fixed run . joint offset), pose12    */
    const double*  mov_axis;     /* [n_mov][3]  joint axis in the joint frame (NOT normalised,
                                    exactly like numbotics/robots/helpers.py:43-55)          */
    int32_t n_frames;            /* one per link of the tree (root included)                 */
    const int32_t* frame_mov;    /* [n_frames] last movable joint on the path, -1 = none     */
    const double*  frame_tail;   /* [n_frames][12] trailing fixed offset (identity if none)  */
    const double*  frame_com;    /* [n_frames][12] link inertial offset (`use_com`)          */
    int32_t n_prims;             /* link collision primitives (first <collision> per link)   */
    const int32_t* prim_frame;   /* [n_prims] owning frame                                   */
    const int32_t* prim_type;    /* [n_prims] NBX_PRIM_*                                     */
    const double*  prim_dims;    /* [n_prims][3]                                             */
    const double*  prim_offset;  /* [n_prims][12] primitive pose in the link frame           */
    const double*  prim_margin;  /* [n_prims] rounding radius (0 = sharp), see DESIGN.md     */
} nbx_chain_desc;

/*
 * World obstacles + the allowed pair lists.  Replaces the per-call set algebra of
 * numbotics/robots/arm.py:190-250 (self_collision_pairs / collision_pairs) and the
 * obstacle geometry / pose reads of numbotics/physics/object.py:200-212,353-399.
 */
typedef struct {
    int32_t n_obs;
    const int32_t* obs_type;     /* [n_obs] NBX_PRIM_*                                       */
    const double*  obs_dims;     /* [n_obs][3]                                               */
    const double*  obs_pose;     /* [n_obs][12] world pose                                   */
    const double*  obs_margin;   /* [n_obs]                                                  */
    int32_t n_self_pairs;
    const int32_t* self_pairs;   /* [n_self_pairs][2] link-primitive indices                 */
    int32_t n_env_pairs;
    const int32_t* env_pairs;    /* [n_env_pairs][2] {link primitive, obstacle}              */
    const uint8_t* self_sharp;   /* [n_self_pairs] 1 = ignore margins for this pair (box-box)*/
    const uint8_t* env_sharp;    /* [n_env_pairs]                                            */
} nbx_scene_desc;

typedef struct nbx_chain nbx_chain;
typedef struct nbx_scene nbx_scene;

NBX_API int  nbx_version(void);
NBX_API const char* nbx_last_error(void);

/* Number of kernels this library has launched in this process (bench.py `gpu_launches`). */
NBX_API int64_t nbx_launch_count(void);

NBX_API int  nbx_chain_create(const nbx_chain_desc* desc, int device, nbx_chain** out);
NBX_API void nbx_chain_destroy(nbx_chain* chain);
NBX_API int  nbx_scene_create(const nbx_chain* chain, const nbx_scene_desc* desc, nbx_scene** out);
NBX_API void nbx_scene_destroy(nbx_scene* scene);

/*
 * Fused forward kinematics + geometric Jacobian of one frame over B configurations.
 * Replaces numbotics/robots/helpers.py:91-113 (nb_compute_transformation),
 * :117-187 (nb_compute_jacobian) and the arithmetic of Arm.forward_kinematics /
 * Arm.jacobian (numbotics/robots/arm.py:394-406, 451-457).
 *
 *   q          (B, dof)      device
 *   base_pose  16 values     HOST, row-major 4x4 (Arm.base_pose), always float64
 *   local_pose (B,4,4) or (4,4) device, nullable; local_pose_stride = 16 or 0 (broadcast)
 *   global_pose same layout, nullable: replaces the end point used by the Jacobian
 *              (numbotics/robots/helpers.py:168-170); does not affect T_out
 *   T_out      (B,4,4)       device, nullable
 *   J_out      (B,6,dof)     device, nullable; rows [v; w]; columns of joints that
 *                            are not on the path are written as 0
 */
NBX_API int nbx_fk_jac_f32(const nbx_chain* chain, int frame, const float* q, int64_t B,
                   const double* base_pose, int use_com,
                   const float* local_pose, int64_t local_pose_stride,
                   const float* global_pose, int64_t global_pose_stride,
                   float* T_out, float* J_out, void* stream);
NBX_API int nbx_fk_jac_f64(const nbx_chain* chain, int frame, const double* q, int64_t B,
                   const double* base_pose, int use_com,
                   const double* local_pose, int64_t local_pose_stride,
                   const double* global_pose, int64_t global_pose_stride,
                   double* T_out, double* J_out, void* stream);

/* Same through HOST buffers (pinned or pageable); chunked, copy/compute overlapped. */
NBX_API int nbx_fk_jac_host_f32(const nbx_chain* chain, int frame, const float* q, int64_t B,
                        const double* base_pose, int use_com, float* T_out, float* J_out);
NBX_API int nbx_fk_jac_host_f64(const nbx_chain* chain, int frame, const double* q, int64_t B,
                        const double* base_pose, int use_com, double* T_out, double* J_out);

/*
 * Batched Arm.in_collision (numbotics/robots/arm.py:555-604): verdict[b] = 1 iff the
 * minimum distance over the allowed pairs is < threshold (any sign: the reference compares
 * Bullet's signed distance).  Optional min_dist[b] (nullable) receives the SIGNED minimum
 * distance: core distance minus the two margins while the cores are apart, and
 * -(margin_a + margin_b) - penetration depth of the cores (EPA) once they overlap; see DESIGN.md section 4.
 */
NBX_API int nbx_in_collision_f32(const nbx_chain* chain, const nbx_scene* scene, const float* q, int64_t B,
                         const double* base_pose, float threshold,
                         uint8_t* verdict, float* min_dist, void* stream);
NBX_API int nbx_in_collision_host_f32(const nbx_chain* chain, const nbx_scene* scene, const float* q, int64_t B,
                              const double* base_pose, float threshold, uint8_t* verdict);

/*
 * Batched DiscreteConnector.connect / .steer
 * (numbotics/planning/sampling_based/connectors.py:57-100) with the validity checker
 * `not Arm.in_collision(q, threshold)` and the default linear trajectory
 * (numbotics/planning/trajectories.py:6-21).
 *   starts, goals (E, dof) device; valid[e] = 1 iff the reference would return a state
 *   (0 = it would return None).  max_distance <= 0 selects `connect`, > 0 selects `steer`
 *   (samples stop at T_f = min(1, max_distance / distance)).
 *   n_samples (nullable, int32[E]) receives the number of interpolation samples of the edge.
 */
NBX_API int nbx_edges_valid_f32(const nbx_chain* chain, const nbx_scene* scene,
                        const float* starts, const float* goals, int64_t E,
                        const double* base_pose, double resolution, double max_distance,
                        float threshold, uint8_t* valid, int32_t* n_samples, void* stream);
/*
 * Same query with per-edge detail.  first_invalid (nullable, int32[E]) receives the index of the FIRST colliding
 * interpolation sample of the edge (0x7fffffff when the edge is valid): the sequential reference stops there
 * (connectors.py:71-73), so  useful samples = first_invalid + 1  for an invalid edge and n_samples for a valid one --
 * the "useful vs issued" accounting of the edge benchmark.  (With first_invalid the narrow phase evaluates every
 * candidate that precedes the current minimum instead of stopping at the edge's first hit.)
 */
NBX_API int nbx_edges_valid_detail_f32(const nbx_chain* chain, const nbx_scene* scene,
                               const float* starts, const float* goals, int64_t E,
                               const double* base_pose, double resolution, double max_distance,
                               float threshold, uint8_t* valid, int32_t* n_samples, int32_t* first_invalid, void* stream);
NBX_API int nbx_edges_valid_host_f32(const nbx_chain* chain, const nbx_scene* scene,
                             const float* starts, const float* goals, int64_t E,
                             const double* base_pose, double resolution, double max_distance,
                             float threshold, uint8_t* valid);

/*
 * Batched Arm.closest_to (numbotics/robots/arm.py:599-600): per configuration the allowed pair with the smallest
 * distance, as the fields of the reference's Proximity record (numbotics/physics/chain.py:952-968), plus the
 * gradient of that distance with respect to q, which is Arm.jacobian_proximity for the closest pair
 * (numbotics/robots/arm.py:621-632:  n . J_v(subject, p_subject) - n . J_v(target, p_target)).
 *   distance (B,), position_on_subject / position_on_target / normal_target_to_subject (B,3) world frame,
 *   subject_prim (B,) index of the link primitive in the chain descriptor,
 *   target (B,): obstacle index in the scene descriptor when the pair is (link, obstacle) -- then is_env bit 30 is
 *   set: target = obstacle | 0x40000000 -- else the chain-descriptor index of the other link primitive,
 *   valid (B,) u8: 0 only when no allowed pair exists; overlapping pairs carry the signed penetration distance, the
 *   push-out normal and the witness points of the EPA.  gradient (nullable, (B,dof)).
 * The scene must be unpruned (every allowed pair present).
 */
NBX_API int nbx_closest_f32(const nbx_chain* chain, const nbx_scene* scene, const float* q, int64_t B, const double* base_pose,
                    float* distance, float* position_on_subject, float* position_on_target, float* normal_target_to_subject,
                    int32_t* subject_prim, int32_t* target, uint8_t* valid, float* gradient, void* stream);

/*
 * Batched Arm.collisions (numbotics/robots/arm.py:555-580): one Proximity per ALLOWED PAIR and configuration.
 * Outputs are indexed [b][p] with p the pair's index in the scene descriptor, self pairs first
 * (p < n_self_pairs: self_pairs[p]; else env_pairs[p - n_self_pairs]); subject = the pair's first entry.
 *   distance (B, P); position_on_subject / position_on_target / normal_target_to_subject (B, P, 3), nullable;
 *   valid (B, P) u8, nullable: always 1 (kept for ABI stability; overlapping pairs report signed EPA distances).
 * The scene must be unpruned.  Arm.distance_to(q, obj, link) (arm.py:607-618) is a column of this result.
 */
NBX_API int nbx_pair_distances_f32(const nbx_chain* chain, const nbx_scene* scene, const float* q, int64_t B,
                           const double* base_pose, float* distance, float* position_on_subject,
                           float* position_on_target, float* normal_target_to_subject, uint8_t* valid, void* stream);

/*
 * Same, plus the gradient of every pair's distance with respect to q: the batched form of Arm.jacobian_proximity(q, obj, link)
 * (numbotics/robots/arm.py:621-632) for EVERY allowed pair,
 *   gradient[b][p][:] = n . J_v(subject, position_on_subject) - n . J_v(target, position_on_target)   (B, P, dof), nullable
 * (the second term only when the target is a link of this chain).  Chains of at most 64 movable joints.
 */
NBX_API int nbx_pair_distances_grad_f32(const nbx_chain* chain, const nbx_scene* scene, const float* q, int64_t B,
                                const double* base_pose, float* distance, float* position_on_subject,
                                float* position_on_target, float* normal_target_to_subject, uint8_t* valid,
                                float* gradient, void* stream);

/*
 * Batched bisection towards the collision boundary: the zero-order-hold counter-example search of IRIS
 * (numbotics/planning/safe_sets.py:124-134, IrisSolver.counter_ex_search_bisection; its caller maps it over the
 * colliding samples with a thread pool, safe_sets.py:198-199).  Per sample b the interval [inside, outside[b]]
 * (inside: one shared point, stride 0 -- the ellipsoid centre -- or one per sample, stride dof) is halved
 * n_bisections times: the half whose far end still satisfies `min distance < threshold` is kept; out[b] is the
 * colliding end.  Interval arithmetic in float64 like the reference; all buffers on the device.
 */
NBX_API int nbx_bisect_f32(const nbx_chain* chain, const nbx_scene* scene, const float* inside, int64_t inside_stride,
                   const float* outside, int64_t B, const double* base_pose, int n_bisections, float threshold,
                   float* out, void* stream);
NBX_API int nbx_bisect_f64(const nbx_chain* chain, const nbx_scene* scene, const double* inside, int64_t inside_stride,
                   const double* outside, int64_t B, const double* base_pose, int n_bisections, float threshold,
                   double* out, void* stream);

/*
 * Exact k nearest neighbours in L2 (float32, brute force): the faiss IndexFlatL2.search behind
 * ApproximateNearestNeighborsIndex.query / k_nearest (numbotics/math/geometry/nearest_neighbors.py:55-85) as the PRM loop
 * uses it (numbotics/planning/sampling_based/planners/prm.py:35-40).
 *   refs (n_refs, dof), queries (n_queries, dof) device, dof <= 64, k <= 32;
 *   causal != 0: queries and refs are the same vertex list and query i only sees refs j < i (incremental roadmap);
 *   idx (n_queries, k) int64 sorted by distance (ties: lower index first), valid (n_queries, k) u8 nullable (0 where fewer
 *   than k references exist), dist2 (n_queries, k) squared distances, nullable.
 */
NBX_API int nbx_knn_f32(const float* refs, int64_t n_refs, const float* queries, int64_t n_queries, int dof, int k,
                int causal, int64_t* idx, uint8_t* valid, float* dist2, int device, void* stream);

/*
 * Batched damped-least-squares inverse kinematics: the reference's whole iteration loop
 * (numbotics/robots/arm.py:464-552, Arm.inverse_kinematics) on the device, one problem per thread.
 *   target_pose (B,4,4) or (4,4) device; target_stride = 16 or 0 (broadcast)
 *   q0          (B,dof) device, initial guess;   q_out (B,dof) device (may alias q0)
 *   lower/upper HOST arrays [dof] (nullable unless use_limits)
 *   success     (B,) u8: |error twist| < tol at exit;  iters_out (nullable) int32 (B,)
 * Defaults of the reference: tol 1e-6, max_iter 100, max_failures 15.
 */
NBX_API int nbx_ik_dls_f64(const nbx_chain* chain, int frame, const double* target_pose, int64_t target_stride,
                   const double* q0, int64_t B, const double* base_pose, int use_com, int use_limits,
                   const double* lower, const double* upper, double tol, int max_iter, int max_failures,
                   double* q_out, uint8_t* success, int32_t* iters_out, void* stream);
NBX_API int nbx_ik_dls_f32(const nbx_chain* chain, int frame, const float* target_pose, int64_t target_stride,
                   const float* q0, int64_t B, const double* base_pose, int use_com, int use_limits,
                   const double* lower, const double* upper, double tol, int max_iter, int max_failures,
                   float* q_out, uint8_t* success, int32_t* iters_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NBX_H_ */
