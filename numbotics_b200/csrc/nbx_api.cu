// nbx_api.cu -- handle lifetime, descriptor flattening and error plumbing of the C-ABI (include/nbx.h).
#include "nbx_common.cuh"
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <algorithm>
#include <mutex>
#include <new>

namespace {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
}  // namespace

void nbx_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void nbx_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// The library's own memory pool per device, with an unbounded release threshold: the default pool hands its memory
// back to the driver at every synchronisation, which turned each call's 100+ MB workspace into a ~10 ms allocation.
int nbx_workspace_pool(int device, cudaMemPool_t* out) {
    static std::mutex mu;
    static cudaMemPool_t pools[64] = {};
    NBX_REQUIRE(device >= 0 && device < 64, "device index out of range");
    std::lock_guard<std::mutex> lock(mu);
    if (pools[device] == nullptr) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        cudaMemPool_t pool = nullptr;
        NBX_CUDA_OK(cudaMemPoolCreate(&pool, &props));
        uint64_t keep = ~0ull;
        NBX_CUDA_OK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        pools[device] = pool;
    }
    *out = pools[device];
    return NBX_OK;
}

// Dynamic shared-memory limit of a kernel: raised once per (kernel, device) instead of a cudaFuncSetAttribute driver
// call in front of every launch (round 1 paid it on each of the ~45 launches of a bisection).
int nbx_func_smem(const void* func, int device, size_t bytes) {
    struct Entry { const void* f; int bytes[64]; };
    static std::mutex mu;
    static std::vector<Entry> table;
    if (device < 0) NBX_CUDA_OK(cudaGetDevice(&device));     // the caller has already selected the chain's device
    NBX_REQUIRE(device >= 0 && device < 64, "device index out of range");
    std::lock_guard<std::mutex> lock(mu);
    Entry* e = nullptr;
    for (Entry& t : table)
        if (t.f == func) { e = &t; break; }
    if (e == nullptr) {
        table.push_back(Entry{func, {}});
        e = &table.back();
    }
    if ((int)bytes <= e->bytes[device]) return NBX_OK;
    NBX_CUDA_OK(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    e->bytes[device] = (int)bytes;
    return NBX_OK;
}

// Library-owned non-blocking streams for the *_host_* entry points: a caller borrows one for the duration of the call
// and hands it back, so concurrent host threads each get their own stream and nothing is created or destroyed per call.
namespace {
std::mutex g_stream_mu;
std::vector<cudaStream_t> g_stream_free[64];
}  // namespace

int nbx_stream_acquire(int device, cudaStream_t* out) {
    NBX_REQUIRE(device >= 0 && device < 64, "device index out of range");
    {
        std::lock_guard<std::mutex> lock(g_stream_mu);
        std::vector<cudaStream_t>& f = g_stream_free[device];
        if (!f.empty()) { *out = f.back(); f.pop_back(); return NBX_OK; }
    }
    NBX_CUDA_OK(cudaSetDevice(device));
    NBX_CUDA_OK(cudaStreamCreateWithFlags(out, cudaStreamNonBlocking));
    return NBX_OK;
}

void nbx_stream_release(int device, cudaStream_t st) {
    if (st == nullptr || device < 0 || device >= 64) return;
    std::lock_guard<std::mutex> lock(g_stream_mu);
    g_stream_free[device].push_back(st);
}



namespace {

template <typename T>
int upload(const std::vector<T>& h, T** d) {
    *d = nullptr;
    const size_t bytes = sizeof(T) * std::max<size_t>(h.size(), 1);
    NBX_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(d), bytes));
    NBX_CUDA_OK(cudaMemset(*d, 0, bytes));
    if (!h.empty()) NBX_CUDA_OK(cudaMemcpy(*d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
    return NBX_OK;
}

bool finite12(const double* p) {
    for (int i = 0; i < 12; ++i)
        if (!(p[i] == p[i]) || p[i] > 1e300 || p[i] < -1e300) return false;
    return true;
}

// bounding radius of the full (un-shrunk) primitive about its centre
float bounding_radius(int type, const double* dims) {
    if (type == NBX_PRIM_SPHERE) return (float)dims[0];
    if (type == NBX_PRIM_BOX) return (float)std::sqrt(dims[0] * dims[0] + dims[1] * dims[1] + dims[2] * dims[2]);
    return (float)std::sqrt(dims[0] * dims[0] + dims[1] * dims[1]);
}

// inscribed-sphere radius of the full primitive about its centre
float inscribed_radius(int type, const double* dims) {
    if (type == NBX_PRIM_SPHERE) return (float)dims[0];
    if (type == NBX_PRIM_BOX) return (float)std::min(dims[0], std::min(dims[1], dims[2]));
    return (float)std::min(dims[0], dims[1]);
}

int prim_pose_floats(int type) { return type == NBX_PRIM_SPHERE ? 3 : (type == NBX_PRIM_CYLINDER ? 6 : 12); }

// core dims = full dims shrunk by the rounding margin (Bullet keeps "implicit shape dimensions" the same way)
int core_dims(int type, const double* dims, double margin, float* out, float* out_margin) {
    if (type == NBX_PRIM_SPHERE) {
        NBX_REQUIRE(dims[0] > 0.0, "sphere radius must be positive");
        out[0] = out[1] = out[2] = 0.f;
        *out_margin = (float)dims[0];   // a sphere is a point core swept by its radius
        return NBX_OK;
    }
    NBX_REQUIRE(margin >= 0.0, "negative margin");
    if (type == NBX_PRIM_BOX) {
        for (int k = 0; k < 3; ++k) {
            NBX_REQUIRE(dims[k] - margin >= 0.0, "box half extent %g smaller than its margin %g", dims[k], margin);
            out[k] = (float)(dims[k] - margin);
        }
    } else if (type == NBX_PRIM_CYLINDER) {
        NBX_REQUIRE(dims[0] - margin >= 0.0 && dims[1] - margin >= 0.0, "cylinder smaller than its margin");
        out[0] = (float)(dims[0] - margin);
        out[1] = (float)(dims[1] - margin);
        out[2] = 0.f;
    } else {
        nbx_set_error("unsupported primitive type %d", type);
        return NBX_ERR_UNSUPPORTED;
    }
    *out_margin = (float)margin;
    return NBX_OK;
}

}  // namespace

extern "C" {

int nbx_version(void) { return NBX_VERSION; }
const char* nbx_last_error(void) { return g_err; }
int64_t nbx_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }

int nbx_chain_create(const nbx_chain_desc* d, int device, nbx_chain** out) {
    NBX_REQUIRE(d != nullptr && out != nullptr, "NULL argument");
    *out = nullptr;
    NBX_REQUIRE(d->dof >= 0 && d->n_mov >= 0 && d->n_frames >= 1 && d->n_prims >= 0, "negative size in chain descriptor");
    NBX_REQUIRE(d->n_prims <= NBX_MAX_PRIMS, "too many link primitives (%d > %d)", d->n_prims, NBX_MAX_PRIMS);
    int ndev = 0;
    NBX_CUDA_OK(cudaGetDeviceCount(&ndev));
    NBX_REQUIRE(device >= 0 && device < ndev, "CUDA device %d not available (%d visible)", device, ndev);
    NBX_CUDA_OK(cudaSetDevice(device));

    nbx_chain* ch = new (std::nothrow) nbx_chain();
    if (!ch) { nbx_set_error("out of host memory"); return NBX_ERR_NOMEM; }
    struct Guard { nbx_chain* c; ~Guard() { if (c) nbx_chain_destroy(c); } } guard{ch};
    ch->device = device;
    {
        cudaDeviceProp prop;
        NBX_CUDA_OK(cudaGetDeviceProperties(&prop, device));
        ch->sm_count = prop.multiProcessorCount;
    }
    ch->dof = d->dof; ch->n_mov = d->n_mov; ch->n_frames = d->n_frames; ch->n_prims = d->n_prims;
    ch->d_tree = nullptr; ch->d_prims = nullptr;
    ch->mov_parent.assign(d->mov_parent, d->mov_parent + d->n_mov);
    ch->mov_type.assign(d->mov_type, d->mov_type + d->n_mov);
    ch->mov_qidx.assign(d->mov_qidx, d->mov_qidx + d->n_mov);
    ch->mov_offset.assign(d->mov_offset, d->mov_offset + 12 * (size_t)d->n_mov);
    ch->mov_axis.assign(d->mov_axis, d->mov_axis + 3 * (size_t)d->n_mov);
    ch->frame_mov.assign(d->frame_mov, d->frame_mov + d->n_frames);
    ch->frame_tail.assign(d->frame_tail, d->frame_tail + 12 * (size_t)d->n_frames);
    ch->frame_com.assign(d->frame_com, d->frame_com + 12 * (size_t)d->n_frames);
    ch->prim_frame.assign(d->prim_frame, d->prim_frame + d->n_prims);
    ch->prim_type.assign(d->prim_type, d->prim_type + d->n_prims);
    ch->prim_dims.assign(d->prim_dims, d->prim_dims + 3 * (size_t)d->n_prims);
    ch->prim_offset.assign(d->prim_offset, d->prim_offset + 12 * (size_t)d->n_prims);
    ch->prim_margin.assign(d->prim_margin, d->prim_margin + d->n_prims);

    for (int j = 0; j < ch->n_mov; ++j) {
        NBX_REQUIRE(ch->mov_parent[j] >= -1 && ch->mov_parent[j] < j, "movable joints must be in topological order (joint %d)", j);
        NBX_REQUIRE(ch->mov_type[j] == NBX_JOINT_REVOLUTE || ch->mov_type[j] == NBX_JOINT_PRISMATIC,
                    "joint %d: only REVOLUTE and PRISMATIC joints are movable on this path (SPHERICAL is broken in the reference's batched path too)", j);
        NBX_REQUIRE(ch->mov_qidx[j] >= 0 && ch->mov_qidx[j] < ch->dof, "joint %d: q index out of range", j);
        NBX_REQUIRE(finite12(&ch->mov_offset[12 * j]), "joint %d: non-finite offset", j);
    }
    ch->frame_path.resize(ch->n_frames);
    for (int f = 0; f < ch->n_frames; ++f) {
        NBX_REQUIRE(ch->frame_mov[f] >= -1 && ch->frame_mov[f] < ch->n_mov, "frame %d: bad movable joint index", f);
        std::vector<int> path;
        for (int j = ch->frame_mov[f]; j >= 0; j = ch->mov_parent[j]) path.push_back(j);
        std::reverse(path.begin(), path.end());
        NBX_REQUIRE((int)path.size() <= NBX_MAX_PATH, "frame %d: path longer than %d movable joints", f, NBX_MAX_PATH);
        ch->frame_path[f] = path;
    }

    // ---- collision walk: primitives sorted by owning joint, branch points saved ------------
    std::vector<int> prim_mov(ch->n_prims);
    for (int p = 0; p < ch->n_prims; ++p) {
        NBX_REQUIRE(ch->prim_frame[p] >= 0 && ch->prim_frame[p] < ch->n_frames, "primitive %d: bad frame", p);
        prim_mov[p] = ch->frame_mov[ch->prim_frame[p]];
    }
    ch->prim_order.resize(ch->n_prims);
    for (int p = 0; p < ch->n_prims; ++p) ch->prim_order[p] = p;
    std::stable_sort(ch->prim_order.begin(), ch->prim_order.end(), [&](int a, int b) { return prim_mov[a] < prim_mov[b]; });
    ch->prim_rank.assign(ch->n_prims, 0);
    for (int i = 0; i < ch->n_prims; ++i) ch->prim_rank[ch->prim_order[i]] = i;
    ch->n_base_prims = 0;
    ch->pose_floats = 0;
    ch->h_prims.resize(ch->n_prims);
    for (int i = 0; i < ch->n_prims; ++i) {
        const int p = ch->prim_order[i];
        PrimK& pk = ch->h_prims[i];
        std::memset(&pk, 0, sizeof(pk));
        double rel[12];
        pose12_mul(&ch->frame_tail[12 * ch->prim_frame[p]], &ch->prim_offset[12 * p], rel);
        for (int k = 0; k < 12; ++k) pk.rel[k] = (float)rel[k];
        pk.type = ch->prim_type[p];
        int rc = core_dims(pk.type, &ch->prim_dims[3 * p], ch->prim_margin[p], pk.dims, &pk.margin);
        if (rc != NBX_OK) return rc;
        pk.mov = prim_mov[p];
        pk.slot = ch->pose_floats;
        ch->pose_floats += prim_pose_floats(pk.type);
        pk.brad = bounding_radius(pk.type, &ch->prim_dims[3 * p]) * 1.000001f;
        pk.rin = inscribed_radius(pk.type, &ch->prim_dims[3 * p]) * 0.999999f;
        pk.desc = p;
        pk.anc = 0ull;
        if (ch->n_mov <= 64)
            for (int j = pk.mov; j >= 0; j = ch->mov_parent[j]) pk.anc |= 1ull << j;
        if (pk.mov < 0) ch->n_base_prims++;
    }
    ch->fast_walk = ch->n_mov > 0 ? 1 : 0;
    ch->h_tree.resize(ch->n_mov);
    std::vector<int> save_slot(ch->n_mov, -1);
    int n_save = 0;
    for (int j = 0; j < ch->n_mov; ++j) {
        const int pj = ch->mov_parent[j];
        if (pj >= 0 && pj != j - 1 && save_slot[pj] < 0) save_slot[pj] = n_save++;
    }
    NBX_REQUIRE(n_save <= 8, "tree has %d branch points below movable joints (max 8)", n_save);
    ch->n_save_slots = n_save;
    for (int j = 0; j < ch->n_mov; ++j) {
        TreeJointK& tj = ch->h_tree[j];
        std::memset(&tj, 0, sizeof(tj));
        JointK<float> jk;
        nbx_fill_joint<float>(jk, &ch->mov_offset[12 * j], &ch->mov_axis[3 * j], ch->mov_type[j], ch->mov_qidx[j]);
        for (int k = 0; k < 9; ++k) { tj.A[k] = jk.A[k]; tj.Bm[k] = jk.Bm[k]; tj.Cm[k] = jk.Cm[k]; }
        for (int k = 0; k < 3; ++k) { tj.t[k] = jk.t[k]; tj.d[k] = jk.d[k]; tj.ax[k] = jk.ax[k]; }
        tj.type = jk.type;
        tj.qidx = jk.qidx;
        tj.kind = jk.kind;
        tj.sgn = jk.sgn;
        tj.okind = jk.okind; tj.oc = jk.oc; tj.os = jk.os;
        if (jk.kind != NBX_KIND_Z || jk.okind == 0 || ch->mov_parent[j] != j - 1) ch->fast_walk = 0;
        const int pj = ch->mov_parent[j];
        tj.load_slot = (pj == j - 1) ? -1 : (pj < 0 ? -2 : save_slot[pj]);
        tj.save_slot = save_slot[j];
        tj.prim_begin = tj.prim_end = 0;
    }
    {
        int i = ch->n_base_prims;
        for (int j = 0; j < ch->n_mov; ++j) {
            ch->h_tree[j].prim_begin = i;
            while (i < ch->n_prims && ch->h_prims[i].mov == j) ++i;
            ch->h_tree[j].prim_end = i;
        }
    }
    int rc = upload(ch->h_tree, &ch->d_tree);
    if (rc != NBX_OK) return rc;
    rc = upload(ch->h_prims, &ch->d_prims);
    if (rc != NBX_OK) return rc;
    guard.c = nullptr;
    *out = ch;
    return NBX_OK;
}

void nbx_chain_destroy(nbx_chain* ch) {
    if (!ch) return;
    cudaSetDevice(ch->device);
    if (ch->d_tree) cudaFree(ch->d_tree);
    if (ch->d_prims) cudaFree(ch->d_prims);
    delete ch;
}

int nbx_scene_create(const nbx_chain* ch, const nbx_scene_desc* d, nbx_scene** out) {
    NBX_REQUIRE(ch != nullptr && d != nullptr && out != nullptr, "NULL argument");
    *out = nullptr;
    NBX_REQUIRE(d->n_obs >= 0 && d->n_obs <= NBX_MAX_OBS, "obstacle count %d out of range [0, %d]", d->n_obs, NBX_MAX_OBS);
    NBX_REQUIRE(d->n_self_pairs >= 0 && d->n_env_pairs >= 0 && d->n_self_pairs + d->n_env_pairs <= NBX_MAX_PAIRS,
                "pair count out of range (max %d)", NBX_MAX_PAIRS);
    NBX_CUDA_OK(cudaSetDevice(ch->device));
    nbx_scene* sc = new (std::nothrow) nbx_scene();
    if (!sc) { nbx_set_error("out of host memory"); return NBX_ERR_NOMEM; }
    struct Guard { nbx_scene* s; ~Guard() { if (s) nbx_scene_destroy(s); } } guard{sc};
    sc->chain = ch;
    sc->device = ch->device;
    sc->n_obs = d->n_obs; sc->n_self = d->n_self_pairs; sc->n_env = d->n_env_pairs;
    sc->d_prims = nullptr; sc->d_obs = nullptr; sc->d_self = nullptr; sc->d_env = nullptr;
    sc->d_pair_desc = nullptr; sc->d_env_begin = nullptr;
    sc->h_prims = ch->h_prims;

    sc->h_obs.resize(d->n_obs);
    for (int o = 0; o < d->n_obs; ++o) {
        ObsK& ok = sc->h_obs[o];
        std::memset(&ok, 0, sizeof(ok));
        const double* P = &d->obs_pose[12 * o];
        NBX_REQUIRE(finite12(P), "obstacle %d: non-finite pose", o);
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) ok.R[r * 3 + c] = (float)P[r * 4 + c];
            ok.c[r] = (float)P[r * 4 + 3];
        }
        ok.type = d->obs_type[o];
        int rc = core_dims(ok.type, &d->obs_dims[3 * o], d->obs_margin[o], ok.dims, &ok.margin);
        if (rc != NBX_OK) return rc;
        ok.brad = bounding_radius(ok.type, &d->obs_dims[3 * o]) * 1.000001f;
        {
            const float rin = inscribed_radius(ok.type, &d->obs_dims[3 * o]) * 0.999999f;
            std::memcpy(&ok.pad[1], &rin, sizeof(float));
        }
        // axis-aligned (rotation exactly the identity): the broad phase skips the rotation into the obstacle frame
        ok.pad[0] = (P[0] == 1.0 && P[5] == 1.0 && P[10] == 1.0 && P[1] == 0.0 && P[2] == 0.0 && P[4] == 0.0 && P[6] == 0.0 &&
                     P[8] == 0.0 && P[9] == 0.0) ? 1 : 0;
    }
    const int np = ch->n_prims;
    std::vector<PairK> self(d->n_self_pairs);
    for (int i = 0; i < d->n_self_pairs; ++i) {
        int a = d->self_pairs[2 * i], b = d->self_pairs[2 * i + 1];
        NBX_REQUIRE(a >= 0 && a < np && b >= 0 && b < np && a != b, "self pair %d: bad primitive index", i);
        a = ch->prim_rank[a]; b = ch->prim_rank[b];
        if (a > b) std::swap(a, b);
        PairK& pr = self[i];
        std::memset(&pr, 0, sizeof(PairK));
        pr.a = (uint8_t)a; pr.b = (uint8_t)b;
        pr.sharp = d->self_sharp ? d->self_sharp[i] : 0;
        const PrimK& pa = sc->h_prims[a];
        const PrimK& pb = sc->h_prims[b];
        pr.both_cyl = (pa.type == NBX_PRIM_CYLINDER && pb.type == NBX_PRIM_CYLINDER) ? 1 : 0;
        pr.off_a = pa.slot * NBX_WARP_TILE * 4; pr.off_b = pb.slot * NBX_WARP_TILE * 4;
        pr.rsum = pa.brad + pb.brad;
    }
    // device order (sorted) -> descriptor order, for the queries that report per pair (nbx_pair_distances_f32)
    auto sorted_with_origin = [](std::vector<PairK>& v, std::vector<int>& origin, int base) {
        std::vector<int> idx(v.size());
        for (size_t i = 0; i < idx.size(); ++i) idx[i] = (int)i;
        std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return v[x].b != v[y].b ? v[x].b < v[y].b : v[x].a < v[y].a; });
        std::vector<PairK> out(v.size());
        for (size_t i = 0; i < idx.size(); ++i) { out[i] = v[idx[i]]; origin.push_back(base + idx[i]); }
        v.swap(out);
    };
    sc->h_pair_desc.clear();            // [n_self + n_env]: descriptor pair index of each device pair
    sorted_with_origin(self, sc->h_pair_desc, 0);
    std::vector<PairK> env(d->n_env_pairs);
    for (int i = 0; i < d->n_env_pairs; ++i) {
        const int a = d->env_pairs[2 * i], o = d->env_pairs[2 * i + 1];
        NBX_REQUIRE(a >= 0 && a < np && o >= 0 && o < d->n_obs, "env pair %d: bad index", i);
        PairK& pr = env[i];
        std::memset(&pr, 0, sizeof(PairK));
        pr.a = (uint8_t)ch->prim_rank[a]; pr.b = (uint8_t)o;
        pr.sharp = d->env_sharp ? d->env_sharp[i] : 0;
        const PrimK& pa = sc->h_prims[pr.a];
        pr.off_a = pa.slot * NBX_WARP_TILE * 4;
        pr.rsum = pa.brad;
    }
    // obstacle-major: the broad phase keeps one obstacle in registers and streams the primitives past it
    sorted_with_origin(env, sc->h_pair_desc, d->n_self_pairs);
    sc->h_self = self;
    sc->h_env = env;
    sc->h_env_begin.assign(d->n_obs + 1, 0);
    for (const PairK& p : env) sc->h_env_begin[p.b + 1]++;
    for (int i = 0; i < d->n_obs; ++i) sc->h_env_begin[i + 1] += sc->h_env_begin[i];
    sc->scratch_floats = ch->pose_floats;

    int rc;
    if ((rc = upload(sc->h_prims, &sc->d_prims)) != NBX_OK) return rc;
    if ((rc = upload(sc->h_obs, &sc->d_obs)) != NBX_OK) return rc;
    if ((rc = upload(sc->h_self, &sc->d_self)) != NBX_OK) return rc;
    if ((rc = upload(sc->h_env, &sc->d_env)) != NBX_OK) return rc;
    if ((rc = upload(sc->h_pair_desc, &sc->d_pair_desc)) != NBX_OK) return rc;
    if ((rc = upload(sc->h_env_begin, &sc->d_env_begin)) != NBX_OK) return rc;
    guard.s = nullptr;
    *out = sc;
    return NBX_OK;
}

void nbx_scene_destroy(nbx_scene* sc) {
    if (!sc) return;
    // the scene's OWN copy of the device index: Python may finalise the chain handle first, and reading sc->chain->device
    // from a freed chain handed cudaSetDevice a garbage ordinal ("invalid device ordinal" surfacing at the next launch)
    cudaSetDevice(sc->device);
    if (sc->d_prims) cudaFree(sc->d_prims);
    if (sc->d_obs) cudaFree(sc->d_obs);
    if (sc->d_self) cudaFree(sc->d_self);
    if (sc->d_env) cudaFree(sc->d_env);
    if (sc->d_pair_desc) cudaFree(sc->d_pair_desc);
    if (sc->d_env_begin) cudaFree(sc->d_env_begin);
    delete sc;
}

}  // extern "C"
