"""Float64 CPU restatement of Arm.in_collision / DiscreteConnector.  TEST INFRASTRUCTURE ONLY.

Two independent routes to the same numbers:

* checker route (`min_distance`, `in_collision`, `edges_valid`): link frames from the NumPy FK
  oracle (oracle/fkjac.py, pinned against the reference's numba kernels), primitive world pose
  = link frame . collision offset (reference physics/helpers.py:231-257, chain.py:423-445),
  pair distances from the C GJK (`oracle_pair_distances`), pair lists straight from the
  Arm's Python sets.  This route does not touch numbotics_b200's flattened descriptors.
* pipeline route (`c_in_collision`, `c_edges_valid`, `c_fk_jac`): everything in C on the flattened
  descriptor, OpenMP over configurations.  Used for the CPU baseline timing; tests check it
  against the checker route.

The narrow phase is PARITY UNPINNED against Bullet (see nbx_oracle.c header).
"""
import ctypes as C

import numpy as np

from . import build as _build
from . import fkjac

PRIM_SPHERE, PRIM_BOX, PRIM_CYL = 0, 1, 2
_lib = None

_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)


class OChain(C.Structure):
    _fields_ = [("dof", C.c_int), ("n_mov", C.c_int), ("n_frames", C.c_int), ("n_prims", C.c_int),
                ("mov_parent", _i32p), ("mov_type", _i32p), ("mov_qidx", _i32p),
                ("mov_offset", _f64p), ("mov_axis", _f64p),
                ("frame_mov", _i32p), ("frame_tail", _f64p), ("frame_com", _f64p),
                ("prim_frame", _i32p), ("prim_type", _i32p),
                ("prim_dims", _f64p), ("prim_offset", _f64p), ("prim_margin", _f64p)]


class OScene(C.Structure):
    _fields_ = [("n_obs", C.c_int), ("obs_type", _i32p), ("obs_dims", _f64p), ("obs_pose", _f64p), ("obs_margin", _f64p),
                ("n_self", C.c_int), ("self_pairs", _i32p), ("n_env", C.c_int), ("env_pairs", _i32p),
                ("self_sharp", _u8p), ("env_sharp", _u8p)]


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build())
        _lib.oracle_num_threads.restype = C.c_int
    return _lib


def _arr(a, dt):
    a = np.ascontiguousarray(a, dtype=dt)
    if a.size == 0:
        a = np.zeros(2, dtype=dt)
    return a


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


class CDesc:
    """ctypes views of the flattened chain / scene arrays (keeps the buffers alive)."""

    def __init__(self, chain_arrays, scene_arrays=None):
        A = chain_arrays
        self.keep = []
        ch = OChain()
        ch.dof = int(A["dof"]); ch.n_mov = len(A["mov_type"]); ch.n_frames = len(A["frame_mov"]); ch.n_prims = len(A["prim_type"])
        for k, ct, dt in (("mov_parent", C.c_int32, np.int32), ("mov_type", C.c_int32, np.int32), ("mov_qidx", C.c_int32, np.int32),
                          ("mov_offset", C.c_double, np.float64), ("mov_axis", C.c_double, np.float64),
                          ("frame_mov", C.c_int32, np.int32), ("frame_tail", C.c_double, np.float64), ("frame_com", C.c_double, np.float64),
                          ("prim_frame", C.c_int32, np.int32), ("prim_type", C.c_int32, np.int32),
                          ("prim_dims", C.c_double, np.float64), ("prim_offset", C.c_double, np.float64), ("prim_margin", C.c_double, np.float64)):
            a = _arr(A[k], dt); self.keep.append(a); setattr(ch, k, _p(a, ct))
        self.chain = ch
        self.dof = ch.dof
        self.scene = None
        if scene_arrays is not None:
            S = scene_arrays
            sc = OScene()
            sc.n_obs = len(S["obs_type"]); sc.n_self = len(S["self_pairs"]); sc.n_env = len(S["env_pairs"])
            for k, ct, dt in (("obs_type", C.c_int32, np.int32), ("obs_dims", C.c_double, np.float64), ("obs_pose", C.c_double, np.float64),
                              ("obs_margin", C.c_double, np.float64), ("self_pairs", C.c_int32, np.int32), ("env_pairs", C.c_int32, np.int32),
                              ("self_sharp", C.c_uint8, np.uint8), ("env_sharp", C.c_uint8, np.uint8)):
                a = _arr(S[k], dt); self.keep.append(a); setattr(sc, k, _p(a, ct))
            self.scene = sc


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def set_threads(n: int):
    lib().oracle_set_threads(C.c_int(int(n)))


# ------------------------------------------------------------------------------------------------
# pipeline route (C, flattened descriptor)
# ------------------------------------------------------------------------------------------------
def c_fk_jac(desc: CDesc, frame_id: int, q, base_pose, use_com=False, want_T=True, want_J=True):
    q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, desc.dof)
    n = q.shape[0]
    base = np.ascontiguousarray(base_pose, dtype=np.float64)
    T = np.empty((n, 4, 4)) if want_T else None
    J = np.empty((n, 6, desc.dof)) if want_J else None
    rc = lib().oracle_fk_jac(C.byref(desc.chain), C.c_int(frame_id), _p(q, C.c_double), C.c_int64(n), _p(base, C.c_double),
                             C.c_int(int(use_com)), _p(T, C.c_double) if want_T else None, _p(J, C.c_double) if want_J else None)
    assert rc == 0
    return T, J


def c_in_collision(desc: CDesc, q, base_pose, threshold=0.0, want_dist=False):
    q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, desc.dof)
    n = q.shape[0]
    base = np.ascontiguousarray(base_pose, dtype=np.float64)
    v = np.empty(n, np.uint8)
    d = np.empty(n, np.float64) if want_dist else None
    rc = lib().oracle_in_collision(C.byref(desc.chain), C.byref(desc.scene), _p(q, C.c_double), C.c_int64(n), _p(base, C.c_double),
                                   C.c_double(threshold), _p(v, C.c_uint8), _p(d, C.c_double) if want_dist else None)
    assert rc == 0
    return (v.astype(bool), d) if want_dist else v.astype(bool)


def c_edges_valid(desc: CDesc, starts, goals, base_pose, resolution, max_distance=None, threshold=0.0, want_min_d=False):
    """(valid, n_samples, n_useful[, min_d]); `want_min_d` evaluates every sample of every edge (no early exit) and adds
    the smallest min-distance over the edge's samples (what decides whether an edge is inside the contact band)."""
    s = np.ascontiguousarray(starts, dtype=np.float64).reshape(-1, desc.dof)
    g = np.ascontiguousarray(goals, dtype=np.float64).reshape(-1, desc.dof)
    E = s.shape[0]
    base = np.ascontiguousarray(base_pose, dtype=np.float64)
    v = np.empty(E, np.uint8); ns = np.empty(E, np.int32); nu = np.empty(E, np.int32)
    md = np.empty(E, np.float64) if want_min_d else None
    rc = lib().oracle_edges_valid_ex(C.byref(desc.chain), C.byref(desc.scene), _p(s, C.c_double), _p(g, C.c_double), C.c_int64(E),
                                     _p(base, C.c_double), C.c_double(resolution),
                                     C.c_double(0.0 if max_distance is None else max_distance), C.c_double(threshold),
                                     _p(v, C.c_uint8), _p(ns, C.c_int32), _p(nu, C.c_int32), _p(md, C.c_double) if want_min_d else None)
    assert rc == 0
    return (v.astype(bool), ns, nu, md) if want_min_d else (v.astype(bool), ns, nu)


# ------------------------------------------------------------------------------------------------
# checker route (NumPy FK oracle + C narrow phase, pair lists from the Arm's Python sets)
# ------------------------------------------------------------------------------------------------
def bullet_margin(shape_name: str, info: dict, margin_model: str):
    """Rounding radius of a primitive: btConvexInternalShape::setSafeMargin as remembered --
    min(CONVEX_DISTANCE_MARGIN = 0.04, 0.1 * smallest half extent).  UNPINNED (no Bullet here)."""
    if margin_model == "sharp":
        return 0.0
    if shape_name in ("CUBE", "CUBOID"):
        return min(0.04, 0.1 * float(np.min(info["half_extents"])))
    if shape_name == "CYLINDER":
        return min(0.04, 0.1 * min(float(info["radius"]), 0.5 * float(info["height"])))
    return 0.0


def _shape_params(cshape, margin_model):
    name = cshape.shape.name
    info = cshape._shape_info
    if name in ("CUBE", "CUBOID"):
        return PRIM_BOX, np.asarray(info["half_extents"], float), bullet_margin(name, info, margin_model)
    if name == "SPHERE":
        return PRIM_SPHERE, np.array([float(info["radius"]), 0.0, 0.0]), 0.0
    if name == "CYLINDER":
        return PRIM_CYL, np.array([float(info["radius"]), 0.5 * float(info["height"]), 0.0]), bullet_margin(name, info, margin_model)
    raise NotImplementedError(name)


class SceneOracle:
    """Built from an Arm-like object exposing `_link_joint_sequence`, `_links_from_nodes`, `dof`, `base_pose`,
    `self_collision_pairs()`, `collision_pairs()` -- only the Python-level model, no device descriptors."""

    def __init__(self, arm, margin_model="bullet", foreign_link_poses=None):
        """`foreign_link_poses`: {Link of ANOTHER chain: 4x4 world pose} -- such links are static targets (reference
        arm.py:239-241); the caller computes their poses independently (e.g. with an ArmOracle of that chain)."""
        self.fk = fkjac.ArmOracle(arm._link_joint_sequence, arm.dof, arm.base_pose)
        self.dof = arm.dof
        foreign = foreign_link_poses or {}
        node_of_link = {link: node for node, link in arm._links_from_nodes.items()}
        self_pairs = arm.self_collision_pairs()
        env_pairs = [p for p in arm.collision_pairs().difference(self_pairs)
                     if (not hasattr(p[1], "_index") or p[1] in foreign) and p[1]._collision_shape.shape.name != "EMPTY"]
        links = []
        for a, b in self_pairs:
            for l in (a, b):
                if l not in links:
                    links.append(l)
        for a, _ in env_pairs:
            if a not in links:
                links.append(a)
        links = [l for l in links if l._collision_shape.shape.name != "EMPTY"]
        objs = []
        for _, o in env_pairs:
            if o not in objs:
                objs.append(o)
        self.link_nodes = [node_of_link[l] for l in links]
        self.link_offsets = [np.asarray(l._collision_shape.offset, float) for l in links]
        types, dims, margins = [], [], []
        for s in [l._collision_shape for l in links] + [o._collision_shape for o in objs]:
            t, d, m = _shape_params(s, margin_model)
            types.append(t); dims.append(d); margins.append(m)
        self.types = np.asarray(types, np.int32)
        self.dims = np.asarray(dims, np.float64).reshape(-1, 3)
        self.margins = np.asarray(margins, np.float64)
        self.static_poses = np.asarray([((foreign[o] if o in foreign else o.pose) @ o._collision_shape.offset)[:3, :4].reshape(12)
                                        for o in objs], np.float64).reshape(-1, 12)
        li = {l: i for i, l in enumerate(links)}
        oi = {o: len(links) + i for i, o in enumerate(objs)}
        pairs = [(li[a], li[b]) for a, b in self_pairs if a in li and b in li]
        pairs += [(li[a], oi[o]) for a, o in env_pairs if a in li]
        self.pairs = np.asarray(pairs, np.int32).reshape(-1, 2)
        self.sharp = np.array([self.types[a] == PRIM_BOX and self.types[b] == PRIM_BOX for a, b in pairs], np.uint8)
        self.n_link = len(links)
        self.links, self.objs = links, objs

    def pair_distances(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, self.dof)
        n = q.shape[0]
        poses = np.empty((n, max(self.n_link, 1), 12))
        for i, (node, off) in enumerate(zip(self.link_nodes, self.link_offsets)):
            poses[:, i] = (self.fk.forward_kinematics(q, node) @ off)[:, :3, :4].reshape(n, 12)
        out = np.empty((n, len(self.pairs)))
        sp = _arr(self.static_poses, np.float64)
        rc = lib().oracle_pair_distances(C.c_int64(n), C.c_int(self.n_link), _p(poses, C.c_double), C.c_int(len(self.static_poses)),
                                         _p(sp, C.c_double), _p(self.types, C.c_int32), _p(np.ascontiguousarray(self.dims), C.c_double),
                                         _p(self.margins, C.c_double), C.c_int(len(self.pairs)), _p(_arr(self.pairs, np.int32), C.c_int32),
                                         _p(_arr(self.sharp, np.uint8), C.c_uint8), _p(out, C.c_double))
        assert rc == 0
        return out

    def link_poses(self, q):
        """World poses (n, n_link, 12) of the link primitives."""
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, self.dof)
        poses = np.empty((q.shape[0], max(self.n_link, 1), 12))
        for i, (node, off) in enumerate(zip(self.link_nodes, self.link_offsets)):
            poses[:, i] = (self.fk.forward_kinematics(q, node) @ off)[:, :3, :4].reshape(-1, 12)
        return poses

    def point_to_shape(self, shape_index, shape_poses12, points, sharp=False):
        """Distance from world points (n,3) to shape `shape_index` (link or obstacle index space of `pairs`) posed at
        `shape_poses12 (n,12)`, through the same C narrow phase (a point is a sphere of radius 0).  `sharp`: the shape
        without its rounding margin (box-box pairs under the Bullet model)."""
        n = points.shape[0]
        poses = np.zeros((n, 2, 12))
        poses[:, 0] = shape_poses12
        poses[:, 1, [0, 5, 10]] = 1.0
        poses[:, 1, [3, 7, 11]] = points
        types = _arr([self.types[shape_index], PRIM_SPHERE], np.int32)
        dims = np.ascontiguousarray(np.vstack([self.dims[shape_index], np.zeros(3)]), dtype=np.float64)
        margins = _arr([self.margins[shape_index], 0.0], np.float64)
        pairs = _arr([[0, 1]], np.int32)
        out = np.empty((n, 1))
        sp = np.zeros((1, 12))
        rc = lib().oracle_pair_distances(C.c_int64(n), C.c_int(2), _p(np.ascontiguousarray(poses), C.c_double), C.c_int(0),
                                         _p(sp, C.c_double), _p(types, C.c_int32), _p(dims, C.c_double), _p(margins, C.c_double),
                                         C.c_int(1), _p(pairs, C.c_int32), _p(_arr([1 if sharp else 0], np.uint8), C.c_uint8), _p(out, C.c_double))
        assert rc == 0
        return out[:, 0]

    def min_distance(self, q):
        shape = np.shape(q)[:-1]
        return self.pair_distances(q).min(axis=1).reshape(shape)

    def in_collision(self, q, threshold=0.0):
        return self.min_distance(q) < threshold

    def edges_valid(self, starts, goals, resolution, max_distance=None, threshold=0.0, return_detail=False):
        """Literal restatement of connectors.py:57-100 (NumPy arange grid, sequential semantics)."""
        starts = np.asarray(starts, np.float64).reshape(-1, self.dof)
        goals = np.asarray(goals, np.float64).reshape(-1, self.dof)
        E = starts.shape[0]
        valid = np.zeros(E, bool)
        n_samples = np.zeros(E, np.int32)
        min_d = np.full(E, np.inf)
        for e in range(E):
            distance = np.linalg.norm(starts[e] - goals[e])
            if distance <= np.finfo(np.float32).eps:
                continue
            t_f = 1.0
            if max_distance is not None and distance > max_distance:
                t_f = max_distance / distance
            T = np.append(np.arange(0.0, t_f, resolution / distance), t_f)
            n_samples[e] = len(T)
            pts = (1.0 - T)[:, None] * starts[e][None] + T[:, None] * goals[e][None]
            d = self.min_distance(pts)
            min_d[e] = d.min()
            valid[e] = bool(np.all(d >= threshold))
        if return_detail:
            return valid, n_samples, min_d
        return valid
