"""Flatten a GraphChain (+ world obstacles + pair sets) into the arrays of include/nbx.h.

This is the B200-side form of the "GraphChain kinematics caching" that
reference numbotics/robots/arm.py:20-62 performs in `Arm.__init__`: per frame, the
run of FIXED joints in front of a movable joint is folded into that joint's offset
(arm.py:36-52) and the trailing FIXED run becomes an extra last offset (arm.py:54-58).
In a tree every movable joint has exactly one path from the root, so the folded
offset is a property of the joint and the trailing offset a property of the frame;
that is how the descriptor stores them (once, instead of once per frame).
"""
import numpy as np

from ..physics import Constraint, Shape
from .. import _cabi

_BULLET_DEFAULT_MARGIN = 0.04   # CONVEX_DISTANCE_MARGIN
_BULLET_SAFE_MARGIN_MULT = 0.1  # btConvexInternalShape::setSafeMargin


def pose12(T):
    return np.asarray(T, dtype=np.float64)[:3, :4].reshape(12).copy()


def primitive_params(cshape, margin_model: str):
    """CollisionShape -> (nbx type, dims[3], margin).

    margin_model:
      'bullet' -- rounded cores the way Bullet builds btBoxShape / btCylinderShapeZ
                  (margin = min(0.04, 0.1 * smallest half extent), core = shape - margin);
                  third-party behaviour restated from the Bullet sources as remembered,
                  NOT verifiable offline: see DESIGN.md "parity unpinned".
      'sharp'  -- exact sharp primitives (margin 0).
    """
    info = cshape._shape_info
    if cshape.shape in (Shape.CUBE, Shape.CUBOID):
        he = np.asarray(info["half_extents"], dtype=np.float64).reshape(3)
        if np.any(he <= 0):
            raise ValueError("box half extents must be positive")
        m = min(_BULLET_DEFAULT_MARGIN, _BULLET_SAFE_MARGIN_MULT * float(he.min())) if margin_model == "bullet" else 0.0
        return _cabi.NBX_PRIM_BOX, he, m
    if cshape.shape == Shape.SPHERE:
        r = float(info["radius"])
        if r <= 0:
            raise ValueError("sphere radius must be positive")
        return _cabi.NBX_PRIM_SPHERE, np.array([r, 0.0, 0.0]), r
    if cshape.shape == Shape.CYLINDER:
        r, hh = float(info["radius"]), 0.5 * float(info["height"])
        if r <= 0 or hh <= 0:
            raise ValueError("cylinder radius and height must be positive")
        m = min(_BULLET_DEFAULT_MARGIN, _BULLET_SAFE_MARGIN_MULT * min(r, hh)) if margin_model == "bullet" else 0.0
        return _cabi.NBX_PRIM_CYLINDER, np.array([r, hh, 0.0]), m
    raise NotImplementedError(
        f"collision shape {cshape.shape.name} is outside the batched collision path "
        "(cylinders, boxes and spheres only)")


def flatten_chain(chain, margin_model: str = "bullet") -> dict:
    """GraphChain -> dict of arrays matching nbx_chain_desc (plus name maps)."""
    if margin_model not in ("bullet", "sharp"):
        raise ValueError("margin_model must be 'bullet' or 'sharp'")
    G = chain._G
    order = chain._node_order
    root = order[0]
    frame_id = {node: i for i, node in enumerate(order)}
    parent_of = {node: next(iter(G.predecessors(node))) for node in order[1:]}

    # movable joints in the reference's q order (physics/chain.py:516-534)
    mov_of_joint = {}
    for joint in chain._joints:
        if joint.type == Constraint.FIXED:
            continue
        if joint.type == Constraint.SPHERICAL:
            raise NotImplementedError(
                "SPHERICAL joints are not supported by the batched path (the reference's batched "
                "helpers cannot evaluate them either: robots/helpers.py:101-102, 144-145)")
        mov_of_joint[joint] = len(mov_of_joint)
    n_mov = len(mov_of_joint)

    mov_parent = np.full(n_mov, -1, np.int32)
    mov_type = np.zeros(n_mov, np.int32)
    mov_qidx = np.zeros(n_mov, np.int32)
    mov_offset = np.zeros((n_mov, 12))
    mov_axis = np.zeros((n_mov, 3))
    frame_mov = np.full(len(order), -1, np.int32)
    frame_tail = np.zeros((len(order), 12))
    frame_com = np.zeros((len(order), 12))

    # walk in topological order carrying (last movable joint, accumulated fixed run)
    state = {root: (-1, np.eye(4))}
    for node in order:
        if node != root:
            par = parent_of[node]
            last, run = state[par]
            joint = G.edges[(par, node)]["joint"]
            if joint.type == Constraint.FIXED:
                state[node] = (last, run @ joint.offset)
            else:
                m = mov_of_joint[joint]
                mov_parent[m] = last
                mov_type[m] = joint.type.value
                mov_qidx[m] = chain._joint_to_index[joint]
                mov_offset[m] = pose12(run @ joint.offset)
                mov_axis[m] = joint.axis
                state[node] = (m, np.eye(4))
        last, run = state[node]
        f = frame_id[node]
        frame_mov[f] = last
        frame_tail[f] = pose12(run)
        frame_com[f] = pose12(G.nodes[node]["link"]._offset)

    # the device walk needs parents before children; topological node order guarantees the same
    # for joints only if their indices follow that order, which they do (chain._joints is in it).
    prim_frame, prim_type, prim_dims, prim_offset, prim_margin, prim_link = [], [], [], [], [], []
    for node in order:
        link = G.nodes[node]["link"]
        cs = link._collision_shape
        if cs.shape == Shape.EMPTY:
            continue
        t, dims, m = primitive_params(cs, margin_model)
        prim_frame.append(frame_id[node])
        prim_type.append(t)
        prim_dims.append(dims)
        prim_offset.append(pose12(cs.offset))
        prim_margin.append(m)
        prim_link.append(link)

    return {
        "dof": chain.dof,
        "mov_parent": mov_parent, "mov_type": mov_type, "mov_qidx": mov_qidx,
        "mov_offset": mov_offset.reshape(-1), "mov_axis": mov_axis.reshape(-1),
        "frame_mov": frame_mov, "frame_tail": frame_tail.reshape(-1), "frame_com": frame_com.reshape(-1),
        "prim_frame": np.asarray(prim_frame, np.int32), "prim_type": np.asarray(prim_type, np.int32),
        "prim_dims": np.asarray(prim_dims, np.float64).reshape(-1),
        "prim_offset": np.asarray(prim_offset, np.float64).reshape(-1),
        "prim_margin": np.asarray(prim_margin, np.float64),
        # host-side maps (not part of the C descriptor)
        "frame_id": frame_id,
        "prim_of_link": {link: i for i, link in enumerate(prim_link)},
        "margin_model": margin_model,
    }


def primitive_reach(chain_arrays: dict, joint_limits, from_first_joint: bool = False):
    """Upper bound, per link primitive, on the distance of ANY of its points from the base-frame origin,
    over every configuration: sum of the translation norms along the path (rotations preserve norms;
    anything behind a prismatic joint is unbounded) + the primitive's own offset + its bounding radius.
    `joint_limits` is accepted for signature stability and not used: verdicts must not depend on limits that
    q is never checked against.

    `from_first_joint=True` returns `(reach, root)` instead: the same bound measured from the origin of the
    primitive's FIRST movable joint (`root[p]`, -1 = the primitive hangs off the base) -- that origin is
    `base_pose . t_root`, the same point for every configuration, so the first link offset does not count
    towards the reach (for a Kinova on a pedestal that is 0.16 m of a 1.3 m bound)."""
    A = chain_arrays
    n_mov = len(A["mov_type"])
    off = np.asarray(A["mov_offset"]).reshape(n_mov, 12)
    reach_j = np.zeros(n_mov)
    root_j = np.zeros(n_mov, dtype=np.int64)
    for j in range(n_mov):
        par = int(A["mov_parent"][j])
        if from_first_joint:
            root_j[j] = j if par < 0 else root_j[par]
            r = 0.0 if par < 0 else reach_j[par] + np.linalg.norm(off[j, [3, 7, 11]])
            if par >= 0 and not np.isfinite(reach_j[par]):
                r = np.inf
        else:
            r = (reach_j[par] if par >= 0 else 0.0) + np.linalg.norm(off[j, [3, 7, 11]])
        if int(A["mov_type"][j]) == _cabi.NBX_JOINT_PRISMATIC:
            # q is never clamped to the joint limits (the reference does not clamp either), so a slider's travel is
            # unbounded as far as pruning is concerned: nothing downstream of a prismatic joint is ever pruned
            r = np.inf
        reach_j[j] = r
    tail = np.asarray(A["frame_tail"]).reshape(-1, 12)
    poff = np.asarray(A["prim_offset"]).reshape(-1, 12)
    dims = np.asarray(A["prim_dims"]).reshape(-1, 3)
    out = np.zeros(len(A["prim_type"]))
    for p in range(len(out)):
        f = int(A["prim_frame"][p])
        j = int(A["frame_mov"][f])
        t = int(A["prim_type"][p])
        brad = dims[p, 0] if t == _cabi.NBX_PRIM_SPHERE else (np.linalg.norm(dims[p]) if t == _cabi.NBX_PRIM_BOX else np.hypot(dims[p, 0], dims[p, 1]))
        out[p] = (reach_j[j] if j >= 0 else 0.0) + np.linalg.norm(tail[f, [3, 7, 11]]) + np.linalg.norm(poff[p, [3, 7, 11]]) + brad
    if from_first_joint:
        root = np.array([root_j[int(A["frame_mov"][int(A["prim_frame"][p])])] if int(A["frame_mov"][int(A["prim_frame"][p])]) >= 0 else -1
                         for p in range(len(out))], dtype=np.int64)
        return out, root
    return out


def lever_arms(chain_arrays: dict) -> np.ndarray:
    """`rho[p, c]`: upper bound, over every configuration, on the speed of ANY point of link primitive `p` per unit rate
    of q column `c` (0 when the joint of column `c` is not on the path from the root to the primitive).

    Revolute joint: a point turns about an axis through the joint's origin, so its speed is at most its distance from
    that origin <= sum of the translation norms of the link offsets between the joint and the primitive's frame + the
    frame's trailing fixed run + the primitive's own offset + its bounding radius (rotations preserve norms).
    Prismatic joint: |axis|.  A prismatic joint BETWEEN a revolute joint and the primitive makes that lever arm
    unbounded (inf).  This is what turns clearances at two samples of a motion into a bound for the motion between them
    (`Arm.edges_clear`, the batched form of reference connectors.py:108-185)."""
    A = chain_arrays
    n_mov = len(A["mov_type"])
    off = np.asarray(A["mov_offset"]).reshape(n_mov, 12)
    axis = np.asarray(A["mov_axis"]).reshape(n_mov, 3)
    tail = np.asarray(A["frame_tail"]).reshape(-1, 12)
    poff = np.asarray(A["prim_offset"]).reshape(-1, 12)
    dims = np.asarray(A["prim_dims"]).reshape(-1, 3)
    n_prims = len(A["prim_type"])
    rho = np.zeros((n_prims, int(A["dof"])))
    for p in range(n_prims):
        f = int(A["prim_frame"][p])
        t = int(A["prim_type"][p])
        brad = dims[p, 0] if t == _cabi.NBX_PRIM_SPHERE else (np.linalg.norm(dims[p]) if t == _cabi.NBX_PRIM_BOX else np.hypot(dims[p, 0], dims[p, 1]))
        arm = np.linalg.norm(tail[f, [3, 7, 11]]) + np.linalg.norm(poff[p, [3, 7, 11]]) + brad      # from the last movable joint's origin
        j = int(A["frame_mov"][f])
        while j >= 0:
            c = int(A["mov_qidx"][j])
            if int(A["mov_type"][j]) == _cabi.NBX_JOINT_PRISMATIC:
                rho[p, c] = np.linalg.norm(axis[j])
                arm = np.inf                                  # the slider's travel is not bounded (q is never clamped)
            else:
                rho[p, c] = arm
            arm = arm + np.linalg.norm(off[j, [3, 7, 11]])
            j = int(A["mov_parent"][j])
    return rho


def clearance_rate_rows(chain_arrays: dict, scene_arrays: dict) -> np.ndarray:
    """Rows `M[r, c]` such that `|d(q') - d(q)| <= max_r sum_c M[r, c] |q'_c - q_c|` for the minimum distance `d` over the
    scene's allowed pairs.  An obstacle pair changes at most as fast as the primitive's points move (`lever_arms`); a
    self pair only through the joints that move ONE of its two links -- a joint above both moves them rigidly together."""
    rho = lever_arms(chain_arrays)
    rows = []
    for a in sorted({int(a) for a, _ in np.asarray(scene_arrays["env_pairs"]).reshape(-1, 2)}):
        rows.append(rho[a])
    for a, b in np.asarray(scene_arrays["self_pairs"]).reshape(-1, 2):
        ra, rb = rho[int(a)], rho[int(b)]
        both = (ra > 0) & (rb > 0)
        rows.append(np.where(both, 0.0, ra + rb))
    if not rows:
        return np.zeros((0, rho.shape[1]))
    return np.unique(np.asarray(rows), axis=0)


def obstacle_gap_from(origin_pose, obj) -> float:
    """Lower bound on the distance from the origin of `origin_pose` (4x4, world) to the obstacle."""
    cs = obj._collision_shape
    P = obj.pose @ cs.offset
    o = origin_pose[:3, 3]
    info = cs._shape_info
    if cs.shape in (Shape.CUBE, Shape.CUBOID):
        loc = P[:3, :3].T @ (o - P[:3, 3])
        he = np.asarray(info["half_extents"], dtype=np.float64).reshape(3)
        return float(np.linalg.norm(np.maximum(np.abs(loc) - he, 0.0)))
    c = float(np.linalg.norm(o - P[:3, 3]))
    if cs.shape == Shape.SPHERE:
        return c - float(info["radius"])
    return c - float(np.hypot(info["radius"], 0.5 * info["height"]))


def flatten_scene(chain_arrays: dict, self_pairs, env_pairs, margin_model: str = "bullet",
                  prune=None) -> dict:
    """Pair sets of Link / PhysicsObject tuples -> arrays matching nbx_scene_desc.

    self_pairs: iterable of (Link, Link); env_pairs: iterable of (Link, PhysicsObject).
    Pairs whose link has no primitive (EMPTY shape) can never produce a proximity in
    the reference either and are dropped.

    prune = (base_pose, joint_limits, threshold) enables STATIC pruning of obstacle pairs: a pair is
    dropped when, for every configuration, the primitive stays farther than `threshold` from the
    obstacle (reach bound, `primitive_reach`).  Verdicts `distance < t` for any t <= threshold are
    unchanged; the pruned scene must not be used for minimum-distance queries.
    """
    prim_of_link = chain_arrays["prim_of_link"]
    prim_type = chain_arrays["prim_type"]
    env_pairs = [(link, obj) for link, obj in env_pairs if link in prim_of_link]
    n_before = len(env_pairs)
    if prune is not None:
        base_pose, joint_limits, thr = prune
        limits = np.asarray(joint_limits, dtype=np.float64).reshape(-1, 2)
        base = np.asarray(base_pose, dtype=np.float64)
        reach = primitive_reach(chain_arrays, limits)
        # the same bound from the (fixed) origin of the primitive's first movable joint: tighter by the first link offset
        reach1, root1 = primitive_reach(chain_arrays, limits, from_first_joint=True)
        n_mov = len(chain_arrays["mov_type"])
        off = np.asarray(chain_arrays["mov_offset"]).reshape(n_mov, 12)
        origin_of = {}
        gaps = {}
        kept = []
        for link, obj in env_pairs:
            p = prim_of_link[link]
            if (obj, -1) not in gaps:
                gaps[(obj, -1)] = obstacle_gap_from(base, obj)
            far = gaps[(obj, -1)] - reach[p] > thr + 1e-9
            r = int(root1[p])
            if not far and r >= 0 and np.isfinite(reach1[p]):
                if (obj, r) not in gaps:
                    if r not in origin_of:
                        X = np.eye(4)
                        X[:3, 3] = off[r, [3, 7, 11]]
                        origin_of[r] = base @ X
                    gaps[(obj, r)] = obstacle_gap_from(origin_of[r], obj)
                far = gaps[(obj, r)] - reach1[p] > thr + 1e-9
            if not far:
                kept.append((link, obj))
        env_pairs = kept
    obstacles = sorted({obj for _, obj in env_pairs}, key=lambda o: o.name)   # independent of set order
    obs_index = {obj: i for i, obj in enumerate(obstacles)}
    env_idx = [(prim_of_link[link], obs_index[obj]) for link, obj in env_pairs]
    # deterministic order independent of set iteration order
    env_idx = sorted(set(env_idx))
    self_idx = sorted({tuple(sorted((prim_of_link[a], prim_of_link[b])))
                       for a, b in self_pairs if a in prim_of_link and b in prim_of_link})

    obs_type, obs_dims, obs_pose, obs_margin = [], [], [], []
    for obj in obstacles:
        t, dims, m = primitive_params(obj._collision_shape, margin_model)
        obs_type.append(t)
        obs_dims.append(dims)
        obs_pose.append(pose12(obj.pose @ obj._collision_shape.offset))
        obs_margin.append(m)
    obs_type = np.asarray(obs_type, np.int32)
    box = _cabi.NBX_PRIM_BOX
    # Bullet dispatches box-box to btBoxBoxCollisionAlgorithm, which works on the sharp boxes
    self_sharp = np.array([prim_type[a] == box and prim_type[b] == box for a, b in self_idx], np.uint8)
    env_sharp = np.array([prim_type[a] == box and obs_type[o] == box for a, o in env_idx], np.uint8)
    return {
        "obs_type": obs_type,
        "obs_dims": np.asarray(obs_dims, np.float64).reshape(-1),
        "obs_pose": np.asarray(obs_pose, np.float64).reshape(-1),
        "obs_margin": np.asarray(obs_margin, np.float64),
        "self_pairs": np.asarray(self_idx, np.int32).reshape(-1, 2),
        "env_pairs": np.asarray(env_idx, np.int32).reshape(-1, 2),
        "self_sharp": self_sharp, "env_sharp": env_sharp,
        "obstacles": obstacles,
        "n_env_pairs_before_pruning": n_before,
    }
