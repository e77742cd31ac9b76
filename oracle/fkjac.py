"""Float64 NumPy restatement of the reference's batched FK / Jacobian.  TEST INFRASTRUCTURE ONLY.

Restates (paths relative to the reference checkout):

* `joint_transform`         numbotics/robots/helpers.py:33-88
* `compute_transformation`  numbotics/robots/helpers.py:91-113
* `compute_jacobian`        numbotics/robots/helpers.py:117-187
* `ArmOracle.forward_kinematics` / `.jacobian`   numbotics/robots/arm.py:369-461

PINNED: tests/test_oracle_pinned.py checks these functions against the reference's own
numba helpers (loaded by oracle/ref_loader.py when the reference checkout exists)
and against tests/golden/fkjac_*.npz, which were produced by the real helpers with
tests/golden/make_golden.py.

Written batch-vectorised (einsum over the batch axis) rather than as the reference's
per-element loops; same arithmetic, different evaluation order (agreement ~1e-15).
"""
import numpy as np

FIXED, REVOLUTE, PRISMATIC, SPHERICAL = 4, 0, 1, 2


def _skew(a):
    return np.array([[0.0, -a[2], a[1]], [a[2], 0.0, -a[0]], [-a[1], a[0], 0.0]])


def joint_transform(offset, axis, q, joint_type):
    """(B,4,4) local transforms of one joint (helpers.py:33-88).

    PRISMATIC is the *intended* X = [R | t + q R a]: the reference's own prismatic branch
    (helpers.py:57-65) leaves the rotation block zero and mis-broadcasts q, so no reference
    parity exists for it.
    """
    q = np.asarray(q, dtype=np.float64)
    B = q.shape[0]
    X = np.zeros((B, 4, 4))
    X[:, 3, 3] = 1.0
    R = offset[:3, :3]
    if joint_type == FIXED:
        X[:] = offset
    elif joint_type == REVOLUTE:
        K = np.outer(axis, axis)
        K_eye = K - np.eye(3)
        K_skew = _skew(axis)
        rod = K[None] - np.cos(q)[:, None, None] * K_eye[None] + np.sin(q)[:, None, None] * K_skew[None]
        X[:, :3, :3] = np.einsum("ij,bjk->bik", R, rod)
        X[:, :3, 3] = offset[:3, 3]
    elif joint_type == PRISMATIC:
        X[:, :3, :3] = R
        X[:, :3, 3] = offset[:3, 3][None] + q[:, None] * (R @ axis)[None]
    elif joint_type == SPHERICAL:
        raise NotImplementedError("spherical joints are not part of the batched path")
    return X


def compute_transformation(T, offsets, axes, joint_types, joint_idxs, q):
    """T_b <- T_b . prod_i X_i over `offsets.shape[0]` entries (helpers.py:91-113)."""
    T = np.array(T, dtype=np.float64, copy=True)
    B = q.shape[0]
    for i in range(offsets.shape[0]):
        jt = int(joint_types[i])
        qi = q[:, int(joint_idxs[i])] if jt != FIXED else np.zeros(B)
        T = np.einsum("bij,bjk->bik", T, joint_transform(offsets[i], axes[i], qi, jt))
    return T


def compute_jacobian(T, T_mats, com_offset, offsets, axes, joint_types, joint_idxs, q,
                     local_pose=None, global_pose=None, use_com=False):
    """(B,6,dof) geometric Jacobian, rows [v; w] (helpers.py:117-187), including its behaviour when
    `offsets` has no trailing entry (the end frame overwrites the last joint's slot, :157-158)."""
    B, dof = q.shape
    I = axes.shape[0]
    J = np.zeros((B, 6, dof))
    T = np.array(T, dtype=np.float64, copy=True)
    T_mats = np.zeros((B, offsets.shape[0], 4, 4)) if T_mats is None else T_mats
    for i in range(I):
        jt = int(joint_types[i])
        if jt == SPHERICAL:
            raise NotImplementedError("Jacobian for spherical joints not implemented")
        qi = q[:, int(joint_idxs[i])] if jt != FIXED else np.zeros(B)
        T = np.einsum("bij,bjk->bik", T, joint_transform(offsets[i], axes[i], qi, jt))
        T_mats[:, i] = T
    T_mats[:, -1] = T @ offsets[-1]
    if use_com:
        T_mats[:, -1] = T_mats[:, -1] @ com_offset
    if local_pose is not None:
        T_mats[:, -1] = np.einsum("bij,bjk->bik", T_mats[:, -1], local_pose)
    if global_pose is not None:
        T_mats[:, -1] = global_pose
    for i in range(I):
        jt = int(joint_types[i])
        if jt == FIXED:
            continue
        idx = int(joint_idxs[i])
        w = np.einsum("bij,j->bi", T_mats[:, i, :3, :3], axes[i])
        if jt == REVOLUTE:
            J[:, :3, idx] = np.cross(w, T_mats[:, -1, :3, 3] - T_mats[:, i, :3, 3])
            J[:, 3:, idx] = w
        elif jt == PRISMATIC:
            J[:, :3, idx] = w
    return J


class ArmOracle:
    """Restatement of the Arm-level wrappers (arm.py:369-461) over a per-frame cache
    `{frame: (offsets, axes, joint_types, joint_idxs)}` shaped like the reference's
    `_link_joint_sequence` (arm.py:61).

    reference_compat=False (default): frames whose path ends in a movable joint get an identity
    trailing offset appended, i.e. the geometrically correct link frame -- what the CUDA path
    computes.  reference_compat=True reproduces arm.py:396-398 literally (the last joint's rotation
    is dropped for such frames).
    """

    def __init__(self, link_joint_sequence, dof, base_pose=None, com_offsets=None, reference_compat=False,
                 helpers=None):
        self.seq = link_joint_sequence
        self.dof = dof
        self.base_pose = np.eye(4) if base_pose is None else np.asarray(base_pose, dtype=np.float64)
        self.com = com_offsets or {}
        self.compat = reference_compat
        # `helpers` lets the pinned tests run the SAME wrappers over the reference's real numba kernels
        self._ct = compute_transformation if helpers is None else helpers.nb_compute_transformation
        self._cj = compute_jacobian if helpers is None else helpers.nb_compute_jacobian

    def _sequence(self, frame):
        seq = self.seq.get(frame)
        if seq is None:
            raise ValueError(f"Frame {frame} not found in chain")
        if len(seq) == 0:
            return (np.eye(4)[None], np.zeros((0, 3)), np.zeros(0, np.int64), np.zeros(0, np.int64))
        offsets, axes, types, idxs = seq
        if not self.compat and offsets.shape[0] == axes.shape[0]:
            offsets = np.concatenate([offsets, np.eye(4)[None]], axis=0)
        return offsets, axes, types, idxs

    def forward_kinematics(self, q, frame, use_com=False, local_pose=None):
        offsets, axes, types, idxs = self._sequence(frame)
        if q.shape[-1] != self.dof:
            raise ValueError(f"q must have {self.dof} elements")
        shape = q.shape
        q2 = np.ascontiguousarray(q.reshape(-1, self.dof), dtype=np.float64)
        if local_pose is not None:
            if local_pose.shape[-2:] != (4, 4):
                raise ValueError("local_pose must be a 4x4 matrix")
            if local_pose.ndim == 2:
                local_pose = np.tile(local_pose[None], (q2.shape[0], 1, 1))
            else:
                if local_pose.shape[:-2] != shape[:-1]:
                    raise ValueError("local_pose must have the same batch dimensions as q")
                local_pose = local_pose.reshape(-1, 4, 4)
        T = np.tile(self.base_pose, (q2.shape[0], 1, 1))
        if len(offsets) > 1:
            T = self._ct(T, np.ascontiguousarray(offsets[:-1]), np.ascontiguousarray(axes), types, idxs, q2)
        T = T @ offsets[-1]
        if use_com:
            T = T @ self.com.get(frame, np.eye(4))
        if local_pose is not None:
            T = np.einsum("bij,bjk->bik", T, local_pose)
        if len(shape) == 1:
            return T[0]
        return T.reshape(*shape[:-1], 4, 4)

    def jacobian(self, q, frame, use_com=False, local_pose=None, global_pose=None):
        offsets, axes, types, idxs = self._sequence(frame)
        if q.shape[-1] != self.dof:
            raise ValueError(f"q must have {self.dof} elements")
        shape = q.shape
        q2 = np.ascontiguousarray(q.reshape(-1, self.dof), dtype=np.float64)
        B = q2.shape[0]

        def _prep(p, name):
            if p is None:
                return None
            if p.shape[-2:] != (4, 4):
                raise ValueError(f"{name} must be a 4x4 matrix")
            if p.ndim == 2:
                return np.ascontiguousarray(np.tile(p[None], (B, 1, 1)))
            if p.shape[:-2] != shape[:-1]:
                raise ValueError(f"{name} must have the same batch dimensions as q")
            return np.ascontiguousarray(p.reshape(-1, 4, 4))

        local_pose = _prep(local_pose, "local_pose")
        if global_pose is not None and local_pose is not None:
            raise ValueError("local_pose and global_pose cannot both be provided")
        global_pose = _prep(global_pose, "global_pose")
        T_mats = np.zeros((B, offsets.shape[0], 4, 4))
        T = np.tile(self.base_pose, (B, 1, 1))
        com = np.ascontiguousarray(self.com.get(frame, np.eye(4)))
        J = np.zeros((B, 6, self.dof))
        if len(offsets) > 1:
            J = self._cj(T, T_mats, com, np.ascontiguousarray(offsets), np.ascontiguousarray(axes), types, idxs, q2,
                         local_pose, global_pose, bool(use_com))
        if len(shape) == 1:
            return J[0]
        return J.reshape(*shape[:-1], 6, self.dof)

    def link_frames(self, q):
        """{frame: (B,4,4)} for every frame -- the all-link FK the collision path needs."""
        return {f: self.forward_kinematics(q, f) for f in self.seq}
