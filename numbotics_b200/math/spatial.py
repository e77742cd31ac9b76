"""Rigid-transform helpers (host side, float64 NumPy).

Behavioural mirror of reference numbotics/math/spatial.py:157-178 (`trans_mat`),
:181-190 (`skew`) and :207-212 (`rot_diff`, the second definition, which is the
one that is live in the reference because it shadows the first).
"""
import numpy as np


def trans_mat(pos=None, orn=None):
    """Homogeneous transform(s) from position(s) `(...,3)` and rotation(s) `(...,3,3)`."""
    pos = np.zeros(3) if pos is None else np.asarray(pos, dtype=np.float64)
    orn = np.eye(3) if orn is None else np.asarray(orn, dtype=np.float64)
    if pos.shape[-1] != 3:
        raise ValueError(f"Position must have 3 elements, got {pos.shape[-1]}")
    if orn.shape[-2:] != (3, 3):
        raise ValueError(f"Orientation must be a 3x3 matrix, got {orn.shape}")
    if pos.ndim != orn.ndim - 1:
        raise ValueError(
            "Position and orientation must have the same number of batch dimensions, "
            f"got {pos.ndim - 1} and {orn.ndim - 2}")
    batch = orn.shape[:-2]
    if pos.shape[:-1] != batch:
        raise ValueError(
            f"Position and orientation must have the same batch dimensions, got {pos.shape[:-1]} and {batch}")
    T = np.zeros(batch + (4, 4))
    T[..., :3, :3] = orn
    T[..., :3, 3] = pos
    T[..., 3, 3] = 1.0
    return T


def skew(v):
    v = np.asarray(v, dtype=np.float64)
    S = np.zeros(v.shape + (3,))
    S[..., 0, 1] = -v[..., 2]
    S[..., 0, 2] = v[..., 1]
    S[..., 1, 0] = v[..., 2]
    S[..., 1, 2] = -v[..., 0]
    S[..., 2, 0] = -v[..., 1]
    S[..., 2, 1] = v[..., 0]
    return S


def rot_diff(A, B):
    """Vee of the skew part of B A^T (reference spatial.py:207-212)."""
    A = np.asarray(A)
    B = np.asarray(B)
    R = B @ np.swapaxes(A, -2, -1)
    D = 0.5 * (R - np.swapaxes(R, -2, -1))
    return np.stack([D[..., 2, 1], D[..., 0, 2], D[..., 1, 0]], axis=-1)


def rpy_matrix(rpy):
    """URDF roll-pitch-yaw -> rotation matrix.

    The reference converts URDF `rpy` with scipy `Rotation.from_euler('xyz', rpy)`
    (numbotics/physics/helpers.py:313-316): lower-case = extrinsic x, y, z, i.e.
    R = Rz(yaw) @ Ry(pitch) @ Rx(roll).  Written out so the loader needs no scipy.
    """
    r, p, y = (float(a) for a in rpy)
    cr, sr = np.cos(r), np.sin(r)
    cp, sp = np.cos(p), np.sin(p)
    cy, sy = np.cos(y), np.sin(y)
    return np.array([
        [cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
        [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
        [-sp, cp * sr, cp * cr],
    ])
