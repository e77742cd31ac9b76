"""Float64 NumPy restatement of Arm.inverse_kinematics.  TEST INFRASTRUCTURE ONLY.

Follows reference numbotics/robots/arm.py:464-552 statement by statement (batched masks included), on
top of the pinned FK / Jacobian oracle (oracle/fkjac.py) and `rot_diff` of numbotics/math/spatial.py:207-212.
"""
import numpy as np


def rot_diff(A, B):
    R = B @ np.swapaxes(A, -2, -1)
    D = 0.5 * (R - np.swapaxes(R, -2, -1))
    return np.stack([D[..., 2, 1], D[..., 0, 2], D[..., 1, 0]], axis=-1)


def inverse_kinematics(arm_oracle, joint_limits, pose, q0, frame, use_com=False, use_limits=False, tol=1e-6,
                       max_iter=100, max_failures=15):
    """`arm_oracle`: oracle.fkjac.ArmOracle.  pose (B,4,4), q0 (B,dof).  Returns (converged (B,), q (B,dof), iters (B,))."""
    dof = arm_oracle.dof
    lower, upper = joint_limits[:, 0], joint_limits[:, 1]
    pose = np.asarray(pose, np.float64).reshape(-1, 4, 4)
    q = np.array(q0, dtype=np.float64).reshape(-1, dof)
    if pose.shape[0] == 1 and q.shape[0] > 1:
        pose = np.tile(pose, (q.shape[0], 1, 1))
    ee = arm_oracle.forward_kinematics(q, frame, use_com)
    diff = np.zeros((q.shape[0], 6))
    diff[:, :3] = pose[:, :3, 3] - ee[:, :3, 3]
    diff[:, 3:] = rot_diff(ee[:, :3, :3], pose[:, :3, :3])
    diff_norm = np.linalg.norm(diff, axis=-1)
    B_I = np.tile(np.eye(6)[None], (q.shape[0], 1, 1))
    lambdas = np.ones(q.shape[0]) * 1e-1
    failures = np.zeros(q.shape[0], dtype=np.int64)
    iters = np.zeros(q.shape[0], dtype=np.int64)
    for _ in range(max_iter):
        running = np.where((diff_norm > tol) & (failures < max_failures))
        if len(running[0]) == 0:
            break
        iters[running] += 1
        J = arm_oracle.jacobian(q[running], frame, use_com)
        J_T = np.swapaxes(J, -2, -1)
        q[running] = q[running] + (J_T @ np.linalg.solve((J @ J_T) + lambdas[running][..., None, None] * B_I[running],
                                                         diff[running][..., None])).squeeze(-1)
        if use_limits:
            q[running] = np.clip(q[running], lower, upper)
        ee = arm_oracle.forward_kinematics(q[running], frame, use_com)
        diff[running, :3] = pose[running, :3, 3] - ee[:, :3, 3]
        diff[running, 3:] = rot_diff(ee[:, :3, :3], pose[running, :3, :3])
        prev = np.copy(diff_norm[running])
        diff_norm = np.linalg.norm(diff, axis=-1)
        rel = diff_norm[running] > prev
        lambdas[running] *= np.where(rel, 1.2, 0.5)
        failures[running[0][rel]] += 1
        failures[running[0][~rel]] = 0
        if np.all(diff_norm < tol):
            break
    return diff_norm < tol, q, iters
