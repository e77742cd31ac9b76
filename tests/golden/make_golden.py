"""Generate tests/golden/*.npz from the REFERENCE's own numba kernels.

Run in the build container only (needs /root/reference and numba):

    python tests/golden/make_golden.py

For every case the inputs (q, the per-frame (offsets, axes, joint_types, joint_idxs) sequence,
base pose, optional com / local / global poses) and the outputs of
`nb_compute_transformation` / `nb_compute_jacobian` (numbotics/robots/helpers.py:91-187),
driven through a restatement of the Arm wrappers (numbotics/robots/arm.py:369-461, with the
reference's literal `offsets[:-1]` handling), are stored as float64.  The golden files are the
pin for oracle/fkjac.py (tests/test_oracle_pinned.py) and for the CUDA path (tests/test_fkjac_gpu.py).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader, fkjac  # noqa: E402
from numbotics_b200 import workloads  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def seq_arrays(arm, frame):
    offsets, axes, types, idxs = arm._link_joint_sequence[frame]
    return dict(offsets=offsets, axes=axes, joint_types=types, joint_idxs=idxs)


def main():
    helpers = ref_loader.load()
    rng = np.random.default_rng(20261018)

    # ---- Kinova 7-DoF ------------------------------------------------------------------------
    arm = workloads.kinova_arm()
    com = {n: l._offset for n, l in arm._links_from_nodes.items()}
    base = arm.base_pose
    ref = fkjac.ArmOracle(arm._link_joint_sequence, arm.dof, base, com, reference_compat=True, helpers=helpers)
    lo, hi = workloads.sample_limits(arm.joint_limits)
    B = 96
    q = rng.uniform(lo, hi, size=(B, arm.dof))
    q[0] = 0.0
    q[1] = lo
    q[2] = hi
    lp = np.tile(np.eye(4), (B, 1, 1))
    ang = rng.uniform(-np.pi, np.pi, B)
    lp[:, 0, 0], lp[:, 0, 1], lp[:, 1, 0], lp[:, 1, 1] = np.cos(ang), -np.sin(ang), np.sin(ang), np.cos(ang)
    lp[:, :3, 3] = rng.normal(scale=0.2, size=(B, 3))
    gp = np.tile(np.eye(4), (B, 1, 1))
    gp[:, :3, 3] = rng.normal(scale=0.5, size=(B, 3))
    out = dict(q=q, base_pose=base, local_pose=lp, global_pose=gp)
    # frames that end in a fixed run: the reference is geometrically right there
    for frame in ["tool_frame", "end_effector_link", "camera", "gripper"]:
        out[f"{frame}/T"] = ref.forward_kinematics(q, frame)
        out[f"{frame}/J"] = ref.jacobian(q, frame)
        out[f"{frame}/T_com"] = ref.forward_kinematics(q, frame, use_com=True)
        out[f"{frame}/J_com"] = ref.jacobian(q, frame, use_com=True)
        out[f"{frame}/com"] = com[frame]
        for k, v in seq_arrays(arm, frame).items():
            out[f"{frame}/{k}"] = v
    out["tool_frame/T_lp"] = ref.forward_kinematics(q, "tool_frame", local_pose=lp)
    out["tool_frame/J_lp"] = ref.jacobian(q, "tool_frame", local_pose=lp)
    out["tool_frame/J_gp"] = ref.jacobian(q, "tool_frame", global_pose=gp)
    out["tool_frame/T_lp1"] = ref.forward_kinematics(q, "tool_frame", local_pose=lp[0])
    # frames that end in a movable joint: golden records the reference's literal (quirky) output
    for frame in ["bracelet_link", "forearm_link", "shoulder_link"]:
        out[f"{frame}/T_compat"] = ref.forward_kinematics(q, frame)
        out[f"{frame}/J_compat"] = ref.jacobian(q, frame)
        for k, v in seq_arrays(arm, frame).items():
            out[f"{frame}/{k}"] = v
    np.savez_compressed(os.path.join(OUT, "fkjac_kinova.npz"), **out)

    # ---- raw helper calls with a non-constant initial transform and interior FIXED entries -----
    offsets, axes, types, idxs = arm._link_joint_sequence["tool_frame"]
    T0 = np.tile(np.eye(4), (B, 1, 1))
    T0[:, :3, 3] = rng.normal(size=(B, 3))
    c, s = np.cos(ang), np.sin(ang)
    T0[:, 1, 1], T0[:, 1, 2], T0[:, 2, 1], T0[:, 2, 2] = c, -s, s, c
    mixed_off = np.concatenate([offsets[:3], offsets[7:8], offsets[3:]], axis=0)          # a FIXED entry in the middle
    mixed_axes = np.concatenate([axes[:3], np.zeros((1, 3)), axes[3:]], axis=0)
    mixed_types = np.concatenate([types[:3], [4], types[3:]]).astype(np.int64)
    mixed_idxs = np.concatenate([idxs[:3], [0], idxs[3:]]).astype(np.int64)
    raw = dict(q=q, T0=T0, offsets=offsets, axes=axes, joint_types=types, joint_idxs=idxs,
               mixed_offsets=mixed_off, mixed_axes=mixed_axes, mixed_types=mixed_types, mixed_idxs=mixed_idxs)
    raw["T"] = helpers.nb_compute_transformation(T0.copy(), np.ascontiguousarray(offsets[:-1]), axes, types, idxs, q)
    raw["J"] = helpers.nb_compute_jacobian(T0.copy(), np.zeros((B, 8, 4, 4)), np.eye(4), offsets, axes, types, idxs, q,
                                           None, None, False)
    raw["T_mixed"] = helpers.nb_compute_transformation(T0.copy(), np.ascontiguousarray(mixed_off[:-1]), mixed_axes,
                                                       mixed_types, mixed_idxs, q)
    raw["J_mixed"] = helpers.nb_compute_jacobian(T0.copy(), np.zeros((B, 9, 4, 4)), np.eye(4), mixed_off, mixed_axes[:8],
                                                 mixed_types[:8], mixed_idxs[:8], q, None, None, False)
    np.savez_compressed(os.path.join(OUT, "helpers_raw.npz"), **raw)

    # ---- 30-DoF serial chain (BASELINE config 5) -------------------------------------------------
    arm30 = workloads.serial_arm(30)
    ref30 = fkjac.ArmOracle(arm30._link_joint_sequence, 30, arm30.base_pose, {}, reference_compat=True, helpers=helpers)
    q30 = rng.uniform(-np.pi, np.pi, size=(48, 30))
    out30 = dict(q=q30, base_pose=arm30.base_pose)
    out30["tool/T"] = ref30.forward_kinematics(q30, "tool")
    out30["tool/J"] = ref30.jacobian(q30, "tool")
    for k, v in seq_arrays(arm30, "tool").items():
        out30[f"tool/{k}"] = v
    np.savez_compressed(os.path.join(OUT, "fkjac_serial30.npz"), **out30)
    for f in ("fkjac_kinova.npz", "helpers_raw.npz", "fkjac_serial30.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()
