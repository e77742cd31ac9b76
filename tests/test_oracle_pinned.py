"""Pin the float64 oracle: against the golden vectors produced by the reference's numba kernels,
and (when the reference checkout is present, i.e. in the build container) against those kernels live."""
import numpy as np
import pytest

from oracle import fkjac, ref_loader, collision
from numbotics_b200 import workloads
from numbotics_b200.robots import descriptor

TOL = 1e-12


def _seq(g, frame):
    return (g[f"{frame}/offsets"], g[f"{frame}/axes"], g[f"{frame}/joint_types"], g[f"{frame}/joint_idxs"])


def test_oracle_matches_golden_kinova(golden_kinova):
    g = golden_kinova
    frames = ["tool_frame", "end_effector_link", "camera", "gripper", "bracelet_link", "forearm_link", "shoulder_link"]
    seq = {f: _seq(g, f) for f in frames}
    com = {f: g[f"{f}/com"] for f in frames if f"{f}/com" in g}
    ora = fkjac.ArmOracle(seq, 7, g["base_pose"], com, reference_compat=True)
    q = g["q"]
    for f in ["tool_frame", "end_effector_link", "camera", "gripper"]:
        assert np.abs(ora.forward_kinematics(q, f) - g[f"{f}/T"]).max() < TOL
        assert np.abs(ora.jacobian(q, f) - g[f"{f}/J"]).max() < TOL
        assert np.abs(ora.forward_kinematics(q, f, use_com=True) - g[f"{f}/T_com"]).max() < TOL
        assert np.abs(ora.jacobian(q, f, use_com=True) - g[f"{f}/J_com"]).max() < TOL
    assert np.abs(ora.forward_kinematics(q, "tool_frame", local_pose=g["local_pose"]) - g["tool_frame/T_lp"]).max() < TOL
    assert np.abs(ora.jacobian(q, "tool_frame", local_pose=g["local_pose"]) - g["tool_frame/J_lp"]).max() < TOL
    assert np.abs(ora.jacobian(q, "tool_frame", global_pose=g["global_pose"]) - g["tool_frame/J_gp"]).max() < TOL
    assert np.abs(ora.forward_kinematics(q, "tool_frame", local_pose=g["local_pose"][0]) - g["tool_frame/T_lp1"]).max() < TOL
    # the reference's literal behaviour on frames that end in a movable joint
    for f in ["bracelet_link", "forearm_link", "shoulder_link"]:
        assert np.abs(ora.forward_kinematics(q, f) - g[f"{f}/T_compat"]).max() < TOL
        assert np.abs(ora.jacobian(q, f) - g[f"{f}/J_compat"]).max() < TOL


def test_oracle_matches_golden_raw_helpers(golden_raw):
    g = golden_raw
    q, T0 = g["q"], g["T0"]
    T = fkjac.compute_transformation(T0, g["offsets"][:-1], g["axes"], g["joint_types"], g["joint_idxs"], q)
    assert np.abs(T - g["T"]).max() < TOL
    J = fkjac.compute_jacobian(T0, None, np.eye(4), g["offsets"], g["axes"], g["joint_types"], g["joint_idxs"], q)
    assert np.abs(J - g["J"]).max() < TOL
    Tm = fkjac.compute_transformation(T0, g["mixed_offsets"][:-1], g["mixed_axes"], g["mixed_types"], g["mixed_idxs"], q)
    assert np.abs(Tm - g["T_mixed"]).max() < TOL
    Jm = fkjac.compute_jacobian(T0, None, np.eye(4), g["mixed_offsets"], g["mixed_axes"][:8], g["mixed_types"][:8],
                                g["mixed_idxs"][:8], q)
    assert np.abs(Jm - g["J_mixed"]).max() < TOL


def test_oracle_matches_golden_serial30(golden_serial30):
    g = golden_serial30
    ora = fkjac.ArmOracle({"tool": _seq(g, "tool")}, 30, g["base_pose"], {}, reference_compat=True)
    assert np.abs(ora.forward_kinematics(g["q"], "tool") - g["tool/T"]).max() < TOL
    assert np.abs(ora.jacobian(g["q"], "tool") - g["tool/J"]).max() < TOL


def test_model_rebuilds_the_golden_sequences(golden_kinova, golden_serial30):
    """The authored URDF + the Arm cache reproduce exactly the sequences the golden files were made with."""
    arm = workloads.kinova_arm()
    for f in ["tool_frame", "camera", "bracelet_link"]:
        for got, want in zip(arm._link_joint_sequence[f], _seq(golden_kinova, f)):
            assert np.array_equal(got, want)
    arm30 = workloads.serial_arm(30)
    for got, want in zip(arm30._link_joint_sequence["tool"], _seq(golden_serial30, "tool")):
        assert np.array_equal(got, want)


def test_c_pipeline_fkjac_matches_numpy_oracle():
    arm = workloads.kinova_arm()
    desc = collision.CDesc(arm._arrays)
    q = workloads.random_configs_np(arm, 257, seed=11)
    com = {n: l._offset for n, l in arm._links_from_nodes.items()}
    ora = fkjac.ArmOracle(arm._link_joint_sequence, 7, arm.base_pose, com)
    for frame in ["tool_frame", "forearm_link", "camera", "base_link"]:
        T, J = collision.c_fk_jac(desc, arm._arrays["frame_id"][frame], q, arm.base_pose)
        assert np.abs(T - ora.forward_kinematics(q, frame)).max() < TOL
        assert np.abs(J - ora.jacobian(q, frame)).max() < TOL
    T, J = collision.c_fk_jac(desc, arm._arrays["frame_id"]["gripper"], q, arm.base_pose, use_com=True)
    assert np.abs(T - ora.forward_kinematics(q, "gripper", use_com=True)).max() < TOL
    assert np.abs(J - ora.jacobian(q, "gripper", use_com=True)).max() < TOL


def test_jacobian_is_the_derivative_of_fk():
    """Independent of any stored vector: central differences of the oracle FK."""
    arm = workloads.kinova_arm()
    ora = fkjac.ArmOracle(arm._link_joint_sequence, 7, arm.base_pose)
    q = workloads.random_configs_np(arm, 16, seed=4)
    J = ora.jacobian(q, "tool_frame")
    h = 1e-6
    for k in range(7):
        dq = np.zeros(7); dq[k] = h
        Tp, Tm = ora.forward_kinematics(q + dq, "tool_frame"), ora.forward_kinematics(q - dq, "tool_frame")
        v = (Tp[:, :3, 3] - Tm[:, :3, 3]) / (2 * h)
        assert np.abs(v - J[:, :3, k]).max() < 1e-8
        dR = (Tp[:, :3, :3] - Tm[:, :3, :3]) / (2 * h)
        W = dR @ np.swapaxes(ora.forward_kinematics(q, "tool_frame")[:, :3, :3], 1, 2)
        w = np.stack([W[:, 2, 1], W[:, 0, 2], W[:, 1, 0]], axis=1)
        assert np.abs(w - J[:, 3:, k]).max() < 1e-8


@pytest.mark.skipif(not ref_loader.available(), reason="reference checkout / numba not present (GPU box)")
def test_oracle_matches_live_reference_kernels():
    helpers = ref_loader.load()
    arm = workloads.kinova_arm()
    com = {n: l._offset for n, l in arm._links_from_nodes.items()}
    q = workloads.random_configs_np(arm, 4096, seed=0)      # BASELINE config 1
    ref = fkjac.ArmOracle(arm._link_joint_sequence, 7, arm.base_pose, com, reference_compat=True, helpers=helpers)
    ora = fkjac.ArmOracle(arm._link_joint_sequence, 7, arm.base_pose, com, reference_compat=True)
    for frame in arm._link_joint_sequence:
        if len(arm._link_joint_sequence[frame]) == 0:
            continue   # the reference cannot unpack the root's empty sequence (arm.py:395)
        assert np.abs(ref.forward_kinematics(q, frame) - ora.forward_kinematics(q, frame)).max() < TOL, frame
        assert np.abs(ref.jacobian(q, frame) - ora.jacobian(q, frame)).max() < TOL, frame


def test_staged_reference_helpers_are_the_reference_file():
    """oracle/_ref/ref_helpers.py (what bench.py's reference-numba leg times on the GPU box) is the reference's
    numbotics/robots/helpers.py byte for byte -- staged by ref_loader.stage(), never edited."""
    import os
    if not os.path.exists(ref_loader.LIVE_PATH):
        pytest.skip("reference checkout not visible (GPU box)")
    staged = ref_loader.stage()
    assert staged == ref_loader.STAGED_PATH
    assert open(staged, "rb").read() == open(ref_loader.LIVE_PATH, "rb").read()
