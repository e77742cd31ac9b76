"""Batched damped-least-squares IK (first 'next' row) vs the float64 restatement of
reference arm.py:464-552, through the C-ABI."""
import numpy as np
import pytest

from numbotics_b200 import workloads
from oracle import fkjac, ik as ik_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup(cuda):
    arm = workloads.kinova_arm(device=0)
    com = {n: l._offset for n, l in arm._links_from_nodes.items()}
    return arm, fkjac.ArmOracle(arm._link_joint_sequence, 7, arm.base_pose, com)


def test_ik_matches_oracle_f64(setup):
    arm, ora = setup
    rng = np.random.default_rng(0)
    q_true = workloads.random_configs_np(arm, 400, seed=1) * 0.6
    pose = ora.forward_kinematics(q_true, "tool_frame")
    q0 = q_true + rng.normal(scale=0.25, size=q_true.shape)
    ok, q, it = arm.inverse_kinematics(pose, q0, "tool_frame", return_iters=True)
    ok_o, q_o, it_o = ik_oracle.inverse_kinematics(ora, arm.joint_limits, pose, q0, "tool_frame")
    assert ok.shape == (400,) and q.shape == (400, 7) and ok.dtype == bool
    assert ok.mean() > 0.9 and np.array_equal(ok, ok_o)
    # the same iteration, joint for joint: identical iteration counts and the same solution
    assert np.array_equal(it, it_o)
    assert np.abs(q - q_o).max() < 1e-7
    # and the solutions really solve the problem
    Tq = ora.forward_kinematics(q[ok], "tool_frame")
    assert np.abs(Tq - pose[ok]).max() < 5e-6


def test_ik_limits_failures_and_shapes(setup):
    arm, ora = setup
    rng = np.random.default_rng(3)
    q_true = workloads.random_configs_np(arm, 64, seed=5) * 0.5
    pose = ora.forward_kinematics(q_true, "end_effector_link")
    q0 = q_true + rng.normal(scale=0.3, size=q_true.shape)
    ok, q = arm.inverse_kinematics(pose, q0, "end_effector_link", use_limits=True, max_iter=40)
    ok_o, q_o, _ = ik_oracle.inverse_kinematics(ora, arm.joint_limits, pose, q0, "end_effector_link", use_limits=True, max_iter=40)
    assert np.array_equal(ok, ok_o) and np.abs(q - q_o).max() < 1e-6
    lim = arm.joint_limits
    assert np.all(q >= lim[:, 0] - 1e-12) and np.all(q <= lim[:, 1] + 1e-12)
    # unreachable targets: never converge, stop by max_failures / max_iter exactly like the oracle
    far = pose.copy(); far[:, :3, 3] += np.array([3.0, 0, 0])
    ok2, q2, it2 = arm.inverse_kinematics(far[:16], q0[:16], "end_effector_link", max_iter=30, return_iters=True)
    ok2_o, q2_o, it2_o = ik_oracle.inverse_kinematics(ora, arm.joint_limits, far[:16], q0[:16], "end_effector_link", max_iter=30)
    assert not ok2.any() and np.array_equal(it2, it2_o)
    # single problem: reference returns q[0] and a (1,) flag array
    ok1, q1 = arm.inverse_kinematics(pose[0], q0[0], "end_effector_link")
    assert q1.shape == (7,) and ok1.shape == (1,)
    # one target pose, many seeds
    okb, qb = arm.inverse_kinematics(pose[0], q0[:8], "end_effector_link")
    assert qb.shape == (8, 7) and okb.shape == (8,)
    with pytest.raises(ValueError, match="q0 must have 7 elements"):
        arm.inverse_kinematics(pose[0], np.zeros(6), "end_effector_link")
    with pytest.raises(ValueError, match="max_iter must be greater than 0"):
        arm.inverse_kinematics(pose[0], q0[0], "end_effector_link", max_iter=0)
    with pytest.raises(ValueError, match="same batch dimensions"):
        arm.inverse_kinematics(pose[:4], q0[:5], "end_effector_link")


def test_ik_torch_f32_and_deep_chain(cuda):
    import torch
    arm = workloads.kinova_arm(device=0)
    ora = fkjac.ArmOracle(arm._link_joint_sequence, 7, arm.base_pose)
    q_true = workloads.random_configs_np(arm, 2048, seed=9) * 0.6
    pose = ora.forward_kinematics(q_true, "tool_frame")
    q0 = q_true + np.random.default_rng(1).normal(scale=0.2, size=q_true.shape)
    ok, q = arm.inverse_kinematics(torch.as_tensor(pose, dtype=torch.float32, device=cuda),
                                   torch.as_tensor(q0, dtype=torch.float32, device=cuda), "tool_frame", tol=2e-5)
    assert q.is_cuda and q.dtype == torch.float32 and ok.dtype == torch.bool
    assert float(ok.float().mean()) > 0.9
    Tq = ora.forward_kinematics(q[ok].double().cpu().numpy(), "tool_frame")
    assert np.abs(Tq - pose[ok.cpu().numpy()]).max() < 1e-4
    # 30-DoF serial chain: registers stay O(1) in depth
    arm30 = workloads.serial_arm(30, device=0)
    ora30 = fkjac.ArmOracle(arm30._link_joint_sequence, 30, arm30.base_pose)
    qt = np.random.default_rng(2).uniform(-0.5, 0.5, size=(128, 30))
    pose30 = ora30.forward_kinematics(qt, "tool")
    q030 = qt + np.random.default_rng(3).normal(scale=0.05, size=qt.shape)
    ok30, q30 = arm30.inverse_kinematics(pose30, q030, "tool")
    ok30_o, q30_o, _ = ik_oracle.inverse_kinematics(ora30, arm30.joint_limits, pose30, q030, "tool")
    assert np.array_equal(ok30, ok30_o) and ok30.mean() > 0.9
    assert np.abs(q30 - q30_o).max() < 1e-6
