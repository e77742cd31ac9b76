"""Deep / wide chains: every kernel family behind `Arm.fk_jacobian` against the float64 oracle.

The dispatcher of `csrc/nbx_fkjac.cu` has four families: dense (<= 8 joints), wide (float, 9..32 joints, path = all q
columns), deep (canonical-axis / prismatic joints, any path, float and double) and the legacy generic kernel (oblique or
non-unit axes).  `NBX_NO_WIDE` / `NBX_GENERIC_LEGACY` (read once per process) force the later families, so the same
checks run again in sub-processes and all families are held to the oracle (restating reference
numbotics/robots/helpers.py:91-187, numbotics/robots/arm.py:369-461) on identical inputs.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import fkjac
from numbotics_b200 import workloads

pytestmark = pytest.mark.gpu

RTOL64 = 1e-11


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = np.maximum(np.abs(b).reshape(b.shape[0], -1).max(axis=1), 1.0).reshape((-1,) + (1,) * (b.ndim - 1))
    return float((np.abs(a - b) / scale).max()) if a.size else 0.0


def _mixed_serial_arm(n, seed, allz=False, prismatic=True, device=0):
    """Serial chain of n movable joints with canonical axes (+-x, +-y, +-z), optional prismatic joints, random
    offsets, a fixed joint every few links and a fixed tool."""
    import networkx as nx
    from numbotics_b200.math import trans_mat
    from numbotics_b200.physics import GraphChain, Joint, Constraint, DummyLink
    from numbotics_b200.robots import Arm
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(seed)
    w = workloads.fresh_world(f"mixed{n}_{seed}")
    G = nx.DiGraph()
    G.add_node("l0", link=DummyLink(1.0, trans_mat(pos=rng.normal(size=3) * 0.02), name="l0"))
    axes = [[0, 0, 1], [0, 0, -1], [1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0]]
    prev, k = "l0", 0
    for i in range(n):
        off = trans_mat(pos=rng.normal(size=3) * 0.08, orn=Rotation.from_rotvec(rng.normal(size=3)).as_matrix())
        ax = np.array([0.0, 0.0, 1.0]) if allz else np.array(axes[int(rng.integers(0, 6))], float)
        pris = prismatic and not allz and rng.random() < 0.15
        name = f"l{i + 1}"
        G.add_node(name, link=DummyLink(1.0, trans_mat(pos=rng.normal(size=3) * 0.01), name=name))
        G.add_edge(prev, name, joint=Joint(off, ax, Constraint.PRISMATIC if pris else Constraint.REVOLUTE, name=f"j{i + 1}",
                                           lower_limit=-0.3 if pris else -np.pi, upper_limit=0.3 if pris else np.pi))
        prev = name
        if i % 5 == 4:   # a fixed joint in the middle of the chain (folded into the next offset)
            k += 1
            fname = f"f{k}"
            G.add_node(fname, link=DummyLink(0.1, np.eye(4), name=fname))
            G.add_edge(prev, fname, joint=Joint(trans_mat(pos=rng.normal(size=3) * 0.03, orn=Rotation.from_rotvec(rng.normal(size=3)).as_matrix()),
                                                np.zeros(3), Constraint.FIXED, name=f"fj{k}"))
            prev = fname
    G.add_node("tool", link=DummyLink(0.1, np.eye(4), name="tool"))
    G.add_edge(prev, "tool", joint=Joint(trans_mat(pos=np.array([0.05, 0.01, -0.02])), np.zeros(3), Constraint.FIXED, name="tool_joint"))
    return Arm(GraphChain(G, world_name=w.name), device=device)


def _check_arm(arm, n, seed, tol32):
    rng = np.random.default_rng(100 + seed)
    com = {nm: l._offset for nm, l in arm._links_from_nodes.items()}
    ora = fkjac.ArmOracle(arm._link_joint_sequence, arm.dof, arm.base_pose, com)
    assert arm.dof == n
    for B in (1000, 37):   # ragged: neither a multiple of 32 nor of 128
        q = workloads.random_configs_np(arm, B, seed=seed)
        frames = ["tool", f"l{n}", f"l{max(1, n // 2)}", "l1"]
        for frame in frames:
            To, Jo = ora.forward_kinematics(q, frame), ora.jacobian(q, frame)
            T64, J64 = arm.fk_jacobian(q, frame)
            assert rel_err(T64, To) <= RTOL64, (n, frame, "f64 T")
            assert rel_err(J64, Jo) <= RTOL64, (n, frame, "f64 J")
            T32, J32 = arm.fk_jacobian(q.astype(np.float32), frame)
            assert T32.dtype == np.float32 and J32.dtype == np.float32
            assert rel_err(T32, To) <= tol32, (n, frame, "f32 T")
            assert rel_err(J32, Jo) <= tol32, (n, frame, "f32 J")
            # the separate calls (FK only / J only launches) give the same bits as the fused call
            assert np.array_equal(arm.forward_kinematics(q.astype(np.float32), frame), T32)
            assert np.array_equal(arm.jacobian(q.astype(np.float32), frame), J32)
        # use_com, per-configuration local_pose, and the Jacobian of a given world point
        lp = np.tile(np.eye(4), (B, 1, 1))
        lp[:, :3, 3] = rng.normal(size=(B, 3)) * 0.1
        from scipy.spatial.transform import Rotation
        lp[:, :3, :3] = Rotation.from_rotvec(rng.normal(size=(B, 3))).as_matrix()
        gp = np.tile(np.eye(4), (B, 1, 1))
        gp[:, :3, 3] = rng.normal(size=(B, 3)) * 0.5
        for dt, tol in ((np.float64, RTOL64), (np.float32, tol32)):
            qd = q.astype(dt)
            assert rel_err(arm.forward_kinematics(qd, "tool", use_com=True), ora.forward_kinematics(q, "tool", use_com=True)) <= tol
            assert rel_err(arm.jacobian(qd, "tool", use_com=True), ora.jacobian(q, "tool", use_com=True)) <= tol
            assert rel_err(arm.forward_kinematics(qd, "tool", local_pose=lp.astype(dt)), ora.forward_kinematics(q, "tool", local_pose=lp)) <= tol
            assert rel_err(arm.jacobian(qd, "tool", local_pose=lp.astype(dt)), ora.jacobian(q, "tool", local_pose=lp)) <= tol
            assert rel_err(arm.jacobian(qd, "tool", global_pose=gp.astype(dt)), ora.jacobian(q, "tool", global_pose=gp)) <= tol
            mid = f"l{max(1, n // 2)}"
            assert rel_err(arm.jacobian(qd, mid, global_pose=gp.astype(dt)), ora.jacobian(q, mid, global_pose=gp)) <= tol


# 8 | 9: dense -> wide boundary; 22 | 23: the two wide translation units; 32 | 33: wide -> deep boundary
@pytest.mark.parametrize("n", [8, 9, 13, 22, 23, 30, 32, 33, 48, 64])
def test_mixed_serial_chains(cuda, n):
    # north_star: 1e-5 relative for fp32 at every depth (measured: 3.5e-7 at 7 joints, 5.5e-7 at 30, 1.1e-6 at 48 on 2^20
    # configurations, profiles/r02_fp32_error_vs_depth.txt; round 1 asserted 3e-5 beyond 16 joints without having measured)
    tol32 = 1e-5
    _check_arm(_mixed_serial_arm(n, seed=n), n, seed=n, tol32=tol32)


@pytest.mark.parametrize("n", [12, 31])
def test_allz_serial_chains(cuda, n):
    _check_arm(_mixed_serial_arm(n, seed=1000 + n, allz=True), n, seed=n, tol32=1e-5)


def _dh_serial_arm(n, seed, device=0):
    """DH-style serial chain: every joint about its own +z or -z, link offsets with a roll-only rotation (angles
    including 0, pi and +-pi/2) and a full translation -- the compile-time `ALLZ == 2` variants, whose axis signs the
    host folds into the offsets (fold_axis_signs, csrc/nbx_fkjac.cu)."""
    import networkx as nx
    from numbotics_b200.math import trans_mat
    from numbotics_b200.physics import GraphChain, Joint, Constraint, DummyLink
    from numbotics_b200.robots import Arm
    rng = np.random.default_rng(seed)
    w = workloads.fresh_world(f"dh{n}_{seed}")
    G = nx.DiGraph()
    G.add_node("l0", link=DummyLink(1.0, trans_mat(pos=rng.normal(size=3) * 0.02), name="l0"))
    special = [0.0, np.pi, np.pi / 2, -np.pi / 2]
    prev = "l0"
    for i in range(n):
        roll = special[int(rng.integers(0, 4))] if rng.random() < 0.5 else float(rng.uniform(-np.pi, np.pi))
        c, s_ = np.cos(roll), np.sin(roll)
        if roll == 0.0:
            c, s_ = 1.0, 0.0
        R = np.array([[1.0, 0.0, 0.0], [0.0, c, -s_], [0.0, s_, c]])
        off = trans_mat(pos=rng.normal(size=3) * 0.08, orn=R)
        ax = np.array([0.0, 0.0, 1.0 if rng.random() < 0.5 else -1.0])
        name = f"l{i + 1}"
        G.add_node(name, link=DummyLink(1.0, trans_mat(pos=rng.normal(size=3) * 0.01), name=name))
        G.add_edge(prev, name, joint=Joint(off, ax, Constraint.REVOLUTE, name=f"j{i + 1}", lower_limit=-np.pi, upper_limit=np.pi))
        prev = name
    G.add_node("tool", link=DummyLink(0.1, np.eye(4), name="tool"))
    G.add_edge(prev, "tool", joint=Joint(trans_mat(pos=np.array([0.05, 0.01, -0.02])), np.zeros(3), Constraint.FIXED, name="tool_joint"))
    return Arm(GraphChain(G, world_name=w.name), device=device)


# 7: dense kernel; 9, 30, 31 (odd row length), 32: wide kernels with folded axis signs; 40: deep kernel
@pytest.mark.parametrize("n", [7, 9, 30, 31, 32, 40])
def test_dh_chains_with_negative_axes(cuda, n):
    arm = _dh_serial_arm(n, seed=2000 + n)
    _check_arm(arm, n, seed=n, tol32=1e-5)


@pytest.mark.parametrize("n", [9, 30, 31])
def test_wide_kernel_huge_angles(cuda, n):
    """|q| > 1e5 leaves the fast sine / cosine range: the wide kernels screen each q tile once and reduce the offenders
    out of line.  float32 q is exact in float64, so the oracle sees the same angles."""
    arm = _dh_serial_arm(n, seed=3000 + n)
    com = {nm: l._offset for nm, l in arm._links_from_nodes.items()}
    ora = fkjac.ArmOracle(arm._link_joint_sequence, arm.dof, arm.base_pose, com)
    rng = np.random.default_rng(n)
    B = 32 * 9 + 5                                     # several tiles per warp + a ragged one
    q = workloads.random_configs_np(arm, B, seed=n).astype(np.float32)
    q[3, 1] = 2.5e5
    q[40, n - 1] = -7.7e8                              # first tile clean, later tiles not
    q[100:108, :] = (rng.uniform(-1, 1, size=(8, n)) * 1e6).astype(np.float32)
    q[B - 1, 0] = 3.0e30                               # ragged last tile
    q64 = q.astype(np.float64)
    T32, J32 = arm.fk_jacobian(q, "tool")
    assert rel_err(T32, ora.forward_kinematics(q64, "tool")) <= 1e-5
    assert rel_err(J32, ora.jacobian(q64, "tool")) <= 1e-5
    # non-finite angles poison only their own configuration
    q[7, 2] = np.inf
    q[8, 0] = np.nan
    T32b, J32b = arm.fk_jacobian(q, "tool")
    assert np.isnan(T32b[7]).any() and np.isnan(T32b[8]).any()
    keep = np.ones(B, bool)
    keep[[7, 8]] = False
    assert np.array_equal(T32b[keep], T32[keep]) and np.array_equal(J32b[keep], J32[keep])


def test_too_long_path_is_rejected(cuda):
    from numbotics_b200 import _cabi
    arm = _mixed_serial_arm(65, seed=65)
    q = workloads.random_configs_np(arm, 4, seed=0)
    # the chain descriptor is refused as a whole (NBX_MAX_PATH = 64): a status code and a message, not a crash
    with pytest.raises(_cabi.NbxError, match="longer than 64 movable joints"):
        arm.fk_jacobian(q, "tool")


@pytest.mark.parametrize("env", ["NBX_NO_WIDE", "NBX_GENERIC_LEGACY"])
def test_other_kernel_families_in_subprocess(cuda, env):
    """Re-run the chain checks with the wide kernels disabled (-> deep kernel) and with the deep kernel disabled too
    (-> legacy generic kernel).  The switches are read once per process, hence the sub-process."""
    if os.environ.get("NBX_SUBPROCESS_GUARD"):
        pytest.skip("already inside the sub-process run")
    e = dict(os.environ, NBX_SUBPROCESS_GUARD="1")
    e[env] = "1"
    if env == "NBX_GENERIC_LEGACY":
        e["NBX_NO_WIDE"] = "1"
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-m", "gpu", "-x", "-q", "-k",
                        "mixed_serial_chains or allz_serial_chains or dh_chains"], env=e, capture_output=True, text=True, timeout=900,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
