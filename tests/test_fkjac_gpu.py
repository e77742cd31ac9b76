"""CUDA FK / Jacobian vs the reference's golden vectors and the float64 oracle (through the C-ABI)."""
import numpy as np
import pytest

from numbotics_b200 import workloads
from oracle import fkjac

pytestmark = pytest.mark.gpu

RTOL32 = 1e-5   # north star: 1e-5 relative (fp32), relative to max(|x|, 1) per matrix
RTOL64 = 1e-12


def rel_err(a, b):
    """max over the batch of ||a-b||_max / max(||b||_max, 1) per matrix."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    num = np.abs(a - b).reshape(a.shape[0], -1).max(axis=1)
    den = np.maximum(np.abs(b).reshape(b.shape[0], -1).max(axis=1), 1.0)
    return float((num / den).max())


@pytest.fixture(scope="module")
def kinova(cuda):
    return workloads.kinova_arm(device=0)


@pytest.mark.parametrize("frame", ["tool_frame", "end_effector_link", "camera", "gripper"])
@pytest.mark.parametrize("dtype,tol", [(np.float32, RTOL32), (np.float64, RTOL64)])
def test_golden_numpy_path(kinova, golden_kinova, frame, dtype, tol):
    g = golden_kinova
    q = g["q"].astype(dtype)
    T = kinova.forward_kinematics(q, frame)
    J = kinova.jacobian(q, frame)
    assert T.dtype == dtype and J.dtype == dtype
    assert T.shape == (q.shape[0], 4, 4) and J.shape == (q.shape[0], 6, 7)
    assert rel_err(T, g[f"{frame}/T"]) <= tol
    assert rel_err(J, g[f"{frame}/J"]) <= tol
    Tc = kinova.forward_kinematics(q, frame, use_com=True)
    Jc = kinova.jacobian(q, frame, use_com=True)
    assert rel_err(Tc, g[f"{frame}/T_com"]) <= tol
    assert rel_err(Jc, g[f"{frame}/J_com"]) <= tol


@pytest.mark.parametrize("tdtype,tol", [("float32", RTOL32), ("float64", RTOL64)])
def test_golden_torch_cuda_path(kinova, golden_kinova, cuda, tdtype, tol):
    import torch
    g = golden_kinova
    dt = getattr(torch, tdtype)
    q = torch.as_tensor(g["q"], dtype=dt, device=cuda)
    T, J = kinova.fk_jacobian(q, "tool_frame")
    assert T.is_cuda and J.is_cuda and T.dtype == dt
    assert rel_err(T.cpu().numpy(), g["tool_frame/T"]) <= tol
    assert rel_err(J.cpu().numpy(), g["tool_frame/J"]) <= tol
    lp = torch.as_tensor(g["local_pose"], dtype=dt, device=cuda)
    gp = torch.as_tensor(g["global_pose"], dtype=dt, device=cuda)
    assert rel_err(kinova.forward_kinematics(q, "tool_frame", local_pose=lp).cpu().numpy(), g["tool_frame/T_lp"]) <= tol
    assert rel_err(kinova.jacobian(q, "tool_frame", local_pose=lp).cpu().numpy(), g["tool_frame/J_lp"]) <= tol
    assert rel_err(kinova.jacobian(q, "tool_frame", global_pose=gp).cpu().numpy(), g["tool_frame/J_gp"]) <= tol
    assert rel_err(kinova.forward_kinematics(q, "tool_frame", local_pose=lp[0]).cpu().numpy(), g["tool_frame/T_lp1"]) <= tol


def test_shapes_and_batch_dims(kinova, golden_kinova):
    g = golden_kinova
    q = g["q"][:24]
    T1 = kinova.forward_kinematics(q[0], "tool_frame")
    J1 = kinova.jacobian(q[0], "tool_frame")
    assert T1.shape == (4, 4) and J1.shape == (6, 7)
    assert rel_err(T1[None], g["tool_frame/T"][:1]) <= RTOL64
    q3 = q.reshape(2, 3, 4, 7)
    T3 = kinova.forward_kinematics(q3, "tool_frame")
    J3 = kinova.jacobian(q3, "tool_frame")
    assert T3.shape == (2, 3, 4, 4, 4) and J3.shape == (2, 3, 4, 6, 7)
    assert rel_err(T3.reshape(-1, 4, 4), g["tool_frame/T"][:24]) <= RTOL64
    assert rel_err(J3.reshape(-1, 6, 7), g["tool_frame/J"][:24]) <= RTOL64
    # bottom row is exactly [0 0 0 1]
    assert np.array_equal(T3[..., 3, :], np.broadcast_to([0.0, 0.0, 0.0, 1.0], T3[..., 3, :].shape))


def test_errors_match_reference(kinova):
    q = np.zeros((3, 7))
    with pytest.raises(ValueError, match="Frame nope not found in chain"):
        kinova.forward_kinematics(q, "nope")
    with pytest.raises(ValueError, match="q must have 7 elements"):
        kinova.forward_kinematics(np.zeros((3, 6)), "tool_frame")
    with pytest.raises(ValueError, match="local_pose must be a 4x4 matrix"):
        kinova.forward_kinematics(q, "tool_frame", local_pose=np.zeros((3, 3)))
    with pytest.raises(ValueError, match="local_pose must have the same batch dimensions as q"):
        kinova.forward_kinematics(q, "tool_frame", local_pose=np.zeros((2, 4, 4)))
    with pytest.raises(ValueError, match="local_pose and global_pose cannot both be provided"):
        kinova.jacobian(q, "tool_frame", local_pose=np.eye(4), global_pose=np.eye(4))
    with pytest.raises(ValueError, match="global_pose must be a 4x4 matrix"):
        kinova.jacobian(q, "tool_frame", global_pose=np.zeros((3, 3)))


def test_every_frame_against_oracle(kinova):
    """All 16 link frames (dense and sparse paths, frames ending in a movable joint included) vs the
    float64 oracle in its geometrically-correct mode."""
    q = workloads.random_configs_np(kinova, 1000 + 37, seed=3)   # ragged last tile
    com = {n: l._offset for n, l in kinova._links_from_nodes.items()}
    ora = fkjac.ArmOracle(kinova._link_joint_sequence, 7, kinova.base_pose, com)
    for frame in kinova._link_joint_sequence:
        To, Jo = ora.forward_kinematics(q, frame), ora.jacobian(q, frame)
        T64, J64 = kinova.fk_jacobian(q, frame)
        assert rel_err(T64, To) <= RTOL64, frame
        assert rel_err(J64, Jo) <= RTOL64, frame
        T32, J32 = kinova.fk_jacobian(q.astype(np.float32), frame)
        assert rel_err(T32, To) <= RTOL32, frame
        assert rel_err(J32, Jo) <= RTOL32, frame


def test_serial30_golden(cuda, golden_serial30):
    g = golden_serial30
    arm = workloads.serial_arm(30, device=0)
    for dtype, tol in ((np.float64, RTOL64), (np.float32, 1e-5)):
        # 30 chained fp32 rotations stay inside north_star's 1e-5 (measured 5.5e-7, profiles/r02_fp32_error_vs_depth.txt)
        T, J = arm.fk_jacobian(g["q"].astype(dtype), "tool")
        assert rel_err(T, g["tool/T"]) <= tol
        assert rel_err(J, g["tool/J"]) <= tol


def test_large_batch_properties(kinova, cuda):
    """BASELINE-size style checks that do not need the oracle at full size: rigid-body invariants and
    agreement between the one-pass fused call and the two separate calls."""
    import torch
    q = workloads.random_configs_cuda(kinova, 1 << 20, seed=5)
    T, J = kinova.fk_jacobian(q, "tool_frame")
    R = T[:, :3, :3]
    eye = torch.eye(3, device=cuda).expand_as(R)
    assert float((R @ R.transpose(1, 2) - eye).abs().max()) < 2e-5
    assert float((torch.linalg.det(R) - 1).abs().max()) < 2e-5
    T2 = kinova.forward_kinematics(q, "tool_frame")
    J2 = kinova.jacobian(q, "tool_frame")
    assert torch.equal(T, T2) and torch.equal(J, J2)
    # angular rows of a revolute column are unit vectors
    assert float((J[:, 3:, :].norm(dim=1) - 1).abs().max()) < 2e-5
    # spot check 4096 rows against the oracle
    idx = torch.randint(0, q.shape[0], (4096,), device=cuda)
    com = {n: l._offset for n, l in kinova._links_from_nodes.items()}
    ora = fkjac.ArmOracle(kinova._link_joint_sequence, 7, kinova.base_pose, com)
    qs = q[idx].double().cpu().numpy()
    assert rel_err(T[idx].cpu().numpy(), ora.forward_kinematics(qs, "tool_frame")) <= RTOL32
    assert rel_err(J[idx].cpu().numpy(), ora.jacobian(qs, "tool_frame")) <= RTOL32


def _random_tree_arm(seed, n_joints=9, device=0):
    """Random branching tree: canonical (+-x/y/z), oblique and NON-unit revolute axes, prismatic and fixed joints."""
    import networkx as nx
    from numbotics_b200.math import trans_mat
    from numbotics_b200.physics import GraphChain, Joint, Constraint, DummyLink
    from numbotics_b200.robots import Arm
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(seed)
    w = workloads.fresh_world("rtree")
    G = nx.DiGraph()
    G.add_node("n0", link=DummyLink(1.0, trans_mat(pos=rng.normal(size=3) * 0.1), name="n0"))
    axes = [[0, 0, 1], [0, 0, -1], [1, 0, 0], [0, -1, 0], [1, 1, 0], [0.3, -0.5, 0.8], [0, 0, 2.0], [0, 1, 0]]
    for i in range(1, n_joints + 1):
        parent = f"n{int(rng.integers(max(0, i - 3), i))}"
        off = trans_mat(pos=rng.normal(size=3) * 0.2, orn=Rotation.from_rotvec(rng.normal(size=3)).as_matrix())
        kind = rng.choice(["rev", "rev", "rev", "pris", "fixed"])
        ax = np.array(axes[int(rng.integers(0, len(axes)))], float)
        if kind in ("rev",) and np.linalg.norm(ax) not in (1.0, 2.0):
            ax = ax / np.linalg.norm(ax)
        ctype = {"rev": Constraint.REVOLUTE, "pris": Constraint.PRISMATIC, "fixed": Constraint.FIXED}[kind]
        G.add_node(f"n{i}", link=DummyLink(1.0, trans_mat(pos=rng.normal(size=3) * 0.05), name=f"n{i}"))
        G.add_edge(parent, f"n{i}", joint=Joint(off, ax, ctype, name=f"j{i}", lower_limit=-2.5, upper_limit=2.5))
    return Arm(GraphChain(G, world_name=w.name), device=device)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_random_trees_all_joint_kinds(cuda, seed):
    arm = _random_tree_arm(seed)
    if arm.dof == 0:
        pytest.skip("no movable joint drawn")
    q = workloads.random_configs_np(arm, 300, seed=seed)
    com = {n: l._offset for n, l in arm._links_from_nodes.items()}
    ora = fkjac.ArmOracle(arm._link_joint_sequence, arm.dof, arm.base_pose, com)
    for frame in arm._link_joint_sequence:
        for use_com in (False, True):
            T64, J64 = arm.fk_jacobian(q, frame, use_com=use_com)
            assert rel_err(T64, ora.forward_kinematics(q, frame, use_com=use_com)) <= RTOL64, frame
            assert rel_err(J64, ora.jacobian(q, frame, use_com=use_com)) <= RTOL64, frame
        T32, J32 = arm.fk_jacobian(q.astype(np.float32), frame)
        assert rel_err(T32, ora.forward_kinematics(q, frame)) <= RTOL32, frame
        assert rel_err(J32, ora.jacobian(q, frame)) <= RTOL32, frame


def test_large_angles_sincos_range(kinova):
    """The custom fp32 sincos must hold 1e-5 well outside [-pi, pi] (continuous joints wind up)."""
    rng = np.random.default_rng(9)
    q = rng.uniform(-200.0, 200.0, size=(2000, 7))
    q[:5] *= 1e4                      # beyond the Cody-Waite range: sincosf() path
    q32 = q.astype(np.float32)
    ora = fkjac.ArmOracle(kinova._link_joint_sequence, 7, kinova.base_pose)
    T, J = kinova.fk_jacobian(q32, "tool_frame")
    q_as_seen = q32.astype(np.float64)
    assert rel_err(T, ora.forward_kinematics(q_as_seen, "tool_frame")) <= RTOL32
    assert rel_err(J, ora.jacobian(q_as_seen, "tool_frame")) <= RTOL32


def test_empty_unaligned_and_noncontiguous_inputs(kinova, cuda):
    """Edge cases of the boundary: B = 0, tensors that are views at odd offsets (pointers not 16-byte aligned:
    the kernels must drop to their non-TMA path), non-contiguous inputs, and every ragged size around a tile."""
    import torch
    ora = fkjac.ArmOracle(kinova._link_joint_sequence, 7, kinova.base_pose)
    T0, J0 = kinova.fk_jacobian(np.zeros((0, 7)), "tool_frame")
    assert T0.shape == (0, 4, 4) and J0.shape == (0, 6, 7)
    T0, J0 = kinova.fk_jacobian(torch.zeros((0, 7), device=cuda), "tool_frame")
    assert T0.shape == (0, 4, 4) and J0.shape == (0, 6, 7)
    q = workloads.random_configs_cuda(kinova, 1000, seed=11)
    To, Jo = ora.forward_kinematics(q.double().cpu().numpy(), "tool_frame"), ora.jacobian(q.double().cpu().numpy(), "tool_frame")
    # misaligned q: a view starting one float into a larger buffer
    buf = torch.empty(1000 * 7 + 1, device=cuda)
    qv = buf[1:].view(1000, 7)
    qv.copy_(q)
    assert qv.data_ptr() % 16 != 0
    T, J = kinova.fk_jacobian(qv, "tool_frame")
    assert rel_err(T.cpu().numpy(), To) <= RTOL32 and rel_err(J.cpu().numpy(), Jo) <= RTOL32
    # non-contiguous q (every other row of a wider tensor)
    wide = torch.zeros((2000, 7), device=cuda)
    wide[::2] = q
    T, J = kinova.fk_jacobian(wide[::2], "tool_frame")
    assert rel_err(T.cpu().numpy(), To) <= RTOL32 and rel_err(J.cpu().numpy(), Jo) <= RTOL32
    # ragged sizes around one and two tiles (two consecutive tiles per CTA)
    for n in (1, 127, 128, 129, 255, 256, 257, 383):
        T, J = kinova.fk_jacobian(q[:n], "tool_frame")
        assert T.shape == (n, 4, 4)
        assert rel_err(T.cpu().numpy(), To[:n]) <= RTOL32 and rel_err(J.cpu().numpy(), Jo[:n]) <= RTOL32
    # sparse-column frame through the generic kernel, ragged
    Tf, Jf = kinova.fk_jacobian(q[:333], "forearm_link")
    assert rel_err(Tf.cpu().numpy(), ora.forward_kinematics(q[:333].double().cpu().numpy(), "forearm_link")) <= RTOL32
    assert rel_err(Jf.cpu().numpy(), ora.jacobian(q[:333].double().cpu().numpy(), "forearm_link")) <= RTOL32
    assert float(Jf[:, :, 4:].abs().max()) == 0.0     # joints past the forearm do not move it


def test_cabi_rejects_bad_arguments(kinova, cuda):
    """Status codes instead of crashes: the C-ABI validates what the Python layer cannot know."""
    import torch
    from numbotics_b200 import _cabi
    lib = _cabi.load()
    h = kinova._handle()
    keep, base = _cabi.base_pose_arg(kinova.base_pose)
    q = torch.zeros((4, 7), device=cuda)
    T = torch.empty((4, 4, 4), device=cuda)
    assert lib.nbx_fk_jac_f32(h.ptr, 999, q.data_ptr(), 4, base, 0, None, 0, None, 0, T.data_ptr(), None, None) == -1
    assert "frame id" in _cabi.last_error()
    assert lib.nbx_fk_jac_f32(h.ptr, 0, q.data_ptr(), -1, base, 0, None, 0, None, 0, T.data_ptr(), None, None) == -1
    assert lib.nbx_fk_jac_f32(h.ptr, 0, None, 4, base, 0, None, 0, None, 0, T.data_ptr(), None, None) == -1
    assert lib.nbx_fk_jac_f32(None, 0, q.data_ptr(), 4, base, 0, None, 0, None, 0, T.data_ptr(), None, None) == -1
    with pytest.raises(_cabi.NbxError, match="frame id"):
        _cabi.check(lib.nbx_fk_jac_f32(h.ptr, -3, q.data_ptr(), 4, base, 0, None, 0, None, 0, T.data_ptr(), None, None))


def test_concurrent_host_threads(kinova):
    """Handles are immutable and the host-buffer pipeline is guarded: several host threads calling at once (NumPy in,
    NumPy out, and device tensors on their own streams) get the bits of a serial call."""
    import threading
    import torch
    q = workloads.random_configs_np(kinova, 40000, seed=17).astype(np.float32)
    T0, J0 = kinova.fk_jacobian(q, "tool_frame")
    qd = torch.as_tensor(q, device="cuda:0")
    ok = [False] * 6

    def host_worker(i):
        good = True
        for _ in range(4):
            T, J = kinova.fk_jacobian(q[i::3], "tool_frame")
            good = good and np.array_equal(T, T0[i::3]) and np.array_equal(J, J0[i::3])
        ok[i] = good

    def device_worker(i):
        good = True
        with torch.cuda.stream(torch.cuda.Stream(device="cuda:0")):
            for _ in range(4):
                T, J = kinova.fk_jacobian(qd[(i - 3)::3].contiguous(), "end_effector_link" if i == 5 else "tool_frame")
                torch.cuda.current_stream().synchronize()
                if i != 5:
                    good = good and np.array_equal(T.cpu().numpy(), T0[(i - 3)::3]) and np.array_equal(J.cpu().numpy(), J0[(i - 3)::3])
        ok[i] = good

    threads = [threading.Thread(target=host_worker, args=(i,)) for i in range(3)] + \
              [threading.Thread(target=device_worker, args=(i,)) for i in range(3, 6)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert all(ok), ok
