"""Host side of the continuous edge check (reference connectors.py:108-185): the bound on the clearance's rate of change
that `Arm.edges_clear` certifies segments with, held against the float64 oracle; `ContinuousConnector` over an arbitrary
scalar checker (the reference's SLSQP loop) on a case with a closed form."""
import numpy as np
import pytest

from numbotics_b200 import workloads
from numbotics_b200.robots import descriptor
from numbotics_b200.planning.sampling_based import ConnectorParams, ContinuousConnector, DiscreteConnector
from oracle import collision, continuous


@pytest.fixture(scope="module")
def scene():
    arm, cubes = workloads.prm_scene(device=None)
    arm._test_keepalive = cubes
    return arm, collision.SceneOracle(arm)


def test_lever_arms_bound_the_speed_of_primitive_centres(scene):
    arm, so = scene
    rho = descriptor.lever_arms(arm._arrays)
    assert rho.shape == (len(arm._arrays["prim_type"]), arm.dof) and np.all(np.isfinite(rho))
    prim = [arm._arrays["prim_of_link"][l] for l in so.links]
    q = workloads.random_configs_np(arm, 64, seed=5)
    h = 1e-6
    P0 = so.link_poses(q)[:, :, [3, 7, 11]]
    worst = 0.0
    for c in range(arm.dof):
        dq = np.zeros(arm.dof); dq[c] = h
        speed = np.linalg.norm(so.link_poses(q + dq)[:, :, [3, 7, 11]] - P0, axis=2) / h        # (n, links)
        bound = rho[prim, c][None]
        assert np.all(speed <= bound + 1e-6), (c, float((speed - bound).max()))
        worst = max(worst, float((speed / np.maximum(bound, 1e-12)).max()))
    assert worst > 0.3            # not vacuous: some centre moves at a good fraction of its bound


def test_clearance_rate_rows_bound_the_minimum_distance(scene):
    arm, so = scene
    M = descriptor.clearance_rate_rows(arm._arrays, arm.scene_arrays(None))
    assert M.shape[1] == arm.dof and M.shape[0] >= 2
    # a self pair is not charged for the joints above both of its links: some row has zeros in its first columns
    assert np.any((M[:, 0] == 0.0) & (M.sum(axis=1) > 0))
    rng = np.random.default_rng(11)
    q = workloads.random_configs_np(arm, 3000, seed=6)
    for scale in (1e-3, 0.05, 0.3):
        dq = rng.normal(size=q.shape) * scale
        d0, d1 = so.min_distance(q), so.min_distance(q + dq)
        bound = (np.abs(dq) @ M.T).max(axis=1)
        slack = bound - np.abs(d1 - d0)
        assert slack.min() > -1e-9, (scale, float(slack.min()))
    # ... and is not vacuous: for small steps the change reaches a fair share of the bound somewhere
    dq = rng.normal(size=q.shape) * 1e-3
    ratio = np.abs(so.min_distance(q + dq) - so.min_distance(q)) / (np.abs(dq) @ M.T).max(axis=1)
    assert ratio.max() > 0.2


def _disc_checker(centre, radius):
    return lambda x: float(np.linalg.norm(np.asarray(x) - centre) - radius)


def test_continuous_connector_slsqp_path_on_a_disc():
    """A point in the plane and a disc: clearance |x - c| - r.  With a coarse grid the samples of a chord through the
    disc can all be free -- DiscreteConnector accepts the edge, ContinuousConnector (reference :126-145) does not."""
    centre, radius = np.array([0.5, 0.0]), 0.2
    chk = _disc_checker(centre, radius)
    params = ConnectorParams(resolution=0.9, max_distance=0.5, validity_checker=chk)
    cont = ContinuousConnector(params)
    disc = DiscreteConnector(ConnectorParams(resolution=0.9, max_distance=0.5, validity_checker=lambda x: chk(x) > 0.0))
    a, b = np.array([0.0, 0.0]), np.array([1.0, 0.0])          # straight through the disc, samples at t = 0, 0.9, 1 are free
    assert disc.connect(a, b) is not None
    assert cont.connect(a, b) is None
    assert continuous.connect_slsqp(chk, a, b, 0.9) is False
    a2, b2 = np.array([0.0, 0.5]), np.array([1.0, 0.5])        # passes 0.3 above the disc
    out = cont.connect(a2, b2)
    assert out is not None and np.array_equal(out, b2) and out is not b2
    assert continuous.connect_slsqp(chk, a2, b2, 0.9) is True
    # steer stops at max_distance: the first 0.25 of the way is free even on the chord through the disc
    short = ContinuousConnector(ConnectorParams(resolution=0.5, max_distance=0.25, validity_checker=chk))
    out = short.steer(a, b)
    assert out is not None and np.allclose(out, [0.25, 0.0])
    assert cont.steer(a, b) is None                             # 0.5 of the way reaches the disc's rim region: |0.5 - 0.5| - 0.2 < 0
    # degenerate edge and is_valid
    assert cont.connect(a, a.copy()) is None
    assert cont.is_valid(a) and not cont.is_valid(centre)


def test_batched_form_needs_the_arm_checker():
    cont = ContinuousConnector(ConnectorParams(validity_checker=_disc_checker(np.zeros(2), 0.1)))
    with pytest.raises(ValueError, match="ArmClearanceChecker"):
        cont.connect_batch(np.zeros((2, 2)), np.ones((2, 2)))


class _FakeArm:
    """Records what the connector asks of the arm (the device path itself is tests/test_continuous_gpu.py)."""

    def __init__(self, verdict):
        self.verdict = verdict
        self.calls = []

    def edges_clear(self, starts, goals, resolution, max_distance=None, threshold=0.0, max_depth=12, return_detail=False):
        self.calls.append((np.asarray(starts).shape, resolution, max_distance, threshold, return_detail))
        return np.full(np.asarray(starts).shape[:-1], self.verdict)

    def min_distance(self, q):
        q = np.asarray(q)
        return np.float32(0.25) if q.ndim == 1 else np.full(q.shape[:-1], 0.25, np.float32)


def test_arm_checker_routes_to_the_batched_query():
    from numbotics_b200.planning.sampling_based import ArmClearanceChecker
    arm = _FakeArm(True)
    chk = ArmClearanceChecker(arm, threshold=0.05)
    cont = ContinuousConnector(ConnectorParams(resolution=0.1, max_distance=0.5, validity_checker=chk))
    a, b = np.zeros(7), np.full(7, 0.4)
    out = cont.connect(a, b)
    assert np.array_equal(out, b) and out is not b
    assert arm.calls[-1] == ((1, 7), 0.1, None, 0.05, False)             # connect: no max_distance
    out = cont.steer(a, b)                                               # |b - a| = 1.058 > 0.5: stops at T_f = 0.5 / 1.058
    assert arm.calls[-1] == ((1, 7), 0.1, 0.5, 0.05, False)
    assert np.allclose(out, b * (0.5 / np.linalg.norm(b)))
    assert cont.connect_batch(np.zeros((3, 7)), np.ones((3, 7))).shape == (3,)
    assert arm.calls[-1][0] == (3, 7)
    # the checker itself: clearance = distance - threshold, float for one configuration
    assert isinstance(chk(a), float) and abs(chk(a) - 0.2) < 1e-6 and chk(np.zeros((4, 7))).shape == (4,)
    assert cont.is_valid(a)
    # an invalid verdict -> None; a custom distance function leaves the fused path (SLSQP over the checker instead)
    arm.verdict = False
    assert cont.connect(a, b) is None and cont.steer(a, b) is None
    n_calls = len(arm.calls)
    out = cont.connect(a, b, distance_func=lambda x, y: float(np.abs(x - y).max()))
    assert len(arm.calls) == n_calls and out is not None                 # clearance 0.2 everywhere: no feasible point


def test_decision_procedure_on_the_cpu_against_the_dense_oracle(scene):
    """`decide_edges` (the host logic of Arm.edges_clear) with the float64 oracle as its clearance: verdicts equal the
    densely sampled, polished minimum clearance outside a small band, every segment resolved, few evaluations."""
    import torch
    from numbotics_b200.planning.sampling_based.continuous import decide_edges
    arm, so = scene
    rng = np.random.default_rng(21)
    nodes = workloads.random_configs_np(arm, 1500, seed=8)
    pick = rng.choice(1500, 90, replace=False)
    d2 = ((nodes[None] - nodes[pick][:, None]) ** 2).sum(-1)
    near = np.argsort(d2, axis=1)[np.arange(90), 1 + rng.integers(0, 10, 90)]
    starts = np.concatenate([nodes[pick], nodes[:30], nodes[:2]])                  # kNN edges, long edges, two degenerate ones
    goals = np.concatenate([nodes[near], nodes[700:730], nodes[:2]])
    M = descriptor.clearance_rate_rows(arm._arrays, arm.scene_arrays(None))
    S, G = torch.from_numpy(starts), torch.from_numpy(goals)
    calls = []

    def clearance(edge, t):
        calls.append(int(edge.numel()))
        e, tt = edge.numpy(), t.numpy()
        q = starts[e] + tt[:, None] * (goals[e] - starts[e])
        return torch.from_numpy(so.min_distance(q) if len(e) else np.zeros(0))

    for resolution, md in ((0.1, None), (0.9, None), (0.1, 0.7)):
        calls.clear()
        clear, det = decide_edges(S, G, resolution, md, M, clearance)
        clear = clear.numpy()
        assert not clear[-2:].any()                                              # shorter than float32 eps -> None
        c_star = np.array([continuous.min_clearance(so, a, b, resolution, md, per_segment=12)[0]
                           for a, b in zip(starts[:-2], goals[:-2])])
        decided = np.abs(c_star) > 1e-5          # same float64 clearance on both sides: only the oracle's own sampling is in the band
        assert decided.mean() > 0.98
        assert np.array_equal(clear[:-2][decided], c_star[decided] > 0.0), (resolution, md)
        assert det["unresolved_segments"] == 0 and det["levels"] <= 12
        assert det["distance_evaluations"] == sum(calls) and calls[0] == det["grid_samples"]
        # a fine grid needs little refinement; a grid nine times coarser about three more halvings per segment
        assert det["distance_evaluations"] < (5 if resolution == 0.1 else 16) * det["grid_samples"]
        assert np.all(det["min_clearance"].numpy()[:-2] >= c_star - 1e-5)
        assert 0 < (~clear[:-2]).sum() < len(c_star)


def _tree_arm():
    """base -(rev z)-> a -(rev y)-> b -(prismatic x)-> c, and a second branch a -(rev x)-> d: spheres on a, b, c, d."""
    import networkx as nx
    from numbotics_b200.math import trans_mat
    from numbotics_b200.physics import GraphChain, Joint, Constraint, BasicLink, Shape
    from numbotics_b200.robots import Arm
    w = workloads.fresh_world("lever_tree")
    G = nx.DiGraph()
    for n in ("base", "a", "b", "c", "d"):
        G.add_node(n, link=BasicLink(1.0, np.eye(4), Shape.SPHERE, radius=0.05, name=n))
    J = lambda off, ax, t, name: Joint(trans_mat(pos=np.array(off)), np.array(ax, float), t, name=name, lower_limit=-1.0, upper_limit=1.0)
    G.add_edge("base", "a", joint=J([0.0, 0.0, 0.3], [0, 0, 1], Constraint.REVOLUTE, "j_a"))
    G.add_edge("a", "b", joint=J([0.4, 0.0, 0.0], [0, 1, 0], Constraint.REVOLUTE, "j_b"))
    G.add_edge("b", "c", joint=J([0.2, 0.0, 0.0], [1, 0, 0], Constraint.PRISMATIC, "j_c"))
    G.add_edge("a", "d", joint=J([0.0, 0.5, 0.0], [1, 0, 0], Constraint.REVOLUTE, "j_d"))
    return Arm(GraphChain(G, world_name=w.name))


def test_lever_arms_of_a_tree_with_a_slider():
    arm = _tree_arm()
    A = arm._arrays
    rho = descriptor.lever_arms(A)
    prim = {l.name.split(":")[-1]: p for l, p in A["prim_of_link"].items()}          # names are world:chain:link
    col = {name: int(arm._chain._joint_to_index[j]) for j in arm._chain._joints for name in [j.name] if j.type.name != "FIXED"}
    r = 0.05
    # a sits on its own joint: only its radius turns about j_a; nothing else moves it
    assert np.isclose(rho[prim["a"], col["j_a"]], r) and rho[prim["a"]].sum() == rho[prim["a"], col["j_a"]]
    # b: 0.4 from j_a's origin, on j_b's origin
    assert np.isclose(rho[prim["b"], col["j_a"]], 0.4 + r) and np.isclose(rho[prim["b"], col["j_b"]], r)
    assert rho[prim["b"], col["j_d"]] == 0.0 and rho[prim["b"], col["j_c"]] == 0.0
    # c rides the slider: unit speed along it, and an unbounded lever arm about every revolute joint above it
    assert np.isclose(rho[prim["c"], col["j_c"]], 1.0)
    assert np.isinf(rho[prim["c"], col["j_b"]]) and np.isinf(rho[prim["c"], col["j_a"]])
    # d hangs off a on the other branch
    assert np.isclose(rho[prim["d"], col["j_a"]], 0.5 + r) and np.isclose(rho[prim["d"], col["j_d"]], r)
    assert rho[prim["d"], col["j_b"]] == 0.0 and rho[prim["base"]].sum() == 0.0
    # rows: the pair (b, d) is not charged for j_a (above both), and is charged for j_b and j_d
    pairs = np.array([[prim["b"], prim["d"]]], np.int32)
    M = descriptor.clearance_rate_rows(A, {"env_pairs": np.zeros((0, 2), np.int32), "self_pairs": pairs})
    assert M.shape == (1, arm.dof)
    assert M[0, col["j_a"]] == 0.0 and np.isclose(M[0, col["j_b"]], r) and np.isclose(M[0, col["j_d"]], r) and M[0, col["j_c"]] == 0.0
    # the slider makes (c, anything) unbounded: Arm.edges_clear refuses such scenes (NotImplementedError) instead of guessing
    M = descriptor.clearance_rate_rows(A, {"env_pairs": np.array([[prim["c"], 0]], np.int32), "self_pairs": np.zeros((0, 2), np.int32)})
    assert not np.all(np.isfinite(M))


def test_reference_slsqp_loop_agrees_with_the_exact_decision(scene):
    """The reference's own per-segment SLSQP search (connectors.py:126-145, restated in oracle/continuous.py and run with
    scipy, its dependency) over the oracle's clearance, on the edges where it matters -- every invalid edge of a sample and
    the valid ones that come closest to contact: same verdicts as the exact decision the batched path implements.
    (On 80 + 80 random PRM edges at resolutions 0.1 and 0.9 the two agreed on all 160; that run takes minutes.)"""
    arm, so = scene
    rng = np.random.default_rng(5)
    nodes = workloads.random_configs_np(arm, 2000, seed=4)
    pick = rng.choice(2000, 80, replace=False)
    d2 = ((nodes[None] - nodes[pick][:, None]) ** 2).sum(-1)
    near = np.argsort(d2, axis=1)[np.arange(80), 1 + rng.integers(0, 10, 80)]
    S, G = nodes[pick], nodes[near]
    res = 0.3
    c_star = np.array([continuous.min_clearance(so, a, b, res, per_segment=12)[0] for a, b in zip(S, G)])
    invalid = np.nonzero(c_star < -1e-4)[0][:4]
    tight = np.argsort(np.where(c_star > 1e-4, c_star, np.inf))[:4]
    assert len(invalid) >= 2
    chk = lambda q: float(so.min_distance(np.asarray(q)[None])[0])
    for e in np.concatenate([invalid, tight]):
        assert continuous.connect_slsqp(chk, S[e], G[e], res) == bool(c_star[e] > 0.0), (int(e), float(c_star[e]))
