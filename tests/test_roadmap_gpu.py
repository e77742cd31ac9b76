"""Batched PRM construction: same edge set as the reference's sequential loop."""
import numpy as np
import pytest

from numbotics_b200 import workloads
from numbotics_b200.planning.sampling_based import ConnectorParams, DiscreteConnector, ArmValidityChecker, BatchedPRM, knn_earlier
from oracle import collision

pytestmark = pytest.mark.gpu


def test_knn_earlier_matches_bruteforce(cuda):
    import torch
    x = torch.rand((700, 7), device=cuda)
    idx, ok = knn_earlier(x, 5, chunk=256)
    xn = x.cpu().numpy().astype(np.float64)
    for i in (0, 1, 3, 5, 6, 100, 699):
        if i == 0:
            assert not ok[0].any()
            continue
        d = np.linalg.norm(xn[:i] - xn[i], axis=1)
        want = set(np.argsort(d)[: min(5, i)].tolist())
        got = set(idx[i][ok[i]].cpu().numpy().tolist())
        assert got == want and int(ok[i].sum()) == min(5, i)


def test_roadmap_equals_sequential_prm_loop(cuda):
    arm, cubes = workloads.prm_scene(device=0)
    so = collision.SceneOracle(arm)
    conn = DiscreteConnector(ConnectorParams(resolution=0.1, max_distance=np.pi, validity_checker=ArmValidityChecker(arm)))
    lo, hi = workloads.sample_limits(arm.joint_limits)
    prm = BatchedPRM(conn, lo, hi, k_nearest=6)
    start = np.zeros(7)
    goal = np.array([0.0, np.pi / 2.0, 0, 0, 0, 0, 0])          # reference tests/_test_rrt.py:137-138
    rm = prm.build(300, start=start, goals=[goal], seed=3)
    nodes = rm.nodes.cpu().numpy().astype(np.float64)
    assert nodes.shape == (302, 7) and np.allclose(nodes[0], start) and np.allclose(nodes[1], goal, atol=1e-6)
    # the sequential loop of planners/prm.py:35-47 on the same vertices, validity from the oracle
    want = set()
    borderline = set()
    for i in range(1, nodes.shape[0]):
        d = np.linalg.norm(nodes[:i] - nodes[i], axis=1)
        for j in np.argsort(d)[: min(6, i)]:
            v, _, md = so.edges_valid(nodes[j][None], nodes[i][None], 0.1, return_detail=True)
            if abs(md[0]) <= 1e-4:
                borderline.add((int(j), i))
            elif v[0]:
                want.add((int(j), i))
    got = {tuple(e) for e in rm.edges.cpu().numpy().tolist()}
    assert got - borderline == want - borderline
    assert rm.candidates == sum(min(6, i) for i in range(1, nodes.shape[0]))
    # node validity and path query
    assert np.array_equal(rm.node_valid.cpu().numpy(), ~so.in_collision(nodes)) or np.abs(so.min_distance(nodes)).min() < 1e-4
    path = rm.shortest_path(0, 1)
    assert path is not None and path[0] == 0 and path[-1] == 1
    for a, b in zip(path[:-1], path[1:]):
        assert (min(a, b), max(a, b)) in got
    G = rm.to_networkx()
    assert G.number_of_nodes() == 302 and G.number_of_edges() == len(got)
