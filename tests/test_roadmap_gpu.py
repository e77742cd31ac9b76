"""Batched PRM construction: same edge set as the reference's sequential loop."""
import numpy as np
import pytest

from numbotics_b200 import workloads
from numbotics_b200.planning.sampling_based import ConnectorParams, DiscreteConnector, ArmValidityChecker, BatchedPRM, knn_earlier
from oracle import collision

pytestmark = pytest.mark.gpu


# (4000, 7, 10), (2500, 8, 10): more than four tiles per query block, so the slices' agreement on the k-th best distance runs
# (8 dimensions: no spare component, no norm filter); offset 50: points far from the origin, where the norm filter's
# a.b - |b|^2 / 2 loses five digits to cancellation and has to stay conservative
@pytest.mark.parametrize("N,dof,k,offset", [(700, 7, 5, 0.0), (1500, 7, 10, 0.0), (900, 3, 16, 0.0), (600, 30, 20, 0.0), (400, 64, 32, 0.0),
                                            (5, 7, 10, 0.0), (4000, 7, 10, 0.0), (2500, 8, 10, 0.0), (3000, 7, 10, 50.0), (2000, 16, 24, -3.0)])
def test_knn_earlier_matches_bruteforce(cuda, N, dof, k, offset):
    """Exact L2 kNN among the EARLIER vertices (the reference's incremental index): every row against NumPy."""
    import torch
    g = torch.Generator(device=cuda); g.manual_seed(N + dof)
    x = torch.rand((N, dof), device=cuda, generator=g) + offset
    idx, ok = knn_earlier(x, k)
    assert idx.shape == (N, k) and idx.dtype == torch.long and ok.dtype == torch.bool
    xn = x.cpu().numpy().astype(np.float64)
    idx_h, ok_h = idx.cpu().numpy(), ok.cpu().numpy()
    assert not ok_h[0].any()
    for i in range(1, N):
        d = np.linalg.norm(xn[:i] - xn[i], axis=1)
        n = min(k, i)
        assert int(ok_h[i].sum()) == n and ok_h[i, :n].all()
        got = idx_h[i, :n]
        assert len(set(got.tolist())) == n and got.max() < i
        # same distances as the true k smallest (sets can differ only between fp32-equal distances), ascending order
        atol = 1e-6 if offset == 0.0 else 2e-5          # float32 coordinates near 50 are 4e-6 apart
        assert np.allclose(np.sort(d)[:n], d[got], rtol=0, atol=atol)
        assert np.all(np.diff(d[got]) >= -atol)


def test_nearest_neighbors_index_mirror(cuda):
    """`ApproximateNearestNeighborsIndex` (reference math/geometry/nearest_neighbors.py:6-85): ids, removal, shapes."""
    from numbotics_b200.math.geometry import ApproximateNearestNeighborsIndex
    rng = np.random.default_rng(0)
    pts = rng.uniform(-1, 1, size=(500, 7))
    index = ApproximateNearestNeighborsIndex(7)
    assert len(index) == 0
    ids = index.add_points(pts[:400])
    assert ids == list(range(400)) and index.add_point(pts[400]) == 400 and len(index) == 401
    q = rng.uniform(-1, 1, size=7)
    d = np.linalg.norm(pts[:401] - q, axis=1)
    item, label = index.nearest(q, return_labels=True)
    assert label == int(np.argmin(d)) and np.array_equal(item, pts[label]) and item.shape == (7,)
    items, labels = index.k_nearest(q, 6, return_labels=True)
    assert labels.tolist() == np.argsort(d)[:6].tolist() and items.shape == (6, 7)
    index.remove_points(labels[:3].tolist())
    assert len(index) == 398
    items2, labels2 = index.k_nearest(q, 3, return_labels=True)
    assert labels2.tolist() == np.argsort(d)[3:6].tolist()          # ids survive the removal
    lab, d2 = index.query_batch(pts[450:470], 4)
    assert lab.shape == (20, 4) and d2.shape == (20, 4) and np.all(np.diff(d2, axis=1) >= 0)
    with pytest.raises(ValueError, match="does not match index dimension"):
        index.add_points(np.zeros((2, 6)))
    with pytest.raises(ValueError, match="Query dimension 6 does not match index dimension 7"):
        index.query(np.zeros(6))
    with pytest.raises(ValueError, match="Point should be a 1D array"):
        index.add_point(np.zeros((2, 7)))
    small = ApproximateNearestNeighborsIndex(7)
    small.add_points(pts[:2])
    assert small.k_nearest(q, 5).shape == (2, 7)                     # k is clipped to the number of points


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


def test_prm_star_and_the_reference_self_slot(cuda):
    """PRM* (reference planners/prm_star.py:20-57): of the k nearest earlier vertices only those strictly inside the
    connection radius of the graph at insertion time are tried; and `reference_self_slot`: the reference's index already
    holds the new vertex when it is queried, so k answers are k - 1 earlier vertices plus the vertex itself, which
    `connect` rejects."""
    import scipy.special
    from numbotics_b200.planning.sampling_based.roadmap import prm_star_radius
    arm, cubes = workloads.prm_scene(device=0)
    so = collision.SceneOracle(arm)
    conn = DiscreteConnector(ConnectorParams(resolution=0.1, max_distance=np.pi, validity_checker=ArmValidityChecker(arm)))
    # a small box around a free configuration, so that the PRM* radius (which shrinks like (ln n / n)^(1/7)) finds neighbours
    centre = np.array([0.3, 0.6, -0.2, 1.2, 0.1, 0.8, 0.0])
    lo, hi = centre - 0.25, centre + 0.25
    # the radius formula itself, against the reference's expression (prm_star.py:20-26)
    dim, k, V = 7.0, 123.0, float(np.prod(hi - lo))
    v_ball = (np.pi ** (dim / 2.0)) / scipy.special.gamma((dim / 2.0) + 1.0)
    gamma = 2.0 * (1.0 + (1.0 / dim)) ** (1.0 / dim) * (v_ball / V) ** (1.0 / dim)
    assert abs(prm_star_radius(123, 7, V) - gamma * (np.log(k) / k) ** (1 / dim)) < 1e-15
    for star, self_slot in ((True, False), (False, True), (True, True)):
        prm = BatchedPRM(conn, lo, hi, k_nearest=5, star=star, reference_self_slot=self_slot)
        rm = prm.build(250, seed=11)
        nodes = rm.nodes.cpu().numpy().astype(np.float64)
        k_eff = 4 if self_slot else 5
        want, borderline, tried = set(), set(), 0
        for i in range(1, nodes.shape[0]):
            d = np.linalg.norm(nodes[:i] - nodes[i], axis=1)
            radius = prm_star_radius(i + 1, 7, V) if star else np.inf
            for j in np.argsort(d, kind="stable")[: min(k_eff, i)]:
                if abs(d[j] - radius) < 1e-6 * max(radius, 1e-9) if star else False:
                    borderline.add((int(j), i))               # float32 distance against a float64 radius
                    continue
                if not d[j] < radius:
                    continue
                tried += 1
                v, _, md = so.edges_valid(nodes[j][None], nodes[i][None], 0.1, return_detail=True)
                if abs(md[0]) <= 1e-4:
                    borderline.add((int(j), i))
                elif v[0]:
                    want.add((int(j), i))
        got = {tuple(e) for e in rm.edges.cpu().numpy().tolist()}
        assert got - borderline == want - borderline, (star, self_slot)
        assert len(want) > 50
        if star:
            assert abs(rm.candidates - tried) <= len(borderline)
