"""`Arm.edges_clear` / `ContinuousConnector` with an `ArmClearanceChecker` (batched form of reference
connectors.py:108-185) against the float64 oracle: the clearance sampled densely inside every grid segment and polished
(`oracle/continuous.py`).  An edge is compared when the oracle's smallest clearance is outside the stated band."""
import numpy as np
import pytest
import torch

from numbotics_b200 import workloads
from numbotics_b200.planning.sampling_based import (ConnectorParams, ContinuousConnector, DiscreteConnector,
                                                    ArmClearanceChecker, ArmValidityChecker)
from oracle import collision, continuous

pytestmark = pytest.mark.gpu

BAND = 3e-4      # m: 1e-4 stated accuracy of the float32 distance query + the oracle's own sampling


@pytest.fixture(scope="module")
def scene(cuda):
    arm, cubes = workloads.prm_scene(device=0)
    arm._test_keepalive = cubes
    return arm, collision.SceneOracle(arm)


def _knn_edges(arm, n_nodes, E, seed):
    rng = np.random.default_rng(seed)
    nodes = workloads.random_configs_np(arm, n_nodes, seed=seed).astype(np.float32)
    idx = rng.choice(n_nodes, E, replace=False)
    d2 = ((nodes[None, :, :] - nodes[idx][:, None, :]) ** 2).sum(-1)
    order = np.argsort(d2, axis=1)
    pick = order[np.arange(E), 1 + rng.integers(0, 10, E)]
    return nodes[idx], nodes[pick]


def _oracle_min(so, starts, goals, resolution, max_distance=None, threshold=0.0, per_segment=16):
    return np.array([continuous.min_clearance(so, s, g, resolution, max_distance, threshold, per_segment)[0]
                     for s, g in zip(starts.astype(np.float64), goals.astype(np.float64))])


def test_edges_clear_matches_the_dense_oracle(scene):
    arm, so = scene
    starts, goals = _knn_edges(arm, 3000, 300, seed=1)
    clear, det = arm.edges_clear(starts, goals, 0.1, return_detail=True)
    assert clear.dtype == np.bool_ and clear.shape == (300,)
    c_star = _oracle_min(so, starts, goals, 0.1)
    decided = np.abs(c_star) > BAND
    assert decided.mean() > 0.97, f"{1.0 - decided.mean():.4f} of the edges are inside the contact band"
    assert np.array_equal(clear[decided], c_star[decided] > 0.0), np.nonzero(clear[decided] != (c_star[decided] > 0.0))[0]
    assert 0.02 < (~clear).mean() < 0.5                          # both verdicts occur
    assert det["unresolved_segments"] == 0
    assert det["distance_evaluations"] < 4 * det["grid_samples"]
    # the smallest clearance the query saw is a clearance of the edge: never below the true minimum by more than the band
    assert np.all(det["min_clearance"] >= c_star - BAND)
    assert np.all(det["min_clearance"][~clear] <= 0.0)


def test_coarse_grid_edges_the_discrete_check_lets_through(scene):
    """resolution 0.9: a handful of samples per edge.  Edges whose samples are all free but which pass THROUGH an obstacle
    or a link in between are accepted by DiscreteConnector and must be rejected here."""
    arm, so = scene
    nodes = workloads.random_configs_np(arm, 1200, seed=2).astype(np.float32)
    starts, goals = nodes[:600], nodes[600:]
    disc = arm.edges_valid(starts, goals, 0.9)
    clear, det = arm.edges_clear(starts, goals, 0.9, return_detail=True)
    c_star = _oracle_min(so, starts, goals, 0.9, per_segment=12)
    decided = np.abs(c_star) > BAND
    assert np.array_equal(clear[decided], c_star[decided] > 0.0)
    missed = disc & (c_star < -1e-3)
    assert missed.sum() >= 5 and not clear[missed].any()
    # continuous is the stricter test: a clear edge is a valid discrete edge (up to grid samples inside the band)
    odd = clear & ~disc
    assert np.all(det["min_clearance"][odd] < BAND), det["min_clearance"][odd]
    assert det["unresolved_segments"] == 0


def test_threshold_and_steer_semantics(scene):
    arm, so = scene
    starts, goals = _knn_edges(arm, 2000, 120, seed=7)
    goals = starts + 3.0 * (goals - starts)                      # long enough for max_distance to bite
    thr, md = 0.02, 0.6
    clear = arm.edges_clear(starts, goals, 0.1, max_distance=md, threshold=thr)
    c_star = _oracle_min(so, starts, goals, 0.1, max_distance=md, threshold=thr)
    decided = np.abs(c_star) > BAND
    assert decided.mean() > 0.95
    assert np.array_equal(clear[decided], c_star[decided] > 0.0)
    full = arm.edges_clear(starts, goals, 0.1, threshold=thr)    # the whole edge is at least as hard as its first part
    assert not np.any(full & ~clear & decided)


def test_connector_surface(scene):
    arm, so = scene
    starts, goals = _knn_edges(arm, 2000, 40, seed=9)
    chk = ArmClearanceChecker(arm)
    cont = ContinuousConnector(ConnectorParams(resolution=0.1, max_distance=0.3, validity_checker=chk))
    batch = cont.connect_batch(starts, goals)
    for e in range(12):
        out = cont.connect(starts[e], goals[e])
        assert (out is not None) == bool(batch[e])
        if out is not None:
            assert np.array_equal(out, goals[e]) and out is not goals[e]
    sb = cont.steer_batch(starts, goals)
    for e in range(6):
        out = cont.steer(starts[e].astype(np.float64), goals[e].astype(np.float64))
        assert (out is not None) == bool(sb[e])
        if out is not None:
            s64, g64 = starts[e].astype(np.float64), goals[e].astype(np.float64)
            t_f = min(1.0, 0.3 / np.linalg.norm(g64 - s64))
            assert np.allclose(out, (1 - t_f) * s64 + t_f * g64, atol=1e-12)
    # the checker: one configuration -> float, a batch -> array; is_valid as the reference (:184-185)
    c0 = chk(starts[0])
    assert isinstance(c0, float) and abs(c0 - float(so.min_distance(starts[0].astype(np.float64)))) < 2e-4
    assert cont.is_valid(starts[0]) == (c0 > 0.0)
    assert chk(starts[:5]).shape == (5,)
    # torch in -> torch out on the same device; degenerate edges are invalid
    ts, tg = torch.from_numpy(starts).cuda(), torch.from_numpy(goals).cuda()
    tb = cont.connect_batch(ts, tg)
    assert tb.is_cuda and tb.dtype == torch.bool and np.array_equal(tb.cpu().numpy(), batch)
    assert not cont.connect_batch(starts[:3], starts[:3].copy()).any()
    with pytest.raises(ValueError, match="ArmClearanceChecker"):
        ContinuousConnector(ConnectorParams(validity_checker=ArmValidityChecker(arm))).connect_batch(starts, goals)
