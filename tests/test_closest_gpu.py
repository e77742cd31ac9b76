"""Batched Arm.closest_to / jacobian_proximity for the closest pair against the float64
oracle: distance and pair identity from `SceneOracle.pair_distances`, witness points by the point-to-shape distance of
the same C narrow phase, the gradient by central differences of the oracle's minimum distance (the envelope theorem:
d(min distance)/dq = n . (J_v(subject point) - J_v(target point)), reference robots/arm.py:621-632)."""
import numpy as np
import pytest

from numbotics_b200 import workloads
from numbotics_b200.robots.arm import Proximity, Proximities
from oracle import collision

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scene(cuda):
    arm, cubes = workloads.prm_scene(device=0)
    arm._test_keepalive = cubes
    return arm, collision.SceneOracle(arm)


def test_closest_pair_distance_points_normal(scene):
    arm, so = scene
    q = workloads.random_configs_np(arm, 3000, seed=21).astype(np.float32)
    q64 = q.astype(np.float64)
    rec = arm.closest_to(q)
    assert isinstance(rec, Proximities) and rec.distance.shape == (3000,) and rec.position_on_subject.shape == (3000, 3)
    pd = so.pair_distances(q64)                     # (n, n_pairs)
    d_or = pd.min(axis=1)
    ok = rec.valid & (d_or > 1e-3)                  # clearly separated surfaces
    assert ok.mean() > 0.9
    assert np.abs(rec.distance[ok] - d_or[ok]).max() < 1e-4
    # overlapping cores: signed penetration distance from the EPA (oracle: float64 EPA), every record valid
    am = pd.argmin(axis=1)
    msum = so.margins[so.pairs[am, 0]] + so.margins[so.pairs[am, 1]]
    deep = d_or <= -msum - 1e-3                      # cores interpenetrate
    assert deep.sum() > 10 and rec.valid.all()
    assert np.abs(rec.distance[deep] - d_or[deep]).max() < 2e-4
    # witness points are |distance| apart along the normal, which is a unit vector from target to subject
    ps, pt, n = rec.position_on_subject[ok], rec.position_on_target[ok], rec.normal_target_to_subject[ok]
    assert np.abs(np.linalg.norm(n, axis=1) - 1).max() < 1e-5
    assert np.abs(np.linalg.norm(ps - pt, axis=1) - rec.distance[ok]).max() < 1e-4
    assert np.abs((ps - pt) - rec.distance[ok, None] * n).max() < 1e-4
    # pair identity: the reported (subject, target) is a pair whose oracle distance is the minimum (ties within 1e-4)
    link_index = {l: i for i, l in enumerate(so.links)}
    obj_index = {o: so.n_link + i for i, o in enumerate(so.objs)}
    pair_col = {}
    for c, (a, b) in enumerate(so.pairs):
        pair_col[(int(a), int(b))] = c
        pair_col[(int(b), int(a))] = c
    idx = np.where(ok)[0]
    poses = so.link_poses(q64)
    sub_i = np.array([link_index[rec.subject[i]] for i in idx])
    tgt_i = np.array([link_index[t] if t in link_index else obj_index[t] for t in rec.target[idx]])
    cols = np.array([pair_col[(a, b)] for a, b in zip(sub_i, tgt_i)])
    assert np.abs(pd[idx, cols] - d_or[idx]).max() < 1e-4
    # each witness point lies on the surface of its own shape (point-to-shape distance ~ 0 under the same model;
    # box-box pairs are evaluated on the sharp boxes)
    sharp = so.sharp[cols].astype(bool)
    for sh in (False, True):
        for k in np.unique(sub_i[sharp == sh]):
            m = (sub_i == k) & (sharp == sh)
            assert np.abs(so.point_to_shape(k, poses[idx[m], k], ps[m].astype(np.float64), sharp=sh)).max() < 2e-4
        for k in np.unique(tgt_i[sharp == sh]):
            m = (tgt_i == k) & (sharp == sh)
            shape_pose = poses[idx[m], k] if k < so.n_link else np.repeat(so.static_poses[k - so.n_link][None], m.sum(), 0)
            assert np.abs(so.point_to_shape(k, shape_pose, pt[m].astype(np.float64), sharp=sh)).max() < 2e-4


def test_distance_gradient_matches_finite_differences_of_the_oracle(scene):
    arm, so = scene
    q = workloads.random_configs_np(arm, 1500, seed=22).astype(np.float32)
    q64 = q.astype(np.float64)
    rec = arm.closest_to(q, gradient=True)
    g = arm.jacobian_proximity_closest(q)
    assert g.shape == (1500, 7) and np.array_equal(g, rec.gradient)
    pd = np.sort(so.pair_distances(q64), axis=1)
    # differentiable points: a unique closest pair (gap to the runner-up) and surfaces clearly apart
    sel = rec.valid & (pd[:, 0] > 5e-3) & (pd[:, 1] - pd[:, 0] > 2e-2)
    assert sel.sum() > 300
    qs = q64[sel]
    h = 1e-5
    fd = np.empty((qs.shape[0], 7))
    for j in range(7):
        e = np.zeros(7); e[j] = h
        fd[:, j] = (so.min_distance(qs + e) - so.min_distance(qs - e)) / (2 * h)
    err = np.abs(g[sel] - fd)
    # contact features can switch inside the +-h stencil (vertex/edge/face transitions): judge the bulk
    assert np.median(err.max(axis=1)) < 2e-4
    assert (err.max(axis=1) < 5e-3).mean() > 0.97


def test_single_configuration_and_types(scene, cuda):
    import torch
    arm, so = scene
    q = workloads.random_configs_np(arm, 64, seed=23)
    one = arm.closest_to(q[0])
    assert isinstance(one, Proximity) and one.position_on_subject.shape == (3,) and isinstance(one.distance, float)
    batch = arm.closest_to(q)
    assert batch.distance.dtype == np.float64
    assert abs(one.distance - batch.distance[0]) < 1e-6 and one.subject is batch.subject[0] and one.target is batch.target[0]
    t = arm.closest_to(torch.as_tensor(q, dtype=torch.float32, device=cuda).reshape(8, 8, 7), gradient=True)
    assert t.distance.is_cuda and t.distance.shape == (8, 8) and t.gradient.shape == (8, 8, 7) and t.subject.shape == (8, 8) and t.subject_index.is_cuda
    assert np.abs(t.distance.cpu().numpy().reshape(-1) - batch.distance).max() < 1e-6
    # consistent with the distance query
    assert np.abs(arm.min_distance(q.astype(np.float32)) - batch.distance.astype(np.float32))[batch.valid].max() < 2e-5
    with pytest.raises(ValueError, match="q must have 7 elements"):
        arm.closest_to(np.zeros((3, 6)))


def test_pair_distances_every_allowed_pair(scene):
    """Batched Arm.collisions: the (B, P) distance matrix against the oracle's, pair by pair, and the single-configuration
    list of Proximity records (reference robots/arm.py:555-580)."""
    arm, so = scene
    q = workloads.random_configs_np(arm, 6000, seed=24).astype(np.float32)
    pairs, d = arm.pair_distances(q)
    pd = so.pair_distances(q.astype(np.float64))
    assert d.shape == pd.shape == (6000, len(pairs)) and d.dtype == np.float32
    # same pairs, possibly in another order: match columns by (subject, target) objects
    link_index = {l: i for i, l in enumerate(so.links)}
    obj_index = {o: so.n_link + i for i, o in enumerate(so.objs)}
    col = {}
    for c, (a, b) in enumerate(so.pairs):
        col[(int(a), int(b))] = c
        col[(int(b), int(a))] = c
    order = [col[(link_index[s_], link_index[t_] if t_ in link_index else obj_index[t_])] for s_, t_ in pairs]
    assert sorted(order) == list(range(len(pairs)))
    ref = pd[:, order]
    sep = ref > 1e-3
    assert sep.mean() > 0.9
    assert (np.abs(d[sep] - ref[sep]) / np.maximum(ref[sep], 1.0)).max() < 1e-4      # fp32: 1e-4 m, relative beyond 1 m
    pen = ref < -1e-3
    assert pen.sum() > 50 and np.abs(d[pen] - ref[pen]).max() < 2e-4                  # signed: EPA depth on both sides
    # the minimum over the pairs is the distance query; the arg-min is the closest pair
    md = arm.min_distance(q)
    assert np.abs(d.min(axis=1) - md)[md > 1e-3].max() < 1e-4
    # one configuration -> list of Proximity, witness points consistent with the distances
    prox = arm.collisions(q[0])
    assert len(prox) == len(pairs) and all(isinstance(p, Proximity) for p in prox)
    for p, (s_, t_), dd in zip(prox, pairs, d[0]):
        assert p.subject is s_ and p.target is t_ and abs(p.distance - dd) < 1e-6
        if dd > 1e-3:
            assert abs(np.linalg.norm(p.position_on_subject - p.position_on_target) - dd) < 1e-4
    selfp = arm.self_collisions(q[0])
    assert len(selfp) == sum(1 for _, t_ in pairs if t_ in link_index) and all(p.target in link_index for p in selfp)
    cube = so.objs[0]
    to_cube = arm.distance_to(q[0], cube)
    assert len(to_cube) == sum(1 for _, t_ in pairs if t_ is cube) and all(p.target is cube for p in to_cube)
    one = arm.distance_to(q[0], cube, link=to_cube[0].subject)
    assert len(one) == 1 and one[0].distance == to_cube[0].distance
    with pytest.raises(ValueError, match="q must be a 1D array with 7 elements"):
        arm.collisions(q[:2])
