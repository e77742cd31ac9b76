"""Signed distances below contact (penetration depth from the EPA, csrc/nbx_gjk.cuh) against the float64 oracle's EPA
(pinned by tests/test_epa_oracle.py): every allowed pair of contact-rich configurations, witness points and push-out
normal of penetrating pairs, the distance gradient of a penetrating closest pair, single-link scenes that drive one
primitive of each kind through an obstacle of each kind (including the exactly symmetric poses that degenerate the GJK
seed), and negative thresholds in edge validation / bisection."""
import numpy as np
import networkx as nx
import pytest

from numbotics_b200 import workloads
from numbotics_b200.math import trans_mat
from numbotics_b200.physics import GraphChain, Joint, Constraint, BasicLink, DummyLink, Shape, Cube, Cuboid, Sphere, Cylinder
from numbotics_b200.robots import Arm
from oracle import collision

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scene(cuda):
    arm, cubes = workloads.prm_scene(device=0)
    arm._test_keepalive = cubes
    return arm, collision.SceneOracle(arm)


def _dense(n, seed):
    rng = np.random.default_rng(seed)
    sgn = lambda k: rng.choice([-1.0, 1.0], k)
    q = rng.normal(scale=0.5, size=(n, 7))
    q[:, 1] = rng.uniform(1.2, 2.41, n) * sgn(n)
    q[:, 3] = rng.uniform(2.2, 2.66, n) * sgn(n)
    q[:, 5] = rng.uniform(1.5, 2.23, n) * sgn(n)
    return q.astype(np.float32)


def _match_columns(so, pairs):
    link_index = {l: i for i, l in enumerate(so.links)}
    obj_index = {o: so.n_link + i for i, o in enumerate(so.objs)}
    col = {}
    for c, (a, b) in enumerate(so.pairs):
        col[(int(a), int(b))] = c
        col[(int(b), int(a))] = c
    order = [col[(link_index[s_], link_index[t_] if t_ in link_index else obj_index[t_])] for s_, t_ in pairs]
    return order, link_index, obj_index


def test_every_pair_signed_distance_and_witnesses(scene):
    arm, so = scene
    q = _dense(4000, seed=31)
    pairs, rec = arm.pair_distances(q, points=True)
    order, link_index, obj_index = _match_columns(so, pairs)
    ref = so.pair_distances(q.astype(np.float64))[:, order]
    d = rec["distance"]
    pen = ref < -1e-3
    assert pen.sum() > 2000 and rec["valid"].all()
    err = np.abs(d - ref)
    assert err[pen].max() < 2e-4, (err[pen].max(), d[pen][err[pen].argmax()], ref[pen][err[pen].argmax()])
    far = ref > 1e-3
    rel = err[far] / np.maximum(np.abs(ref[far]), 1.0)
    assert rel.max() < 3e-4 and np.mean(rel > 1e-4) < 2e-5      # contact-rich configurations: 1 pair in 10^5 stops a little early
    assert err[~pen & ~far].max() < 5e-4 and np.mean(err[~pen & ~far] > 1e-4) < 0.01      # |d| < 1 mm: see tests/test_gjk_host_harness.py
    # penetrating pairs: unit normal, witness points |d| apart along it (subject behind target: ps - pt = d * n, d < 0)
    ps, pt, n = rec["position_on_subject"][pen], rec["position_on_target"][pen], rec["normal_target_to_subject"][pen]
    assert np.abs(np.linalg.norm(n, axis=1) - 1).max() < 1e-4
    assert np.abs((ps - pt) - d[pen][:, None] * n).max() < 1e-4
    # ... and each witness point lies on the surface of its own shape -- or, on a curved feature, up to the sagitta of the
    # polytope face inside it: the EPA's witnesses are convex combinations of support points (points OF the two shapes whose
    # difference is the minimum translation), never outside
    poses = so.link_poses(q.astype(np.float64))
    bi, pi = np.where(pen)
    for col in np.unique(pi)[:12]:
        m = pi == col
        s_, t_ = pairs[col]
        k = link_index[s_]
        dsub = so.point_to_shape(k, poses[bi[m], k], ps[m].astype(np.float64))
        assert dsub.max() < 3e-4 and dsub.min() > -5e-3 and np.median(np.abs(dsub)) < 1e-4
        kt = link_index[t_] if t_ in link_index else obj_index[t_]
        tp = poses[bi[m], kt] if kt < so.n_link else np.repeat(so.static_poses[kt - so.n_link][None], m.sum(), 0)
        dtgt = so.point_to_shape(kt, tp, pt[m].astype(np.float64))
        assert dtgt.max() < 3e-4 and dtgt.min() > -5e-3 and np.median(np.abs(dtgt)) < 1e-4


def test_gradient_of_a_penetrating_closest_pair(scene):
    arm, so = scene
    q = _dense(3000, seed=32)
    q64 = q.astype(np.float64)
    rec = arm.closest_to(q, gradient=True)
    pd = np.sort(so.pair_distances(q64), axis=1)
    sel = (pd[:, 0] < -1e-2) & (pd[:, 1] - pd[:, 0] > 2e-2)        # clearly penetrating, unique deepest pair
    assert sel.sum() > 100
    qs = q64[sel][:300]
    h = 1e-5
    fd = np.empty((qs.shape[0], 7))
    for j in range(7):
        e = np.zeros(7); e[j] = h
        fd[:, j] = (so.min_distance(qs + e) - so.min_distance(qs - e)) / (2 * h)
    err = np.abs(rec.gradient[sel][:300] - fd).max(axis=1)
    assert np.median(err) < 1e-3 and (err < 2e-2).mean() > 0.9


def _one_link_arm(world, shape, **kw):
    """A slider along world x carrying one primitive: q[0] = x position of its centre."""
    G = nx.DiGraph()
    G.add_node("base", link=DummyLink(1.0, np.eye(4), name="base"))
    G.add_node("body", link=BasicLink(1.0, np.eye(4), shape, name="body", **kw))
    G.add_edge("base", "body", joint=Joint(np.eye(4), np.array([1.0, 0.0, 0.0]), Constraint.PRISMATIC, name="slide",
                                           lower_limit=-5.0, upper_limit=5.0))
    return Arm(GraphChain(G, world_name=world.name), device=0)


@pytest.mark.parametrize("link_kind", ["sphere", "cylinder", "box"])
@pytest.mark.parametrize("obs_kind", ["cube", "cuboid", "sphere", "cylinder"])
def test_each_primitive_through_each_obstacle(cuda, link_kind, obs_kind):
    w = workloads.fresh_world("pen")
    kw = {"sphere": dict(shape=Shape.SPHERE, radius=0.07), "cylinder": dict(shape=Shape.CYLINDER, radius=0.05, height=0.3),
          "box": dict(shape=Shape.CUBOID, half_extents=np.array([0.06, 0.09, 0.12]))}[link_kind]
    arm = _one_link_arm(w, **kw)
    if obs_kind == "cube":
        o = Cube(half_extent=0.3, mass=1.0, world_name=w.name)
    elif obs_kind == "cuboid":
        o = Cuboid(half_extents=np.array([0.3, 0.2, 0.25]), mass=1.0, world_name=w.name)
    elif obs_kind == "sphere":
        o = Sphere(radius=0.3, mass=1.0, world_name=w.name)
    else:
        o = Cylinder(radius=0.25, height=0.5, mass=1.0, world_name=w.name)
    o.position = np.array([0.0, 0.013, -0.021])          # slightly off the slider's line ...
    so = collision.SceneOracle(arm)
    x = np.linspace(-0.6, 0.6, 241).astype(np.float32)[:, None]
    d = arm.min_distance(x)
    ref = so.min_distance(x.astype(np.float64))
    assert (ref < -0.05).sum() > 50
    assert np.abs(d - ref).max() < 3e-4, (np.abs(d - ref).max(), x[np.abs(d - ref).argmax()])
    o.position = np.zeros(3)                                # ... and exactly on it: symmetric, degenerate GJK seeds
    so = collision.SceneOracle(arm)
    d = arm.min_distance(x)
    ref = so.min_distance(x.astype(np.float64))
    assert np.abs(d - ref).max() < 4e-4, (np.abs(d - ref).max(), x[np.abs(d - ref).argmax()])
    del o


def test_negative_threshold_edges_and_bisection(scene):
    arm, so = scene
    q = _dense(1200, seed=33)
    s, g = q[:600], q[600:]
    g = (s + 0.15 * (g - s)).astype(np.float32)
    for thr in (-0.003, -0.03):
        valid = arm.edges_valid(s, g, 0.1, threshold=thr)
        ov, _, md = so.edges_valid(s.astype(np.float64), g.astype(np.float64), 0.1, threshold=thr, return_detail=True)
        clear = np.abs(md - thr) > 2e-4
        assert clear.mean() > 0.9 and (~ov).sum() > 20
        assert np.array_equal(valid[clear], ov[clear]), thr
    # bisection towards the surface "5 mm inside": the far end keeps d < -5 mm, the near end does not
    d = so.min_distance(q.astype(np.float64))
    far = q[d < -0.02][:200]
    centre = np.zeros(7, np.float32)
    assert so.min_distance(centre[None].astype(np.float64))[0] > 0.01
    out = arm.bisect_to_collision(centre, far, 12, -0.005)
    d_out = so.min_distance(out.astype(np.float64))
    near = out - (far - centre) / 2 ** 12
    d_near = so.min_distance(near.astype(np.float64))
    assert (d_out < -0.005 + 2e-4).all() and (d_near > -0.005 - 2e-4).all()


def test_epa_queue_overflow_of_the_distance_query_changes_nothing(scene):
    """The distance query hands overlapping candidates to `epa_kernel` through a global queue of B / 4 entries (at least 4096)
    and evaluates in place what the queue cannot take.  On contact-rich configurations more than a quarter of the batch
    overlaps: one call on the whole batch (queue overflows) must give, bit for bit, the distances and closest pairs of calls
    on chunks small enough for the queue; and both agree with the float64 oracle."""
    arm, so = scene
    q = _dense(40000, seed=77)
    d_all = arm.min_distance(q)
    assert (d_all < -1e-3).mean() > 0.3                     # > 12 000 configurations with an overlapping pair against 10 000 slots
    d_chunks = np.concatenate([arm.min_distance(q[i:i + 4000]) for i in range(0, len(q), 4000)])
    assert np.array_equal(d_all, d_chunks)
    c_all = arm.closest_to(q[:12000])
    c_chunks = [arm.closest_to(q[i:i + 3000]) for i in range(0, 12000, 3000)]
    assert np.array_equal(c_all.distance, np.concatenate([c.distance for c in c_chunks]))
    assert np.array_equal(c_all.position_on_subject, np.concatenate([c.position_on_subject for c in c_chunks]))
    ref = so.min_distance(q[:3000].astype(np.float64))
    err = np.abs(d_all[:3000] - ref)
    pen = ref < -1e-3
    assert pen.sum() > 800 and err[pen].max() < 2e-4 and err[~pen].max() < 5e-4
