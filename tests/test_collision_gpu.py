"""CUDA Arm.in_collision / edge validation vs the float64 oracle (checker route), through the C-ABI."""
import numpy as np
import pytest

from numbotics_b200 import workloads
from numbotics_b200.physics import Cube, Cuboid, Sphere, Cylinder
from numbotics_b200.planning.sampling_based import ConnectorParams, DiscreteConnector, ArmValidityChecker
from oracle import collision

pytestmark = pytest.mark.gpu

EPS = 1e-4   # stated contact band (metres): verdicts may differ only where |oracle distance - threshold| <= EPS


@pytest.fixture(scope="module")
def scene(cuda):
    arm, cubes = workloads.prm_scene(device=0)
    return arm, cubes, collision.SceneOracle(arm)


def check_verdicts(hit, d, thr=0.0, max_excluded=0.005):
    """Verdicts equal the oracle's outside the stated contact band; the band may exclude at most `max_excluded` of the
    samples (measured: ~1e-4 of uniformly random configurations, ~2e-3 of the contact-rich ones)."""
    clear = np.abs(d - thr) > EPS
    excluded = 1.0 - clear.mean()
    assert excluded <= max_excluded, f"the contact band excludes {excluded:.2e} of the samples"
    assert np.array_equal(hit[clear], (d < thr)[clear])
    return excluded


@pytest.mark.parametrize("margin_model", ["bullet", "sharp"])
def test_in_collision_matches_oracle(cuda, margin_model):
    arm, cubes = workloads.prm_scene(device=0, margin_model=margin_model)
    so = collision.SceneOracle(arm, margin_model)
    q = workloads.random_configs_np(arm, 6000 + 13, seed=1)
    hit = arm.in_collision(q.astype(np.float32))
    d = so.min_distance(q.astype(np.float32).astype(np.float64))
    assert hit.dtype == bool and hit.shape == (q.shape[0],)
    assert 50 < hit.sum() < q.shape[0] // 2
    check_verdicts(hit, d)


def _dense_contact_configs(n, seed):
    rng = np.random.default_rng(seed)
    sgn = lambda k: rng.choice([-1.0, 1.0], k)
    q = rng.normal(scale=0.5, size=(n, 7))
    q[:, 1] = rng.uniform(1.2, 2.41, n) * sgn(n)     # shoulder down
    q[:, 3] = rng.uniform(2.2, 2.66, n) * sgn(n)     # elbow folded
    q[:, 5] = rng.uniform(1.5, 2.23, n) * sgn(n)     # wrist folded
    return q.astype(np.float32)


def test_in_collision_dense_contacts(scene):
    """Configurations concentrated where contacts happen: arm folded on itself / reaching into the cubes."""
    arm, cubes, so = scene
    rng = np.random.default_rng(7)
    sgn = lambda n: rng.choice([-1.0, 1.0], n)
    q = rng.normal(scale=0.5, size=(4000, 7))
    q[:, 1] = rng.uniform(1.2, 2.41, 4000) * sgn(4000)     # shoulder down
    q[:, 3] = rng.uniform(2.2, 2.66, 4000) * sgn(4000)     # elbow folded
    q[:, 5] = rng.uniform(1.5, 2.23, 4000) * sgn(4000)     # wrist folded
    q = q.astype(np.float32)
    hit = arm.in_collision(q)
    d = so.min_distance(q.astype(np.float64))
    assert 0.15 < hit.mean() < 0.95
    check_verdicts(hit, d)


def test_min_distance_matches_oracle(scene):
    arm, cubes, so = scene
    q = workloads.random_configs_np(arm, 3000, seed=5).astype(np.float32)
    dist = arm.min_distance(q)
    d = so.min_distance(q.astype(np.float64))
    sep = d > 0
    err = np.abs(dist[sep] - d[sep])
    assert err.max() < 1e-4, (err.max(), np.argmax(err), dist[sep][np.argmax(err)], d[sep][np.argmax(err)])
    # penetration: signed distance from the EPA on both sides (float32 kernel / float64 oracle)
    assert (~sep).sum() > 30
    errn = np.abs(dist[~sep] - d[~sep])
    assert errn.max() < 2e-4, (errn.max(), dist[~sep][np.argmax(errn)], d[~sep][np.argmax(errn)])


def test_threshold_and_single_config_api(scene):
    arm, cubes, so = scene
    q = workloads.random_configs_np(arm, 2000, seed=6).astype(np.float32)
    d = so.min_distance(q.astype(np.float64))
    for thr in (0.02, 0.1):
        check_verdicts(arm.in_collision(q, thr), d, thr)
    one = arm.in_collision(q[0].astype(np.float64))
    assert isinstance(one, bool) and one == bool(d[0] < 0)
    q3 = q[:24].reshape(2, 3, 4, 7)
    assert arm.in_collision(q3).shape == (2, 3, 4)
    with pytest.raises(ValueError, match="q must have 7 elements"):
        arm.in_collision(np.zeros(6))
    # negative thresholds (reference arm.py:603-604 compares Bullet's signed distance): inside the margins, and beyond them
    # where only interpenetrating cores count
    qd = np.concatenate([q, _dense_contact_configs(3000, seed=9)])
    dd = so.min_distance(qd.astype(np.float64))
    for thr in (-0.003, -0.02, -0.06):
        hit = arm.in_collision(qd, thr)
        clear = np.abs(dd - thr) > 2 * EPS
        assert (dd < thr).sum() > 20
        assert np.array_equal(hit[clear], (dd < thr)[clear]), thr


def test_torch_in_torch_out(scene, cuda):
    import torch
    arm, cubes, so = scene
    q = workloads.random_configs_cuda(arm, 5000, seed=2)
    hit = arm.in_collision(q)
    assert hit.is_cuda and hit.dtype == torch.bool and hit.shape == (5000,)
    d = so.min_distance(q.double().cpu().numpy())
    check_verdicts(hit.cpu().numpy(), d)


def test_moving_an_obstacle_refreshes_the_scene(cuda):
    arm, cubes = workloads.prm_scene(device=0)
    q = np.zeros(7, np.float32)            # arm straight up
    assert arm.in_collision(q) is False
    cubes[2].position = np.array([0.0, 0.0, 1.0])   # drop a cube on the arm
    assert arm.in_collision(q) is True
    arm.remove_collision_pair("forearm_link", cubes[2].name)
    so = collision.SceneOracle(arm)
    qs = workloads.random_configs_np(arm, 1500, seed=8).astype(np.float32)
    check_verdicts(arm.in_collision(qs), so.min_distance(qs.astype(np.float64)))


def test_mixed_primitives_and_tree_branches(cuda):
    """Spheres, boxes and cylinders on a branching tree with a prismatic joint, obstacles of all three kinds."""
    import networkx as nx
    from numbotics_b200.math import trans_mat
    from numbotics_b200.physics import GraphChain, Joint, Constraint, BasicLink, Link, Shape, CollisionShape
    from numbotics_b200.robots import Arm
    w = workloads.fresh_world("mixed")
    G = nx.DiGraph()
    G.add_node("base", link=BasicLink(1.0, np.eye(4), Shape.CUBOID, half_extents=np.array([0.1, 0.1, 0.05]), name="base"))
    G.add_node("l1", link=Link(1.0, np.eye(4), CollisionShape(Shape.CYLINDER, radius=0.04, height=0.3,
                                                              offset=trans_mat(pos=np.array([0, 0, 0.15]))), name="l1"))
    G.add_node("l2", link=BasicLink(1.0, np.eye(4), Shape.SPHERE, radius=0.06, name="l2"))
    G.add_node("l3", link=Link(1.0, np.eye(4), CollisionShape(Shape.CUBOID, half_extents=np.array([0.05, 0.03, 0.12]),
                                                              offset=trans_mat(pos=np.array([0, 0, 0.12]))), name="l3"))
    G.add_node("l4", link=Link(1.0, np.eye(4), CollisionShape(
        Shape.CYLINDER, radius=0.03, height=0.25,
        offset=trans_mat(pos=np.array([0.125, 0, 0]), orn=np.array([[0, 0, 1.0], [0, 1, 0], [-1, 0, 0]]))), name="l4"))
    G.add_node("l5", link=BasicLink(1.0, np.eye(4), Shape.SPHERE, radius=0.05, name="l5"))
    R = lambda ax, **kw: Joint(kw.pop("off"), np.array(ax, float), kw.pop("t", Constraint.REVOLUTE), **kw)
    G.add_edge("base", "l1", joint=R([0, 0, 1], off=trans_mat(pos=np.array([0, 0, 0.05])), lower_limit=-3, upper_limit=3))
    G.add_edge("l1", "l2", joint=R([0, 1, 0], off=trans_mat(pos=np.array([0, 0, 0.3])), lower_limit=-2, upper_limit=2))
    G.add_edge("l2", "l3", joint=R([0, 0, 1], off=trans_mat(pos=np.array([0, 0, 0.0])), t=Constraint.PRISMATIC, lower_limit=0.0, upper_limit=0.3))
    G.add_edge("l2", "l4", joint=R([1, 0, 0], off=trans_mat(pos=np.array([0.05, 0, 0])), lower_limit=-2, upper_limit=2))   # branch
    G.add_edge("l4", "l5", joint=R([0, 1, 0], off=trans_mat(pos=np.array([0.3, 0, 0])), lower_limit=-2, upper_limit=2))
    arm = Arm(GraphChain(G, world_name=w.name), device=0)
    obs = [Cuboid(mass=1.0, half_extents=np.array([0.2, 0.3, 0.1]), world_name=w.name),
           Sphere(mass=1.0, radius=0.15, world_name=w.name), Cylinder(mass=1.0, radius=0.1, height=0.5, world_name=w.name)]
    obs[0].pose = trans_mat(pos=np.array([0.45, 0.0, 0.4]), orn=np.array([[0.8, -0.6, 0], [0.6, 0.8, 0], [0, 0, 1.0]]))
    obs[1].position = np.array([-0.3, 0.3, 0.5])
    obs[2].pose = trans_mat(pos=np.array([0.0, -0.45, 0.45]), orn=np.array([[1.0, 0, 0], [0, 0, -1], [0, 1, 0]]))
    assert arm.dof == 5
    q = workloads.random_configs_np(arm, 5000, seed=3).astype(np.float32)
    so = collision.SceneOracle(arm)
    d = so.min_distance(q.astype(np.float64))
    hit = arm.in_collision(q)
    assert 0.05 < hit.mean() < 0.95
    check_verdicts(hit, d, max_excluded=0.01)       # a scene built so that most samples touch something: 8e-3 are inside the band
    dist = arm.min_distance(q)
    assert np.abs(dist[d > 0] - d[d > 0]).max() < 1e-4


def test_no_pairs_raises_like_min_of_empty(cuda):
    arm = workloads.serial_arm(4, device=0)     # no geometry, no obstacles
    with pytest.raises(ValueError, match="empty sequence"):
        arm.in_collision(np.zeros(4))


# ------------------------------------------------------------------------------------------------ edges
def check_edges(valid, ov, md, thr=0.0):
    clear = np.abs(md - thr) > EPS
    assert np.array_equal(valid[clear], ov[clear])


def test_edges_match_oracle(scene):
    arm, cubes, so = scene
    q = workloads.random_configs_np(arm, 700, seed=12).astype(np.float32)
    s, g = q[:350], q[350:]
    # shorten half of the edges so both valid and invalid ones occur
    g[:200] = s[:200] + 0.15 * (g[:200] - s[:200])
    valid, ns = arm.edges_valid(s, g, 0.1, return_samples=True)
    ov, ons, md = so.edges_valid(s.astype(np.float64), g.astype(np.float64), 0.1, return_detail=True)
    assert np.array_equal(ns, ons)                      # numpy.arange sample count reproduced exactly
    assert (~valid).sum() >= 10 and valid.sum() >= 10
    check_edges(valid, ov, md)
    # steer semantics
    valid2, ns2 = arm.edges_valid(s, g, 0.1, max_distance=1.0, return_samples=True)
    ov2, ons2, md2 = so.edges_valid(s.astype(np.float64), g.astype(np.float64), 0.1, max_distance=1.0, return_detail=True)
    assert np.array_equal(ns2, ons2)
    check_edges(valid2, ov2, md2)
    assert valid2.sum() >= valid.sum()


def test_edge_corner_cases(scene):
    arm, cubes, so = scene
    free = np.zeros(7, np.float32)
    assert arm.in_collision(free) is False
    s = np.stack([free, free, free, free])
    g = s.copy()
    g[1, 0] = 1e-8          # distance <= float32 eps -> None
    g[2, 0] = 0.5           # distance/resolution = 5 (arange boundary)
    g[3, 0] = 0.05          # shorter than the resolution: samples {0, 1}
    valid, ns = arm.edges_valid(s, g, 0.1, return_samples=True)
    assert list(valid) == [False, False, True, True]
    ov, ons, _ = so.edges_valid(s.astype(np.float64), g.astype(np.float64), 0.1, return_detail=True)
    assert list(ns) == list(ons) and list(ns[:2]) == [0, 0] and ns[3] == 2
    assert arm.edges_valid(np.zeros((0, 7), np.float32), np.zeros((0, 7), np.float32), 0.1).shape == (0,)
    with pytest.raises(ValueError, match="same shape"):
        arm.edges_valid(s, g[:2], 0.1)


def test_discrete_connector_surface(scene, cuda):
    arm, cubes, so = scene
    chk = ArmValidityChecker(arm)
    conn = DiscreteConnector(ConnectorParams(resolution=0.1, max_distance=np.pi, validity_checker=chk))
    q = workloads.random_configs_np(arm, 80, seed=21)
    free = q[~so.in_collision(q)]
    a, b = free[0], free[0] + 0.2 * (free[1] - free[0])
    out = conn.connect(a, b)
    ov = so.edges_valid(a[None].astype(np.float32), b[None].astype(np.float32), 0.1)[0]
    assert (out is not None) == bool(ov)
    if out is not None:
        assert np.array_equal(out, b) and out is not b
    assert conn.connect(a, a) is None
    assert conn.is_valid(a) is True
    # a colliding goal can never connect
    bad = q[so.in_collision(q)][0] if so.in_collision(q).any() else None
    if bad is not None:
        assert conn.connect(a, bad) is None
    # steer returns the clipped state
    far = free[2]
    st = conn.steer(a, far)
    dist = np.linalg.norm(far - a)
    if st is not None:
        t_f = min(1.0, np.pi / dist)
        assert np.allclose(st, (1 - t_f) * a + t_f * far)
    # custom distance function -> host sample grid + batched validity, same verdict as the oracle
    out2 = conn.connect(a, b, distance_func=lambda x, y: float(np.linalg.norm(x - y)))
    assert (out2 is not None) == (out is not None)
    # batched surface
    s, g = free[:20], free[20:40]
    vb = conn.connect_batch(s, g)
    ovb, _, md = so.edges_valid(s.astype(np.float32), g.astype(np.float32), 0.1, return_detail=True)
    check_edges(vb, ovb, md)
    vs, ends = conn.steer_batch(s, g)
    assert ends.shape == s.shape


def test_static_pruning_never_changes_a_verdict(scene):
    """Obstacle pairs that the reach bound proves can never come within the threshold are dropped from the
    verdict scene; the unpruned scene (used by min_distance) must give the same answers."""
    arm, cubes, so = scene
    pruned = arm.scene_arrays(0.05)
    assert pruned["env_pairs"].shape[0] < pruned["n_env_pairs_before_pruning"] == 30
    q = workloads.random_configs_np(arm, 20000, seed=31).astype(np.float32)
    for thr in (0.0, 0.04, 0.2):                     # 0.2 forces a rebuild with a larger static threshold
        hit = arm.in_collision(q, thr)
        dist = arm.min_distance(q)
        clear = np.abs(dist - thr) > 1e-4
        assert np.array_equal(hit[clear], (dist < thr)[clear])
    # the reach bound itself, against sampled forward kinematics
    from numbotics_b200.robots import descriptor
    reach = descriptor.primitive_reach(arm._arrays, arm.joint_limits)
    base = arm.base_pose[:3, 3]
    for p, (node, off) in enumerate(zip(so.link_nodes, so.link_offsets)):
        centers = (so.fk.forward_kinematics(q[:3000].astype(np.float64), node) @ off)[:, :3, 3]
        link = arm._links_from_nodes[node]
        assert np.linalg.norm(centers - base, axis=1).max() <= reach[arm._arrays["prim_of_link"][link]] + 1e-9


def test_host_buffer_entry_points_and_concurrent_threads(scene):
    """`nbx_in_collision_host_f32` / `nbx_edges_valid_host_f32` (HOST pointers in, HOST verdicts out) give the bits of the
    device path, also when several host threads call at once on the same immutable handles (the IRIS thread-pool caller,
    reference utils/threading.py:33-37)."""
    import threading
    from numbotics_b200 import _cabi
    arm, cubes, so = scene
    lib = _cabi.load()
    q = workloads.random_configs_np(arm, 20000 + 7, seed=31).astype(np.float32)
    expect = arm.in_collision(q)
    sc = arm._scene(0.0)
    keep, base = _cabi.base_pose_arg(arm.base_pose)
    h = arm._handle()

    def host_call(rows):
        out = np.full(rows.shape[0], 7, np.uint8)
        _cabi.check(lib.nbx_in_collision_host_f32(h.ptr, sc.ptr, rows.ctypes.data, rows.shape[0], base, 0.0, out.ctypes.data))
        return out.astype(bool)

    assert np.array_equal(host_call(q), expect)
    s, g = np.ascontiguousarray(q[:5000]), np.ascontiguousarray(q[5000:10000])
    ev = np.full(5000, 7, np.uint8)
    _cabi.check(lib.nbx_edges_valid_host_f32(h.ptr, sc.ptr, s.ctypes.data, g.ctypes.data, 5000, base, 0.1, 0.0, 0.0, ev.ctypes.data))
    assert np.array_equal(ev.astype(bool), arm.edges_valid(s, g, 0.1))
    assert lib.nbx_in_collision_host_f32(h.ptr, sc.ptr, None, 4, base, 0.0, ev.ctypes.data) == -1      # NULL q: status, not a crash
    # 8 threads, each its own slice, repeated
    results = [None] * 8
    def worker(i):
        rows = np.ascontiguousarray(q[i::8])
        ok = True
        for _ in range(5):
            ok = ok and np.array_equal(host_call(rows), expect[i::8])
        results[i] = ok
    threads = [threading.Thread(target=worker, args=(i,)) for i in range(8)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert all(results)


def test_crowded_scene_overflow_paths(cuda):
    """A workspace crowded with obstacles: dozens of candidate pairs per configuration, so a tile of 32 configurations
    overflows the warp's pending list (-> exhaustive evaluation of the tile) and the batch overflows the global candidate
    queue (-> in-place evaluation).  Those are the cold paths of csrc/nbx_collide.cu; verdicts, distances and edge verdicts
    must still equal the oracle's."""
    w = workloads.fresh_world("crowded")
    arm = workloads.kinova_arm(w, device=0)
    workloads.apply_rrt_pair_removals(arm)
    rng = np.random.default_rng(11)
    obs = []
    for i in range(36):                          # small cubes scattered through the reachable volume
        c = Cube(half_extent=float(rng.uniform(0.04, 0.09)), mass=1.0, world_name=w.name)
        r, th, z = rng.uniform(0.25, 0.8), rng.uniform(-np.pi, np.pi), rng.uniform(0.1, 1.0)
        c.position = np.array([r * np.cos(th), r * np.sin(th), z])
        obs.append(c)
    so = collision.SceneOracle(arm)
    q = workloads.random_configs_np(arm, 3000, seed=12).astype(np.float32)
    d = so.min_distance(q.astype(np.float64))
    hit = arm.in_collision(q)
    assert hit.mean() > 0.5                      # nearly everything touches something
    check_verdicts(hit, d)
    # a threshold that makes (almost) every pair a candidate
    hit2 = arm.in_collision(q, 0.3)
    clear = np.abs(d - 0.3) > EPS
    assert np.array_equal(hit2[clear], (d < 0.3)[clear]) and hit2.mean() > 0.95
    dist = arm.min_distance(q[:1000])
    sep = d[:1000] > 0
    assert np.abs(dist[sep] - d[:1000][sep]).max() < 1e-4
    assert np.abs(dist[~sep] - d[:1000][~sep]).max() < 3e-4       # penetration depth into the small cubes (EPA both sides)
    # edges through the clutter
    s, g = q[:300], q[300:600].copy()
    g[:150] = s[:150] + 0.1 * (g[:150] - s[:150])
    valid = arm.edges_valid(s, g, 0.1)
    ov, _, md = so.edges_valid(s.astype(np.float64), g.astype(np.float64), 0.1, return_detail=True)
    check_edges(valid, ov, md)
    del obs


def test_multi_tile_ranges_partial_tiles_and_unaligned_q(scene, cuda):
    """Ranges of several tiles per warp (the next q tile is requested with cp.async while the current one is culled),
    a partial last tile, and a q buffer that is only 4-byte aligned (the 4-byte cp.async path): same verdicts and
    distances as single-tile calls on the same rows, and as the oracle on a sub-sample."""
    import torch
    arm, cubes, so = scene
    B = 1_250_000 + 19                      # > 2 tiles per warp with the default range length, ragged end
    q = workloads.random_configs_cuda(arm, B, seed=11)
    hit = arm.in_collision(q)
    dist = arm.min_distance(q[:200_003])
    # (a) chunks small enough for one tile per warp
    ref = torch.cat([arm.in_collision(q[i:i + 4096]) for i in range(0, B, 4096)])
    assert torch.equal(hit, ref)
    # (b) only 4-byte aligned
    store = torch.empty(B * arm.dof + 1, dtype=torch.float32, device=q.device)
    qu = store[1:].view(B, arm.dof)
    qu.copy_(q)
    assert qu.data_ptr() % 16 == 4
    assert torch.equal(arm.in_collision(qu), hit)
    assert torch.equal(arm.min_distance(qu[:200_003]), dist)
    # (c) oracle on a strided sub-sample
    idx = torch.arange(0, B, 97, device=q.device)
    d = so.min_distance(q[idx].cpu().numpy().astype(np.float64))
    check_verdicts(hit[idx].cpu().numpy(), d)
