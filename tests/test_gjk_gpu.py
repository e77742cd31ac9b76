"""The fp32 CUDA narrow phase on its own: one primitive on a 6-DoF free-flyer chain against one static
obstacle, every type combination, random relative poses, vs the float64 oracle -- through the C-ABI."""
import numpy as np
import networkx as nx
import pytest

from numbotics_b200 import workloads
from numbotics_b200.math import trans_mat
from numbotics_b200.physics import (GraphChain, Joint, Constraint, DummyLink, Link, Shape, CollisionShape,
                                    Cuboid, Sphere, Cylinder)
from numbotics_b200.robots import Arm
from oracle import collision

pytestmark = pytest.mark.gpu


def flyer(shape: CollisionShape, device=0, margin_model="bullet"):
    """base -(Px)-(Py)-(Pz)-(Rz)-(Ry)-(Rx)-> body carrying `shape`."""
    w = workloads.fresh_world("flyer")
    G = nx.DiGraph()
    names = ["base", "tx", "ty", "tz", "rz", "ry", "body"]
    for n in names[:-1]:
        G.add_node(n, link=DummyLink(1.0, np.eye(4), name=n))
    G.add_node("body", link=Link(1.0, np.eye(4), shape, name="body"))
    specs = [((1, 0, 0), Constraint.PRISMATIC), ((0, 1, 0), Constraint.PRISMATIC), ((0, 0, 1), Constraint.PRISMATIC),
             ((0, 0, 1), Constraint.REVOLUTE), ((0, 1, 0), Constraint.REVOLUTE), ((1, 0, 0), Constraint.REVOLUTE)]
    for (a, b), (ax, t) in zip(zip(names[:-1], names[1:]), specs):
        G.add_edge(a, b, joint=Joint(np.eye(4), np.array(ax, float), t, name=f"j_{b}", lower_limit=-3.0, upper_limit=3.0))
    return Arm(GraphChain(G, world_name=w.name), device=device, margin_model=margin_model), w


SHAPES = {
    "sphere": lambda: CollisionShape(Shape.SPHERE, radius=0.11),
    "box": lambda: CollisionShape(Shape.CUBOID, half_extents=np.array([0.2, 0.07, 0.12]), offset=trans_mat(pos=np.array([0.02, 0, -0.03]))),
    "cyl": lambda: CollisionShape(Shape.CYLINDER, radius=0.05, height=0.4),
    "disc": lambda: CollisionShape(Shape.CYLINDER, radius=0.2, height=0.04),
}


def make_obstacle(kind, w):
    if kind == "box":
        o = Cuboid(mass=1.0, half_extents=np.array([0.3, 0.2, 0.1]), world_name=w.name)
    elif kind == "sphere":
        o = Sphere(mass=1.0, radius=0.25, world_name=w.name)
    else:
        o = Cylinder(mass=1.0, radius=0.12, height=0.5, world_name=w.name)
    c, s = np.cos(0.6), np.sin(0.6)
    o.pose = trans_mat(pos=np.array([0.1, -0.05, 0.08]), orn=np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]]) @ np.array([[1, 0, 0], [0, 0.8, -0.6], [0, 0.6, 0.8]]))
    return o


@pytest.mark.parametrize("margin_model", ["bullet", "sharp"])
@pytest.mark.parametrize("obs_kind", ["box", "sphere", "cyl"])
@pytest.mark.parametrize("link_kind", ["sphere", "box", "cyl", "disc"])
def test_pair_distance_and_verdict(cuda, link_kind, obs_kind, margin_model):
    arm, w = flyer(SHAPES[link_kind](), margin_model=margin_model)
    obs = make_obstacle(obs_kind, w)
    import zlib
    rng = np.random.default_rng(zlib.crc32(f"{link_kind}-{obs_kind}".encode()))
    n = 12000
    q = np.zeros((n, 6))
    direction = rng.normal(size=(n, 3)); direction /= np.linalg.norm(direction, axis=1, keepdims=True)
    q[:, :3] = direction * rng.uniform(0.0, 0.9, (n, 1))          # from deep overlap to well separated
    q[:, 3:] = rng.uniform(-3.0, 3.0, (n, 3))
    q = q.astype(np.float32)
    so = collision.SceneOracle(arm, margin_model)
    d = so.min_distance(q.astype(np.float64))
    assert 0.1 < (d < 0).mean() < 0.9
    hit = arm.in_collision(q)
    clear = np.abs(d) > 1e-4
    assert np.array_equal(hit[clear], (d < 0)[clear])
    for thr in (0.03,):
        hit_t = arm.in_collision(q, thr)
        clear = np.abs(d - thr) > 1e-4
        assert np.array_equal(hit_t[clear], (d < thr)[clear])
    dist = arm.min_distance(q)
    err = np.abs(dist[d > 0] - d[d > 0])
    assert err.max() < 1e-4, (err.max(), d[d > 0][np.argmax(err)], dist[d > 0][np.argmax(err)])   # = the stated band
    neg = d < -2e-4
    assert np.all(dist[neg] <= 0), (d[neg][dist[neg] > 0][:5], dist[neg][dist[neg] > 0][:5], q[neg][dist[neg] > 0][:3])
    errn = np.abs(dist[d < 0] - d[d < 0])                      # signed: penetration depth from the EPA on both sides
    assert errn.max() < 5e-4 and np.mean(errn > 1e-4) < 0.005, (errn.max(), d[d < 0][np.argmax(errn)], dist[d < 0][np.argmax(errn)])


def test_known_answers_on_device(cuda):
    """Closed forms: sphere vs box face / edge / corner, through the free-flyer."""
    arm, w = flyer(CollisionShape(Shape.SPHERE, radius=0.1), margin_model="sharp")
    box = Cuboid(mass=1.0, half_extents=np.array([0.5, 0.4, 0.3]), world_name=w.name)
    q = np.zeros((3, 6), np.float32)
    q[0, :3] = [1.0, 0.1, 0.1]
    q[1, :3] = [1.0, 0.9, 0.0]
    q[2, :3] = [1.0, 0.9, 0.9]
    want = np.array([0.4, np.hypot(0.5, 0.5) - 0.1, np.sqrt(0.25 + 0.25 + 0.36) - 0.1])
    assert np.abs(arm.min_distance(q) - want).max() < 2e-6
