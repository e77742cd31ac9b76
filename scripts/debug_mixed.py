import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, networkx as nx
from numbotics_b200 import workloads
from numbotics_b200.math import trans_mat
from numbotics_b200.physics import GraphChain, Joint, Constraint, BasicLink, Link, Shape, CollisionShape, Cuboid, Sphere, Cylinder
from numbotics_b200.robots import Arm
from oracle import collision

def build():
    w = workloads.fresh_world("mixed")
    G = nx.DiGraph()
    G.add_node("base", link=BasicLink(1.0, np.eye(4), Shape.CUBOID, half_extents=np.array([0.1, 0.1, 0.05]), name="base"))
    G.add_node("l1", link=Link(1.0, np.eye(4), CollisionShape(Shape.CYLINDER, radius=0.04, height=0.3, offset=trans_mat(pos=np.array([0, 0, 0.15]))), name="l1"))
    G.add_node("l2", link=BasicLink(1.0, np.eye(4), Shape.SPHERE, radius=0.06, name="l2"))
    G.add_node("l3", link=Link(1.0, np.eye(4), CollisionShape(Shape.CUBOID, half_extents=np.array([0.05, 0.03, 0.12]), offset=trans_mat(pos=np.array([0, 0, 0.12]))), name="l3"))
    G.add_node("l4", link=Link(1.0, np.eye(4), CollisionShape(Shape.CYLINDER, radius=0.03, height=0.25, offset=trans_mat(pos=np.array([0.125, 0, 0]), orn=np.array([[0, 0, 1.0], [0, 1, 0], [-1, 0, 0]]))), name="l4"))
    G.add_node("l5", link=BasicLink(1.0, np.eye(4), Shape.SPHERE, radius=0.05, name="l5"))
    R = lambda ax, **kw: Joint(kw.pop("off"), np.array(ax, float), kw.pop("t", Constraint.REVOLUTE), **kw)
    G.add_edge("base", "l1", joint=R([0, 0, 1], off=trans_mat(pos=np.array([0, 0, 0.05])), lower_limit=-3, upper_limit=3))
    G.add_edge("l1", "l2", joint=R([0, 1, 0], off=trans_mat(pos=np.array([0, 0, 0.3])), lower_limit=-2, upper_limit=2))
    G.add_edge("l2", "l3", joint=R([0, 0, 1], off=trans_mat(pos=np.array([0, 0, 0.0])), t=Constraint.PRISMATIC, lower_limit=0.0, upper_limit=0.3))
    G.add_edge("l2", "l4", joint=R([1, 0, 0], off=trans_mat(pos=np.array([0.05, 0, 0])), lower_limit=-2, upper_limit=2))
    G.add_edge("l4", "l5", joint=R([0, 1, 0], off=trans_mat(pos=np.array([0.3, 0, 0])), lower_limit=-2, upper_limit=2))
    arm = Arm(GraphChain(G, world_name=w.name), device=0)
    obs = [Cuboid(mass=1.0, half_extents=np.array([0.2, 0.3, 0.1]), world_name=w.name), Sphere(mass=1.0, radius=0.15, world_name=w.name), Cylinder(mass=1.0, radius=0.1, height=0.5, world_name=w.name)]
    obs[0].pose = trans_mat(pos=np.array([0.45, 0.0, 0.4]), orn=np.array([[0.8, -0.6, 0], [0.6, 0.8, 0], [0, 0, 1.0]]))
    obs[1].position = np.array([-0.3, 0.3, 0.5])
    obs[2].pose = trans_mat(pos=np.array([0.0, -0.45, 0.45]), orn=np.array([[1.0, 0, 0], [0, 0, -1], [0, 1, 0]]))
    return arm, obs

arm, obs = build()
q = workloads.random_configs_np(arm, 5000, seed=3).astype(np.float32)
so = collision.SceneOracle(arm)
pd = so.pair_distances(q.astype(np.float64))
d = pd.min(axis=1)
hit = arm.in_collision(q); dist = arm.min_distance(q)
bad = np.where((np.abs(d) > 1e-4) & (hit != (d < 0)))[0]
print("n bad verdicts", len(bad))
names = [n for n in so.link_nodes] + [o._name for o in obs]
for i in bad[:10]:
    k = np.argmin(pd[i]); a, b = so.pairs[k]
    print(i, "oracle d", d[i], "pair", names[a], names[b], "types", so.types[a], so.types[b], "gpu hit", hit[i], "gpu dist", dist[i], "q", q[i])
err = np.abs(dist - d); err[d <= 0] = 0
order = np.argsort(-err)[:10]
print("worst distance errors")
for i in order:
    k = np.argmin(pd[i]); a, b = so.pairs[k]
    print(i, "oracle", d[i], "gpu", dist[i], "pair", names[a], names[b], so.types[a], so.types[b])
# kinova
arm2, cubes = workloads.prm_scene(device=0)
so2 = collision.SceneOracle(arm2)
q2 = workloads.random_configs_np(arm2, 3000, seed=5).astype(np.float32)
pd2 = so2.pair_distances(q2.astype(np.float64)); d2 = pd2.min(axis=1)
dist2 = arm2.min_distance(q2)
err = np.abs(dist2 - d2); err[d2 <= 0] = 0
names2 = [n for n in so2.link_nodes] + ["cube"]*3
print("kinova worst distance errors")
for i in np.argsort(-err)[:10]:
    k = np.argmin(pd2[i]); a, b = so2.pairs[k]
    print(i, "oracle", d2[i], "gpu", dist2[i], "pair", names2[a], names2[min(b, len(names2)-1)], so2.types[a], so2.types[b])
