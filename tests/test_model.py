"""Host-side mirror of the reference's model surface: URDF semantics, q ordering, the per-frame cache,
pair-set rules and their error behaviour (no GPU needed)."""
import numpy as np
import networkx as nx
import pytest

from numbotics_b200 import workloads
from numbotics_b200.math import trans_mat, rpy_matrix
from numbotics_b200.physics import (World, GraphChain, Cube, Cuboid, Sphere, Joint, Constraint, BasicLink, DummyLink,
                                    Shape, CollisionShape)
from numbotics_b200.physics.urdf import graph_from_urdf
from numbotics_b200.robots import Arm
from numbotics_b200.robots import descriptor

URDF_SMALL = """<?xml version="1.0"?>
<robot name="t">
  <link name="a"><collision><origin xyz="0 0 0.1" rpy="0 0 0"/><geometry><box size="0.2 0.4 0.6"/></geometry></collision>
                 <collision><geometry><sphere radius="9"/></geometry></collision></link>
  <link name="b"><inertial><origin xyz="0.1 0.2 0.3" rpy="0.1 0.2 0.3"/><mass value="2"/><inertia ixx="1" ixy="0" ixz="0" iyy="1" iyz="0" izz="1"/></inertial>
                 <collision><geometry><cylinder radius="0.05" length="0.3"/></geometry></collision></link>
  <link name="c"><collision><geometry><sphere radius="0.07"/></geometry></collision></link>
  <link name="d"/>
  <joint name="j_ab" type="continuous"><origin xyz="0 0 1" rpy="0.3 -0.2 0.9"/><parent link="a"/><child link="b"/><axis xyz="0 1 0"/></joint>
  <joint name="j_bc" type="prismatic"><origin xyz="0.5 0 0"/><parent link="b"/><child link="c"/><axis xyz="1 0 0"/>
         <limit lower="-0.2" upper="0.4" effort="1" velocity="1"/><dynamics damping="0.5"/></joint>
  <joint name="j_cd" type="fixed"><origin xyz="0 0 0.1"/><parent link="c"/><child link="d"/></joint>
</robot>
"""


@pytest.fixture
def small_urdf(tmp_path):
    p = tmp_path / "small.urdf"
    p.write_text(URDF_SMALL)
    return str(p)


def test_urdf_semantics(small_urdf):
    G = graph_from_urdf(small_urdf)
    assert list(G.nodes) == ["a", "b", "c", "d"]
    a, b, c = (G.nodes[n]["link"] for n in "abc")
    # first <collision> only; box -> half extents
    assert a._collision_shape.shape == Shape.CUBOID
    assert np.allclose(a._collision_shape._shape_info["half_extents"], [0.1, 0.2, 0.3])
    assert np.allclose(a._collision_shape.offset[:3, 3], [0, 0, 0.1])
    assert b._collision_shape.shape == Shape.CYLINDER and b._collision_shape._shape_info["height"] == 0.3
    assert c._collision_shape.shape == Shape.SPHERE
    assert G.nodes["d"]["link"]._collision_shape.shape == Shape.EMPTY
    # inertial origin -> Link._offset
    assert np.allclose(b._offset[:3, 3], [0.1, 0.2, 0.3]) and np.allclose(b._offset[:3, :3], rpy_matrix([0.1, 0.2, 0.3]))
    j_ab, j_bc, j_cd = G.edges[("a", "b")]["joint"], G.edges[("b", "c")]["joint"], G.edges[("c", "d")]["joint"]
    assert j_ab.type == Constraint.REVOLUTE          # continuous -> REVOLUTE
    assert j_ab.lower_limit == -np.inf and j_ab.upper_limit == np.inf   # no <limit> tag
    assert j_bc.type == Constraint.PRISMATIC and (j_bc.lower_limit, j_bc.upper_limit) == (-0.2, 0.4)
    assert j_cd.type == Constraint.FIXED and np.all(j_cd.axis == 0)
    # rpy = extrinsic xyz = Rz Ry Rx, identical to scipy's from_euler('xyz', ...)
    from scipy.spatial.transform import Rotation
    assert np.allclose(j_ab.offset[:3, :3], Rotation.from_euler("xyz", [0.3, -0.2, 0.9]).as_matrix(), atol=1e-15)


def test_urdf_errors(tmp_path):
    p = tmp_path / "bad.urdf"
    p.write_text('<robot name="x"><link name="a"/><link name="b"/><joint name="j" type="revolute"><parent link="a"/><child link="b"/></joint></robot>')
    with pytest.raises(ValueError, match="has no <axis>"):
        graph_from_urdf(str(p))
    p.write_text('<robot name="x"><link name="a"/><link name="b"/><joint name="j" type="floating"><parent link="a"/><child link="b"/></joint></robot>')
    with pytest.raises(ValueError, match="unsupported joint type"):
        graph_from_urdf(str(p))


def test_q_ordering_and_frame_cache_kinova():
    arm = workloads.kinova_arm()
    chain = arm._chain
    assert arm.dof == 7 and chain.joint_limits.shape == (7, 2)
    names = [j.name for j in chain._joints if j.type != Constraint.FIXED]
    assert names == [f"joint_{i}" for i in range(1, 8)]
    assert [chain._joint_to_index[j] for j in chain._joints if j.type != Constraint.FIXED] == list(range(7))
    assert arm._link_joint_sequence["base_link"] == tuple()
    off, axes, types, idxs = arm._link_joint_sequence["tool_frame"]
    assert off.shape == (8, 4, 4) and axes.shape == (7, 3) and list(idxs) == list(range(7)) and np.all(types == 0)
    # trailing fixed run of tool_frame = end_effector joint . tool_frame_joint
    G = chain._G
    tail = G.edges[("bracelet_link", "end_effector_link")]["joint"].offset @ G.edges[("end_effector_link", "tool_frame")]["joint"].offset
    assert np.allclose(off[-1], tail)
    off_b = arm._link_joint_sequence["bracelet_link"][0]
    assert off_b.shape == (7, 4, 4)            # no trailing offset for a frame that ends in a movable joint
    # flattened descriptor: folded joint offsets are the same matrices
    A = arm._arrays
    assert np.allclose(A["mov_offset"].reshape(7, 12), off[:7, :3, :].reshape(7, 12))
    f = A["frame_id"]["tool_frame"]
    assert A["frame_mov"][f] == 6 and np.allclose(A["frame_tail"].reshape(-1, 12)[f], tail[:3].reshape(12))
    assert np.allclose(A["frame_tail"].reshape(-1, 12)[A["frame_id"]["bracelet_link"]], np.eye(4)[:3].reshape(12))


def test_fixed_runs_fold_into_next_movable_joint():
    """arm.py:36-52: FIXED joints before a movable joint multiply into its offset; branches share parents."""
    w = workloads.fresh_world("fold")
    G = nx.DiGraph()
    for n in ["r", "f1", "m1", "f2", "m2", "side"]:
        G.add_node(n, link=DummyLink(1.0, np.eye(4), name=n))
    F1, M1 = trans_mat(pos=np.array([0.1, 0, 0])), trans_mat(pos=np.array([0, 0.2, 0]))
    F2, M2 = trans_mat(pos=np.array([0, 0, 0.3])), trans_mat(pos=np.array([0.4, 0, 0]))
    G.add_edge("r", "f1", joint=Joint(F1, np.zeros(3), Constraint.FIXED, name="jf1"))
    G.add_edge("f1", "m1", joint=Joint(M1, np.array([0, 0, 1.0]), Constraint.REVOLUTE, name="jm1"))
    G.add_edge("m1", "f2", joint=Joint(F2, np.zeros(3), Constraint.FIXED, name="jf2"))
    G.add_edge("f2", "m2", joint=Joint(M2, np.array([1.0, 0, 0]), Constraint.PRISMATIC, name="jm2"))
    G.add_edge("m1", "side", joint=Joint(M2, np.array([0, 1.0, 0]), Constraint.REVOLUTE, name="jside"))
    arm = Arm(GraphChain(G, world_name=w.name))
    A = arm._arrays
    assert arm.dof == 3
    mo = A["mov_offset"].reshape(-1, 12)
    q_of = {j.name: arm._chain._joint_to_index[j] for j in arm._chain._joints if j.type != Constraint.FIXED}
    order = [j.name for j in arm._chain._joints if j.type != Constraint.FIXED]
    i1, i2, iS = order.index("jm1"), order.index("jm2"), order.index("jside")
    assert np.allclose(mo[i1], (F1 @ M1)[:3].reshape(12)) and np.allclose(mo[i2], (F2 @ M2)[:3].reshape(12))
    assert A["mov_parent"][i1] == -1 and A["mov_parent"][i2] == i1 and A["mov_parent"][iS] == i1
    assert list(A["mov_qidx"]) == [q_of[n] for n in order]
    assert A["mov_type"][i2] == 1
    off, axes, types, idxs = arm._link_joint_sequence["m2"]
    assert off.shape == (2, 4, 4) and np.allclose(off[0], F1 @ M1) and np.allclose(off[1], F2 @ M2)


def test_graphchain_validation():
    w = workloads.fresh_world("val")
    G = nx.DiGraph()
    G.add_node("a", link=DummyLink(1.0, np.eye(4), name="a"))
    G.add_node("b", link=DummyLink(1.0, np.eye(4), name="b"))
    G.add_node("c", link=DummyLink(1.0, np.eye(4), name="c"))
    J = lambda: Joint(np.eye(4), np.array([0, 0, 1.0]), Constraint.REVOLUTE)
    G.add_edge("a", "b", joint=J()); G.add_edge("b", "c", joint=J()); G.add_edge("a", "c", joint=J())
    with pytest.raises(ValueError, match="Chain graph must be a tree"):
        GraphChain(G, world_name=w.name)
    G2 = nx.DiGraph()
    G2.add_node("a", link=DummyLink(1.0, np.eye(4), name="zzz"))
    G2.add_node("b", link=DummyLink(1.0, np.eye(4), name="b"))
    G2.add_edge("a", "b", joint=J())
    with pytest.raises(ValueError, match="must match root node name"):
        GraphChain(G2, world_name=w.name)
    with pytest.raises(ValueError, match="Link mass must be positive"):
        DummyLink(-1.0, np.eye(4))
    with pytest.raises(ValueError, match="Invalid shape type"):
        CollisionShape("box")


def test_self_collision_pair_rule_and_removals():
    """Effective rule of arm.py:190-223: all non-adjacent pairs of non-EMPTY links (the 'two joints apart'
    filter never runs in the reference); removals follow tests/_test_rrt.py:38-61."""
    world = workloads.fresh_world("pairs")
    arm = workloads.kinova_arm(world)
    geo = [l for l in arm._chain._links if l._collision_shape.shape != Shape.EMPTY]
    assert len(geo) == 10
    Gu = arm._chain._G.to_undirected()
    expect = {frozenset((a._name, b._name)) for i, a in enumerate(geo) for b in geo[i + 1:] if not Gu.has_edge(a._name, b._name)}
    got = {frozenset((a._name, b._name)) for a, b in arm.self_collision_pairs()}
    assert got == expect and len(got) == 38
    cubes = workloads.add_prm_obstacles(world)
    assert len(arm.collision_pairs()) == 38 + 30
    workloads.apply_rrt_pair_removals(arm)
    got = {frozenset((a._name, b._name)) for a, b in arm.self_collision_pairs()}
    assert frozenset(("base_link", "half_arm_1_link")) not in got
    assert frozenset(("spherical_wrist_1_link", "bracelet_link")) not in got
    assert frozenset(("base_link", "gripper")) in got
    S = arm.scene_arrays()
    assert S["self_pairs"].shape == (len(got), 2) and S["env_pairs"].shape == (30, 2)
    assert len(S["obstacles"]) == 3 and S["env_sharp"].sum() == 3       # camera box vs the three cubes
    # add back / remove an obstacle pair by name
    arm.add_collision_pair("base_link", "half_arm_1_link")
    assert frozenset(("base_link", "half_arm_1_link")) in {frozenset((a._name, b._name)) for a, b in arm.self_collision_pairs()}
    arm.remove_collision_pair("base_link", cubes[0].name)
    assert arm.scene_arrays()["env_pairs"].shape == (29, 2)
    with pytest.raises(ValueError, match="must be a valid object in the world"):
        arm.remove_collision_pair("base_link", "no_such_thing")


def test_margin_rule_and_descriptor_values():
    t, dims, m = descriptor.primitive_params(CollisionShape(Shape.CUBE, half_extents=np.array([0.5, 0.5, 0.5])), "bullet")
    assert (t, m) == (1, 0.04)
    t, dims, m = descriptor.primitive_params(CollisionShape(Shape.CYLINDER, radius=0.046, height=0.21), "bullet")
    assert t == 2 and m == pytest.approx(0.0046) and np.allclose(dims, [0.046, 0.105, 0])
    t, dims, m = descriptor.primitive_params(CollisionShape(Shape.SPHERE, radius=0.2), "bullet")
    assert (t, m) == (0, 0.2)
    assert descriptor.primitive_params(CollisionShape(Shape.CUBOID, half_extents=np.array([0.1, 0.2, 0.3])), "sharp")[2] == 0.0
    with pytest.raises(NotImplementedError):
        descriptor.primitive_params(CollisionShape(Shape.MESH, filename="x.obj"), "bullet")


def test_world_revision_tracks_obstacle_motion():
    world = workloads.fresh_world("rev")
    arm = workloads.kinova_arm(world)
    r0 = world.revision
    c = Cube(half_extent=0.1, mass=1.0, world_name=world.name)
    assert world.revision > r0
    r1 = world.revision
    c.position = np.array([1.0, 0, 0])
    assert world.revision > r1 and np.allclose(c.pose[:3, 3], [1, 0, 0])
    s = Sphere(mass=1.0, radius=0.1, world_name=world.name)
    cb = Cuboid(mass=0.0, half_extents=np.array([0.1, 0.2, 0.3]), static=True, world_name=world.name)
    assert {o.name for o in world.objects()} >= {c.name, s.name, cb.name}
    assert world.get_object(c.name) is c
    assert world.get_object(f"{world.name}:{arm._chain._name}:tool_frame") is arm._links_from_nodes["tool_frame"]


def test_spherical_joints_rejected():
    w = workloads.fresh_world("sph")
    G = nx.DiGraph()
    G.add_node("a", link=DummyLink(1.0, np.eye(4), name="a"))
    G.add_node("b", link=DummyLink(1.0, np.eye(4), name="b"))
    G.add_edge("a", "b", joint=Joint(np.eye(4), np.array([0, 0, 1.0]), Constraint.SPHERICAL))
    with pytest.raises(NotImplementedError, match="SPHERICAL"):
        Arm(GraphChain(G, world_name=w.name))


def test_static_pruning_bounds_hold_on_sampled_kinematics():
    """The two reach bounds behind the static pruning of obstacle pairs (descriptor.flatten_scene): every point of a link
    primitive stays within `reach` of the base origin and within `reach1` of the FIXED origin of its first movable joint;
    every obstacle pair the PRM scene drops at 0.05 m really stays farther than that from its obstacle (float64 oracle)."""
    from oracle import collision
    arm, cubes = workloads.prm_scene(device=None)
    so = collision.SceneOracle(arm)
    q = workloads.random_configs_np(arm, 4000, seed=5)
    A = arm._arrays
    reach = descriptor.primitive_reach(A, arm.joint_limits)
    reach1, root1 = descriptor.primitive_reach(A, arm.joint_limits, from_first_joint=True)
    assert np.all(reach1 <= reach + 1e-12) and set(root1.tolist()) == {-1, 0}      # -1: the base link's own primitive
    off = np.asarray(A["mov_offset"]).reshape(-1, 12)
    X = np.eye(4)
    X[:3, 3] = off[0, [3, 7, 11]]
    origin1 = (arm.base_pose @ X)[:3, 3]
    for node, o in zip(so.link_nodes, so.link_offsets):
        p = A["prim_of_link"][arm._links_from_nodes[node]]
        centers = (so.fk.forward_kinematics(q, node) @ o)[:, :3, 3]
        t = int(A["prim_type"][p])
        d = np.asarray(A["prim_dims"]).reshape(-1, 3)[p]
        brad = d[0] if t == 0 else (np.linalg.norm(d) if t == 1 else np.hypot(d[0], d[1]))
        assert np.linalg.norm(centers - arm.base_pose[:3, 3], axis=1).max() + brad <= reach[p] + 1e-9
        if root1[p] == 0:
            assert np.linalg.norm(centers - origin1, axis=1).max() + brad <= reach1[p] + 1e-9
    pruned = arm.scene_arrays(0.05)
    kept = {(int(a), pruned["obstacles"][int(o)].name) for a, o in pruned["env_pairs"]}
    assert len(kept) == 5 and pruned["n_env_pairs_before_pruning"] == 30
    D = so.pair_distances(q)
    n_dropped = 0
    for k, (a, b) in enumerate(so.pairs):
        if b < so.n_link:
            continue
        key = (A["prim_of_link"][so.links[a]], so.objs[b - so.n_link].name)
        if key not in kept:
            n_dropped += 1
            assert D[:, k].min() > 0.05
    assert n_dropped == 25
