"""Collision targets beyond world obstacles, and per-pair distance gradients:

* the links of ANOTHER chain in the world are targets of `Arm.collisions / closest_to / in_collision`
  (reference robots/arm.py:239-241), posed by that chain's current `configuration`;
* `Arm.jacobian_proximity(q, obj, link)` (reference robots/arm.py:621-632) for arbitrary allowed pairs, batched --
  against central differences of the oracle's pair distances."""
import numpy as np
import pytest

from numbotics_b200 import workloads
from numbotics_b200.math import trans_mat
from numbotics_b200.physics import GraphChain
from numbotics_b200.robots import Arm
from oracle import collision, fkjac

pytestmark = pytest.mark.gpu


def _second_robot(world, q_state):
    G = workloads.serial_chain_graph(4, link_length=0.22, with_geometry=True)
    chain = GraphChain(G, world_name=world.name, base_pose=trans_mat(pos=np.array([0.45, 0.05, 0.25])))
    chain.configuration = q_state
    return chain


def _foreign_poses(chain):
    """Link poses of the other chain from the float64 FK oracle (independent of the CUDA path)."""
    model = Arm(chain, device=0)              # only its reference-shaped per-frame cache is read
    ora = fkjac.ArmOracle(model._link_joint_sequence, model.dof, chain.base_pose)
    return {link: ora.forward_kinematics(chain.configuration, node) for node, link in model._links_from_nodes.items()}


def test_links_of_another_chain_are_collision_targets(cuda):
    world = workloads.fresh_world("two_robots")
    arm = workloads.kinova_arm(world, device=0)
    workloads.apply_rrt_pair_removals(arm)      # without them neighbouring wrist links overlap in every configuration
    other = _second_robot(world, np.array([0.3, -0.5, 0.7, 0.2]))
    pairs = arm.collision_pairs()
    foreign_links = [l for l in other._links if l._collision_shape.shape.name != "EMPTY"]
    assert len(foreign_links) == 5
    assert all(any(p[1] is fl for p in pairs) for fl in foreign_links)
    so = collision.SceneOracle(arm, foreign_link_poses=_foreign_poses(other))
    q = workloads.random_configs_np(arm, 3000, seed=41).astype(np.float32)
    ref = so.min_distance(q.astype(np.float64))
    d = arm.min_distance(q)
    assert (ref < 0).sum() > 100 and np.abs(d - ref).max() < 3e-4
    hit = arm.in_collision(q)
    clear = np.abs(ref) > 1e-4
    assert np.array_equal(hit[clear], (ref < 0)[clear])
    # the closest target can be a link of the other robot, reported as that Link object
    rec = arm.closest_to(q)
    tg = rec.target
    assert any(t in foreign_links for t in tg)
    # the other robot moves: same arm, new verdicts
    other.configuration = np.array([-1.2, 0.9, -0.3, 0.5])
    so2 = collision.SceneOracle(arm, foreign_link_poses=_foreign_poses(other))
    ref2 = so2.min_distance(q.astype(np.float64))
    d2 = arm.min_distance(q)
    assert np.abs(ref2 - ref).max() > 0.02 and np.abs(d2 - ref2).max() < 3e-4
    # pairs with a foreign link can be voided like any other (by name, reference arm.py:311-366)
    n0 = len(arm.pair_distances(q[:4])[0])
    arm.remove_collision_pair("half_arm_1_link", foreign_links[0].name)
    assert len(arm.pair_distances(q[:4])[0]) == n0 - 1


def test_jacobian_proximity_for_arbitrary_pairs(cuda):
    arm, cubes = workloads.prm_scene(device=0)
    so = collision.SceneOracle(arm)
    q = workloads.random_configs_np(arm, 400, seed=42)
    cube = cubes[2]
    link = arm._links_from_nodes["bracelet_link"]
    # batched, one pair -> (..., dof)
    J = arm.jacobian_proximity(q, cube, link)
    assert J.shape == (400, 7)
    # the same row inside the all-pairs gradient, and the reference's shape rules for one configuration
    pairs, rec = arm.pair_distances(q, gradient=True)
    col = [i for i, (s_, t_) in enumerate(pairs) if s_ is link and t_ is cube]
    assert len(col) == 1 and np.array_equal(rec["gradient"][:, col[0]], J)
    J_all = arm.jacobian_proximity(q[0], cube)
    n_cube = sum(1 for _, t_ in pairs if t_ is cube)
    assert J_all.shape == (n_cube, 7) and arm.jacobian_proximity(q[0], cube, link).shape == (7,)
    with pytest.raises(ValueError, match="not valid"):
        arm.jacobian_proximity(q[0], cube, arm._links_from_nodes["tool_frame"])
    # central differences of the oracle's distance of that pair
    li = so.links.index(link)
    oi = so.n_link + so.objs.index(cube)
    c = [k for k, (a, b) in enumerate(so.pairs) if (a, b) == (li, oi)][0]
    h = 1e-5
    fd = np.empty((400, 7))
    for j in range(7):
        e = np.zeros(7); e[j] = h
        fd[:, j] = (so.pair_distances(q + e)[:, c] - so.pair_distances(q - e)[:, c]) / (2 * h)
    err = np.abs(J - fd).max(axis=1)
    assert np.median(err) < 2e-4 and (err < 5e-3).mean() > 0.95
    # a self pair: joints above both links cancel, the gradient is that of the relative motion
    la, lb = arm._links_from_nodes["shoulder_link"], arm._links_from_nodes["bracelet_link"]
    Js = arm.jacobian_proximity(q, la, lb) if any((s_ is lb and t_ is la) or (s_ is la and t_ is lb) for s_, t_ in pairs) else None
    if Js is not None:
        assert Js.shape == (400, 7) and np.abs(Js[:, 0]).max() < 1e-6      # joint 1 moves both links together
