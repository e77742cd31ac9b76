"""Synthetic workloads of BASELINE.json's configs (shared by tests, smoke and bench.py).

* Kinova 7-DoF arm from the authored `assets/kinova_cyl.urdf` (the reference's own URDF is
  git-ignored).
* The PRM scene of reference numbotics/tests/_test_rrt.py:26-61: three cubes of half extent
  0.5 and the manual collision-pair removals.
* A programmatic N-DoF serial GraphChain (BASELINE config 5), built the way the reference
  requires: an nx.DiGraph whose node keys equal each link's name, edges carrying Joints
  (numbotics/physics/chain.py:1084-1145).
"""
import os
from itertools import combinations

import numpy as np
import networkx as nx

from .math import trans_mat
from .physics import World, GraphChain, Cube, Joint, Constraint, BasicLink, DummyLink, Shape
from .physics.world import WORLD_INSTANCES

KINOVA_URDF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets", "kinova_cyl.urdf")

RRT_WRIST_LINKS = [
    "spherical_wrist_1_link", "spherical_wrist_2_link", "bracelet_link", "end_effector_link",
    "camera_link", "camera_depth_frame", "camera_color_frame", "tool_frame",
    "robotiq_arg2f_base_link", "gripper", "camera",
]


def fresh_world(prefix: str = "w") -> World:
    """A new uniquely named World that becomes the selected one."""
    i = 0
    while f"{prefix}_{i}" in WORLD_INSTANCES:
        i += 1
    return World(name=f"{prefix}_{i}")


def kinova_arm(world: World | None = None, device=None, margin_model: str = "bullet"):
    from .robots import Arm
    world = world or fresh_world("kinova")
    chain = GraphChain.from_urdf(KINOVA_URDF, world_name=world.name)
    return Arm(chain, device=device, margin_model=margin_model)


def add_prm_obstacles(world: World):
    """Cubes of reference tests/_test_rrt.py:29-35 (kept alive by the returned list)."""
    cubes = [Cube(half_extent=0.5, mass=1.0, world_name=world.name) for _ in range(3)]
    cubes[0].position = np.array([1.35, -1.5, 0.25])
    cubes[1].position = np.array([1.35, 1.5, 0.25])
    cubes[2].position = np.array([1.35, 0.0, 1.1])
    return cubes


def apply_rrt_pair_removals(arm):
    """Manual pair pruning of reference tests/_test_rrt.py:38-61."""
    arm.remove_collision_pair("base_link", "half_arm_1_link")
    arm.remove_collision_pair("half_arm_2_link", "spherical_wrist_1_link")
    for a, b in combinations(RRT_WRIST_LINKS, 2):
        arm.remove_collision_pair(a, b)


def prm_scene(device=None, margin_model: str = "bullet"):
    """(arm, cubes): Kinova + the three cubes + the pair removals (BASELINE configs 3 and 4)."""
    world = fresh_world("prm")
    arm = kinova_arm(world, device=device, margin_model=margin_model)
    cubes = add_prm_obstacles(world)
    apply_rrt_pair_removals(arm)
    return arm, cubes


def serial_chain_graph(n_dof: int, link_length: float = 0.1, with_geometry: bool = False) -> nx.DiGraph:
    """N revolute joints about local z, each offset = Trans(link_length, 0, 0) . Rx(+-pi/2), + a fixed tool."""
    G = nx.DiGraph()

    def link(name):
        if with_geometry:
            return BasicLink(1.0, np.eye(4), Shape.SPHERE, radius=0.03, name=name)
        return DummyLink(1.0, np.eye(4), name=name)

    G.add_node("link_0", link=link("link_0"))
    for i in range(n_dof):
        roll = (np.pi / 2.0) if i % 2 == 0 else (-np.pi / 2.0)
        c, s = np.cos(roll), np.sin(roll)
        R = np.array([[1.0, 0.0, 0.0], [0.0, c, -s], [0.0, s, c]])
        off = trans_mat(pos=np.array([link_length, 0.0, 0.0]), orn=R)
        G.add_node(f"link_{i + 1}", link=link(f"link_{i + 1}"))
        G.add_edge(f"link_{i}", f"link_{i + 1}",
                   joint=Joint(off, np.array([0.0, 0.0, 1.0]), Constraint.REVOLUTE, name=f"joint_{i + 1}",
                               lower_limit=-np.pi, upper_limit=np.pi))
    G.add_node("tool", link=DummyLink(0.1, np.eye(4), name="tool"))
    G.add_edge(f"link_{n_dof}", "tool",
               joint=Joint(trans_mat(pos=np.array([0.05, 0.0, 0.0])), np.zeros(3), Constraint.FIXED, name="tool_joint"))
    return G


def serial_arm(n_dof: int, device=None, with_geometry: bool = False):
    from .robots import Arm
    world = fresh_world(f"serial{n_dof}")
    chain = GraphChain(serial_chain_graph(n_dof, with_geometry=with_geometry), world_name=world.name)
    return Arm(chain, device=device)


def sample_limits(limits: np.ndarray):
    """Per-joint sampling box; infinite (continuous) limits fall back to [-pi, pi]."""
    lo = np.where(np.isfinite(limits[:, 0]), limits[:, 0], -np.pi)
    hi = np.where(np.isfinite(limits[:, 1]), limits[:, 1], np.pi)
    return lo, hi


def random_configs_np(arm, n: int, seed: int = 0, dtype=np.float64) -> np.ndarray:
    lo, hi = sample_limits(arm.joint_limits)
    return np.random.default_rng(seed).uniform(lo, hi, size=(n, arm.dof)).astype(dtype)


def random_configs_cuda(arm, n: int, seed: int = 0, dtype=None, device=None):
    import torch
    dtype = dtype or torch.float32
    dev = torch.device("cuda", arm.device_index if device is None else device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    lo, hi = sample_limits(arm.joint_limits)
    lo = torch.as_tensor(lo, dtype=dtype, device=dev)
    hi = torch.as_tensor(hi, dtype=dtype, device=dev)
    u = torch.rand((n, arm.dof), generator=g, dtype=dtype, device=dev)
    return lo + u * (hi - lo)


def knn_edges_cuda(nodes, k: int, chunk: int = 4096):
    """Candidate PRM edges: each node -> its k nearest neighbours in L2 (BASELINE config 4).

    Brute-force kNN with torch on the device; this is workload generation, not the timed path.
    Returns (starts, goals) of shape (N*k, dof).
    """
    import torch
    N = nodes.shape[0]
    idx = torch.empty((N, k), dtype=torch.long, device=nodes.device)
    sq = (nodes * nodes).sum(-1)
    for s in range(0, N, chunk):
        block = nodes[s:s + chunk]
        d2 = sq[s:s + chunk, None] - 2.0 * block @ nodes.T + sq[None, :]
        ar = torch.arange(block.shape[0], device=nodes.device)
        d2[ar, ar + s] = float("inf")
        idx[s:s + chunk] = torch.topk(d2, k, dim=1, largest=False).indices
    starts = nodes[:, None, :].expand(N, k, nodes.shape[1]).reshape(N * k, -1).contiguous()
    goals = nodes[idx.reshape(-1)].contiguous()
    return starts, goals
