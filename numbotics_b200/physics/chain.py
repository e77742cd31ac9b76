"""Kinematic tree model: the *source* the batched path flattens into descriptors.

Mirror of the model half of reference numbotics/physics/chain.py:

* `Link` / `BasicLink` / `DummyLink`   -> :28-72, :381-405
* `Chain` joint-index maps             -> :516-536 (order of q columns!)
* `GraphChain(G)` tree validation and link / joint ordering -> :1084-1182
* `GraphChain.from_urdf`               -> :1185-1187

Everything PyBullet-backed (multibody creation, dynamics, contacts, wrenches,
mass matrix) is out of scope.  What *is* reproduced
exactly is the ordering contract: `_links = [root] + topological order`,
`_joints[i]` is the edge into `_links[i+1]`, and q-column `j` is the j-th
non-fixed joint in that order.
"""
import copy

import numpy as np
import networkx as nx

from .shape import Shape, CollisionShape, parse_shape_kwargs
from .constraint import Constraint, Joint
from .world import get_world


class Link:

    def __init__(self, mass: float, offset, collision_shape=None, visual_shape=None, **kwargs):
        if mass < 0:
            raise ValueError("Link mass must be positive")
        self._offset = np.asarray(offset, dtype=np.float64)
        if self._offset.shape != (4, 4):
            raise ValueError("Link offset must be a 4x4 matrix")
        self._mass = mass
        self._collision_shape = collision_shape if collision_shape is not None else CollisionShape(Shape.EMPTY)
        self._body_name = None
        self._index = None
        self._world_name = None
        self._inertia_diagonal = kwargs.pop("inertia_diagonal", None)
        self._name = kwargs.pop("name", None)
        for key in list(kwargs.keys()):
            if hasattr(self, key):
                setattr(self, key, kwargs.pop(key))
        if kwargs:
            raise ValueError(f"Unexpected keyword arguments: {', '.join(kwargs.keys())}")

    @property
    def index(self):
        return self._index

    @index.setter
    def index(self, index: int):
        self._index = index
        if self._name is None:
            self._name = f"link_{self._index}"

    @property
    def name(self):
        if self._body_name is None or self._world_name is None:
            raise ValueError("Link is not registered in the world")
        return f"{self._world_name}:{self._body_name}:{self._name}"

    def __hash__(self):
        return id(self)

    def __eq__(self, other):
        return self is other

    def __repr__(self):
        return f"<Link {self._name} {self._collision_shape.shape.name}>"


class DummyLink(Link):
    def __init__(self, mass: float, offset, **kwargs):
        super().__init__(mass, offset, None, None, **kwargs)


class BasicLink(Link):
    def __init__(self, mass: float, offset, shape: Shape, **kwargs):
        kwargs, shape_info = parse_shape_kwargs(kwargs)
        super().__init__(mass, offset, CollisionShape(shape, **shape_info), None, **kwargs)


class Chain:
    """Ordered links + joints with the reference's q-index map."""

    def __init__(self, links, joints, static_base: bool = True, **kwargs):
        world = get_world(name=kwargs.pop("world_name", getattr(self, "_world_name", None)))
        self._world_name = world.name
        self._entity_id = world._new_entity_id()
        self._static = False
        self._static_base = static_base if links[0]._mass > 0 else True
        self._links = links
        self._joints = joints
        self._name = kwargs.pop("name", f"chain_{self._entity_id}")
        base_pose = kwargs.pop("base_pose", None)
        if kwargs:
            raise ValueError(f"Unexpected keyword arguments: {', '.join(kwargs.keys())}")

        links[0].index = -1
        for i, link in enumerate(links[1:]):
            link.index = i
        for link in links:
            link._body_name = self._name
            link._world_name = self._world_name

        # q-column order: reference physics/chain.py:516-534
        self._joint_to_index = {}
        j_idx = 0
        for joint in self._joints:
            if joint.type in (Constraint.PRISMATIC, Constraint.REVOLUTE):
                self._joint_to_index[joint] = j_idx
                j_idx += 1
            elif joint.type == Constraint.SPHERICAL:
                self._joint_to_index[joint] = [j_idx, j_idx + 1, j_idx + 2]
                j_idx += 3
        self._dof = j_idx
        self.joint_limits = np.array(
            [[j.lower_limit, j.upper_limit] for j in self._joints if j.type != Constraint.FIXED],
            dtype=np.float64).reshape(-1, 2)
        # PyBullet reports the pose of the base *inertial* frame (reference chain.py:972-975 with
        # baseInertialFramePosition = base_link._offset, :433-434): a chain created at the world
        # origin therefore reports base_link._offset.  The hot path takes it as a given 4x4.
        self._base_pose = np.array(links[0]._offset, dtype=np.float64) if base_pose is None else np.asarray(base_pose, dtype=np.float64)
        # the chain's CURRENT joint state (reference physics/chain.py:679-685 keeps it in PyBullet): only read when this chain
        # is a collision TARGET of another arm (reference robots/arm.py:239-241), whose scene then holds its links as posed here
        self._configuration = np.zeros(self._dof, dtype=np.float64)
        world.register(self)

    def __hash__(self):
        return hash(self.name)

    def __eq__(self, other):
        return self is other

    @property
    def name(self):
        return f"{self._world_name}:{self._name}"

    @property
    def world(self):
        return get_world(name=self._world_name)

    @property
    def dof(self):
        return self._dof

    @property
    def configuration(self):
        return self._configuration.copy()

    @configuration.setter
    def configuration(self, q):
        q = np.asarray(q, dtype=np.float64).reshape(-1)
        if q.shape[0] != self._dof:
            raise ValueError(f"q must have {self._dof} elements")
        self._configuration = q.copy()
        self.world.touch()

    @property
    def base_pose(self):
        return self._base_pose.copy()

    @base_pose.setter
    def base_pose(self, T):
        T = np.asarray(T, dtype=np.float64)
        if T.shape != (4, 4):
            raise ValueError("base_pose must be a 4x4 matrix")
        self._base_pose = T.copy()
        self.world.touch()


class GraphChain(Chain):
    """Tree of links (nodes, attribute `link`) and joints (edges, attribute `joint`)."""

    def __init__(self, G: nx.DiGraph, static_base: bool = False, **kwargs):
        if not nx.is_tree(G):
            raise ValueError("Chain graph must be a tree")
        if not nx.is_directed_acyclic_graph(G):
            raise ValueError("Chain graph must be a directed acyclic graph")
        G = copy.deepcopy(G)
        order = list(nx.topological_sort(G))
        root_node = order[0]
        base_link = G.nodes[root_node]["link"]
        if base_link._name is not None and root_node != base_link._name:
            raise ValueError(f"Base link name {base_link._name} must match root node name {root_node}")
        if base_link._name is None:
            base_link._name = root_node if isinstance(root_node, str) else None

        links, joints = [], []
        for node in order[1:]:
            parents = list(G.predecessors(node))
            if len(parents) != 1:
                raise ValueError(f"Node {node} has {len(parents)} parents, expected 1")
            link = G.nodes[node]["link"]
            joint = G.edges[(parents[0], node)]["joint"]
            if link._name is not None and link._name != node:
                raise ValueError(f"Link name {link._name} must match node name {node}")
            if link._name is None and isinstance(node, str):
                link._name = node
            links.append(link)
            joints.append(joint)
        self._G = G
        self._node_order = order
        super().__init__([base_link] + links, joints, static_base, **kwargs)

    @classmethod
    def from_urdf(cls, urdf_path: str, **kwargs):
        from .urdf import chain_from_urdf
        return chain_from_urdf(urdf_path, **kwargs)
