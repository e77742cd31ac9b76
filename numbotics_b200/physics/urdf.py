"""PyBullet-free URDF -> GraphChain front-end.

Follows the URDF semantics of reference numbotics/physics/helpers.py:176-356
(`_chain_from_urdf`) without urdf_parser_py / PyBullet:

* link inertial origin -> `Link._offset`                      (helpers.py:200-229)
* only the FIRST `<collision>` of a link is read              (helpers.py:233)
* box -> CUBOID with half_extents = size / 2                  (helpers.py:243-245)
* sphere -> SPHERE radius; cylinder -> CYLINDER radius,height=length (helpers.py:246-251)
* mesh -> MESH (parsed, rejected later by the descriptor builder)
* rpy -> scipy `from_euler('xyz', rpy)` = Rz(y) Ry(p) Rx(r)   (helpers.py:313-316)
* `continuous` -> REVOLUTE; other types by upper-cased name   (helpers.py:306-310)
* `<limit>` lower/upper default to 0.0 when the tag exists but the attribute does
  not (urdf_parser_py `JointLimit` defaults), +-inf when there is no tag
  (`Joint` dataclass defaults, constraint.py:27-28)
* `<dynamics damping>` is dropped: the reference reads it into `joint_args`
  before `joint_args` is (re)assigned (helpers.py:320-330), so it never reaches
  the `Joint`; damping is not on the kinematics path anyway.
* nodes are added in `<link>` document order and edges in `<joint>` document
  order, which is what fixes networkx's topological order and therefore the
  order of q columns (chain.py:1096-1145, 516-534).
"""
import xml.etree.ElementTree as ET

import numpy as np
import networkx as nx

from ..math.spatial import trans_mat, rpy_matrix
from .shape import Shape, CollisionShape
from .constraint import Constraint, Joint
from .chain import Link, GraphChain


def _floats(text, n=None):
    vals = [float(x) for x in text.split()]
    if n is not None and len(vals) != n:
        raise ValueError(f"expected {n} numbers, got {text!r}")
    return np.array(vals, dtype=np.float64)


def _origin(elem):
    """`<origin xyz rpy>` child of `elem` -> 4x4 (identity when absent)."""
    if elem is None:
        return np.eye(4)
    o = elem.find("origin")
    if o is None:
        return np.eye(4)
    xyz = _floats(o.get("xyz", "0 0 0"), 3)
    rpy = _floats(o.get("rpy", "0 0 0"), 3)
    return trans_mat(pos=xyz, orn=rpy_matrix(rpy))


def _collision_shape(link_elem):
    col = link_elem.find("collision")  # first one only
    if col is None:
        return CollisionShape(Shape.EMPTY)
    offset = _origin(col)
    geom = col.find("geometry")
    if geom is None:
        raise ValueError("<collision> without <geometry>")
    box, sph, cyl, mesh = (geom.find(t) for t in ("box", "sphere", "cylinder", "mesh"))
    if box is not None:
        return CollisionShape(Shape.CUBOID, offset=offset, half_extents=_floats(box.get("size"), 3) / 2.0)
    if sph is not None:
        return CollisionShape(Shape.SPHERE, offset=offset, radius=float(sph.get("radius")))
    if cyl is not None:
        return CollisionShape(Shape.CYLINDER, offset=offset,
                              radius=float(cyl.get("radius")), height=float(cyl.get("length")))
    if mesh is not None:
        scale = _floats(mesh.get("scale"), 3) if mesh.get("scale") else np.ones(3)
        return CollisionShape(Shape.MESH, offset=offset, filename=mesh.get("filename"), mesh_scale=scale)
    raise ValueError("unsupported collision geometry")


def graph_from_urdf(urdf_path: str) -> nx.DiGraph:
    root = ET.parse(urdf_path).getroot()
    if root.tag != "robot":
        raise ValueError("not a URDF: root element must be <robot>")
    G = nx.DiGraph()
    for le in root.findall("link"):
        inertial = le.find("inertial")
        mass = 0.0
        if inertial is not None and inertial.find("mass") is not None:
            mass = float(inertial.find("mass").get("value"))
        link = Link(mass=mass, offset=_origin(inertial), collision_shape=_collision_shape(le),
                    name=le.get("name"))
        G.add_node(le.get("name"), link=link)

    for je in root.findall("joint"):
        jt = je.get("type")
        if jt == "continuous":
            ctype = Constraint.REVOLUTE
        else:
            try:
                ctype = getattr(Constraint, jt.upper())
            except AttributeError:
                raise ValueError(f"unsupported joint type: {jt}") from None
        ax = je.find("axis")
        if ax is not None:
            axis = _floats(ax.get("xyz", "1 0 0"), 3)
        else:
            if jt != "fixed":
                raise ValueError(f"joint {je.get('name')} is {jt} but has no <axis>")
            axis = np.zeros(3)
        args = {}
        lim = je.find("limit")
        if lim is not None:
            args["lower_limit"] = float(lim.get("lower", 0.0))
            args["upper_limit"] = float(lim.get("upper", 0.0))
            if lim.get("velocity") is not None:
                args["max_velocity"] = float(lim.get("velocity"))
            if lim.get("effort") is not None:
                args["max_effort"] = float(lim.get("effort"))
        joint = Joint(name=je.get("name"), offset=_origin(je), axis=axis, type=ctype, **args)
        parent = je.find("parent").get("link")
        child = je.find("child").get("link")
        for n in (parent, child):
            if n not in G.nodes:
                raise ValueError(f"joint {je.get('name')} references unknown link {n}")
        G.add_edge(parent, child, joint=joint)
    return G


def chain_from_urdf(urdf_path: str, **kwargs) -> GraphChain:
    return GraphChain(graph_from_urdf(urdf_path), **kwargs)
