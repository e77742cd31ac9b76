"""Host-side model types read by the batched kinematics / collision path.

Same names as reference numbotics/physics/__init__.py:1-51, restricted to what
the hot path consumes: joint/link/shape vocabulary, the
tree model with the reference's q ordering, obstacles with a pose, and a registry.
"""
from .shape import Shape, CollisionShape, parse_shape_kwargs
from .constraint import Constraint, Joint
from .world import World, get_world
from .object import PhysicsObject, Cube, Cuboid, Sphere, Cylinder
from .chain import Link, DummyLink, BasicLink, Chain, GraphChain

__all__ = [
    "Shape", "CollisionShape", "parse_shape_kwargs", "Constraint", "Joint", "World", "get_world",
    "PhysicsObject", "Cube", "Cuboid", "Sphere", "Cylinder",
    "Link", "DummyLink", "BasicLink", "Chain", "GraphChain",
]
