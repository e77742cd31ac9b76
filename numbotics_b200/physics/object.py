"""World obstacles: geometry + pose, nothing else.

Mirror of the constructor / pose surface of reference numbotics/physics/object.py
(`PhysicsObject` :15-66, pose/position/orientation :200-250, `Cube` :353-374,
`Cuboid` :378-399, `Sphere` :403-424, `Cylinder`).  Mass, friction, damping,
velocities and contact queries belong to the PyBullet simulation passthrough and
are out of scope; `mass` is accepted and stored so call sites stay unchanged.
"""
import numpy as np

from .shape import Shape, CollisionShape, parse_shape_kwargs
from .world import get_world


class PhysicsObject:

    def __init__(self, mass: float, static: bool, collision_shape=None, visual_shape=None, **kwargs):
        if mass < 0:
            raise ValueError("Link mass must be positive")
        self._mass = mass if not static else 0
        self._static = static
        world = get_world(name=kwargs.pop("world_name", None))
        self._collision_shape = collision_shape if collision_shape is not None else CollisionShape(Shape.EMPTY)
        self._world_name = world.name
        self._entity_id = world._new_entity_id()
        self._pose = np.eye(4)
        self._name = kwargs.pop("name", f"physics_object_{self._entity_id}")
        for key in list(kwargs.keys()):
            if hasattr(self, key):
                setattr(self, key, kwargs.pop(key))
        if kwargs:
            raise ValueError(f"Unexpected keyword arguments: {', '.join(kwargs.keys())}")
        world.register(self)

    def __hash__(self):
        return hash(self.name)

    def __eq__(self, other):
        return self is other

    def __repr__(self):
        return f"<{type(self).__name__} {self.name}>"

    @property
    def name(self):
        return f"{self._world_name}:{self._name}"

    @property
    def world(self):
        return get_world(name=self._world_name)

    @property
    def mass(self):
        return self._mass

    @property
    def pose(self):
        return self._pose.copy()

    @pose.setter
    def pose(self, T):
        T = np.asarray(T, dtype=np.float64)
        if T.shape != (4, 4):
            raise ValueError("pose must be a 4x4 matrix")
        self._pose = T.copy()
        self.world.touch()

    @property
    def position(self):
        return self._pose[:3, 3].copy()

    @position.setter
    def position(self, p):
        self._pose[:3, 3] = np.asarray(p, dtype=np.float64)
        self.world.touch()

    @property
    def orientation(self):
        return self._pose[:3, :3].copy()

    @orientation.setter
    def orientation(self, R):
        R = np.asarray(R, dtype=np.float64)
        if R.shape != (3, 3):
            raise ValueError("orientation must be a 3x3 matrix")
        self._pose[:3, :3] = R
        self.world.touch()


class Cube(PhysicsObject):
    def __init__(self, mass: float, half_extent: float, static: bool = False, **kwargs):
        kwargs, shape_info = parse_shape_kwargs(kwargs)
        he = np.array([half_extent, half_extent, half_extent], dtype=np.float64)
        super().__init__(mass, static, CollisionShape(Shape.CUBE, half_extents=he, **shape_info), None, **kwargs)


class Cuboid(PhysicsObject):
    def __init__(self, mass: float, half_extents, static: bool = False, **kwargs):
        kwargs, shape_info = parse_shape_kwargs(kwargs)
        he = np.asarray(half_extents, dtype=np.float64)
        super().__init__(mass, static, CollisionShape(Shape.CUBOID, half_extents=he, **shape_info), None, **kwargs)


class Sphere(PhysicsObject):
    def __init__(self, mass: float, radius: float, static: bool = False, **kwargs):
        kwargs, shape_info = parse_shape_kwargs(kwargs)
        super().__init__(mass, static, CollisionShape(Shape.SPHERE, radius=radius, **shape_info), None, **kwargs)


class Cylinder(PhysicsObject):
    def __init__(self, mass: float, radius: float, height: float, static: bool = False, **kwargs):
        kwargs, shape_info = parse_shape_kwargs(kwargs)
        super().__init__(mass, static,
                         CollisionShape(Shape.CYLINDER, radius=radius, height=height, **shape_info), None, **kwargs)
