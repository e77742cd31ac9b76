"""Collision primitive vocabulary of the hot path.

Mirror of the *collision* half of reference numbotics/utils/shape.py:17-25,53-114
and numbotics/physics/collision.py:36-46.  The reference hands these parameters to
`pybullet.createCollisionShape`; here they are flattened into the device scene
descriptor instead (robots/descriptor.py).  Parameterisation kept identical:

* CUBE / CUBOID : `half_extents (3,)`
* SPHERE        : `radius`
* CYLINDER      : `radius`, `height` (= full length, axis = local z)
* every shape   : optional 4x4 `offset` in the owning link's frame

MESH / PLANE / CAPSULE are accepted as vocabulary (so that reference scripts
parse) but are rejected by the descriptor builder: the north star scopes the path
to cylinders / boxes / spheres.
"""
from enum import Enum, auto

import numpy as np


class Shape(Enum):
    CUBE = auto()
    CUBOID = auto()
    SPHERE = auto()
    CYLINDER = auto()
    CAPSULE = auto()
    MESH = auto()
    PLANE = auto()
    EMPTY = auto()


_SHAPE_KWARGS = {
    "offset", "half_extents", "radius", "height", "width", "normal", "filename",
    "color", "mesh_scale", "auto_center", "convex_decomposition",
}


def parse_shape_kwargs(kwargs: dict):
    """Split `kwargs` into (remaining, shape_info) exactly like reference utils/shape.py:130-135."""
    shape_info = {}
    for key in list(kwargs.keys()):
        if key in _SHAPE_KWARGS:
            shape_info[key] = kwargs.pop(key)
    return kwargs, shape_info


class CollisionShape:
    """A primitive type plus its parameters (reference physics/collision.py:36-46)."""

    def __init__(self, shape: Shape, **kwargs):
        if not isinstance(shape, Shape):
            raise ValueError(f"Invalid shape type: {shape}")
        self.shape = shape
        self._shape_info = parse_shape_kwargs(kwargs)[1]

    @property
    def offset(self) -> np.ndarray:
        return np.asarray(self._shape_info.get("offset", np.eye(4)), dtype=np.float64)

    def __repr__(self):
        return f"CollisionShape({self.shape.name}, {self._shape_info})"
