"""Joint vocabulary on the boundary of the hot path.

Mirror of reference numbotics/physics/constraint.py:11-59.  The enum values are
PyBullet's `JOINT_*` integers (REVOLUTE 0, PRISMATIC 1, SPHERICAL 2, FIXED 4);
they cross the C-ABI unchanged as `joint_types` (include/nbx.h).
"""
import enum
from dataclasses import dataclass
from typing import Optional

import numpy as np


class Constraint(enum.Enum):
    FIXED = 4
    REVOLUTE = 0
    PRISMATIC = 1
    SPHERICAL = 2


@dataclass(frozen=True, eq=False)
class Joint:
    offset: np.ndarray
    axis: np.ndarray
    type: Constraint
    name: Optional[str] = None
    parent_pose: Optional[np.ndarray] = None
    child_pose: Optional[np.ndarray] = None
    damping: float = 0.01
    lower_limit: float = -np.inf
    upper_limit: float = np.inf
    max_velocity: float = np.inf
    max_effort: float = np.inf

    def __post_init__(self):
        off = np.asarray(self.offset).astype(np.float64)
        if off.shape != (4, 4):
            raise ValueError(f"Joint offset must be a 4x4 matrix, got {off.shape}")
        object.__setattr__(self, "offset", off)
        if self.type == Constraint.FIXED:
            object.__setattr__(self, "axis", np.zeros((3,), dtype=np.float64))
        else:
            object.__setattr__(self, "axis", np.asarray(self.axis).astype(np.float64))

    def __hash__(self):
        if self.name is None:
            return id(self)
        return hash(self.name)

    def __eq__(self, other):
        return self is other

    @property
    def dof(self):
        if self.type == Constraint.FIXED:
            return 0
        if self.type in (Constraint.REVOLUTE, Constraint.PRISMATIC):
            return 1
        if self.type == Constraint.SPHERICAL:
            return 3
        raise ValueError(f"Invalid constraint type: {self.type}")
