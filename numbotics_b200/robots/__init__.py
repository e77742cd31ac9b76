"""Batched kinematics / collision call surface (reference numbotics/robots/__init__.py)."""
from .arm import Arm
from .helpers import nb_compute_transformation, nb_compute_jacobian

__all__ = ["Arm", "nb_compute_transformation", "nb_compute_jacobian"]
