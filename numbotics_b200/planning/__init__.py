"""Edge-validation half of the reference's planning package (numbotics/planning/__init__.py).

Only what the hot path needs: `unit_bspline` (the default trajectory) and the
`sampling_based` connectors.  Planners, IRIS and the roadmap graph are callers of
this path, not part of it.
"""
from .trajectories import unit_bspline

__all__ = ["unit_bspline"]
