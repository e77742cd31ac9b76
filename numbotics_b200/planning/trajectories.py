"""`unit_bspline`: clamped B-spline through control points on t in [0, 1].

Mirror of reference numbotics/planning/trajectories.py:6-21 (same validation, same
knot vector).  With two control points and degree 1 -- the default
`ConnectorParams.trajectory_func` (connectors.py:16) -- this is the linear
interpolation `(1 - t) * start + t * goal`; `LinearTrajectory` evaluates exactly
that without scipy and is what the fused edge kernel implements on the device.
"""
import numpy as np


class LinearTrajectory:
    """Degree-1, two-control-point special case of `unit_bspline`."""

    def __init__(self, start, goal):
        self.start = np.asarray(start, dtype=np.float64)
        self.goal = np.asarray(goal, dtype=np.float64)

    def __call__(self, t):
        t = np.asarray(t, dtype=np.float64)
        if t.ndim == 0:
            return (1.0 - t) * self.start + t * self.goal
        return (1.0 - t)[..., None] * self.start + t[..., None] * self.goal


def unit_bspline(control_points, degree: int = 1):
    control_points = np.asarray(control_points)
    if control_points.ndim != 2:
        raise ValueError("control_points must be a 2D array (B x n)")
    B, _ = control_points.shape
    if degree >= B:
        raise ValueError("Degree must be less than the number of control points")
    if degree == 1 and B == 2:
        return LinearTrajectory(control_points[0], control_points[1])
    from scipy.interpolate import BSpline
    knots = np.concatenate((np.zeros(degree), np.linspace(0, 1, B - degree + 1), np.ones(degree)))
    return BSpline(knots, control_points, degree)
