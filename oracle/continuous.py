"""Float64 restatement of what `ContinuousConnector` decides.  TEST INFRASTRUCTURE ONLY.

Reference numbotics/planning/sampling_based/connectors.py:108-185: the edge `start -> goal` (up to `T_f` under steer
semantics, :160-164) is cut at the grid `arange(0, T_f, resolution / distance) U {T_f}` (:123-124 / :166-167) and
rejected when, on some segment, SLSQP finds a `t` with `validity_checker(trajectory(t)) <= 0` (:131-143).  For an exact
solver that is: the edge is valid iff the clearance stays positive on the whole of `[0, T_f]`.  This module states that
directly -- the clearance (`oracle.collision.SceneOracle.min_distance`, parity unpinned against PyBullet like every
collision distance here: DESIGN.md section 4) sampled densely inside every grid segment, the smallest sample polished by
a bounded scalar minimisation -- and also runs the reference's own SLSQP loop over the same clearance for small cases.
"""
import numpy as np
from scipy.optimize import NonlinearConstraint, minimize, minimize_scalar


def grid(start, goal, resolution, max_distance=None):
    """(T, T_f) of connectors.py:117-124 / :155-167; T is None for an edge shorter than float32 eps."""
    distance = np.linalg.norm(np.asarray(start, np.float64) - np.asarray(goal, np.float64))
    if distance <= np.finfo(np.float32).eps:
        return None, 0.0
    t_f = 1.0
    if max_distance is not None and distance > max_distance:
        t_f = max_distance / distance
    T = np.append(np.arange(0.0, t_f, resolution / distance), t_f)
    return T, t_f


def min_clearance(scene_oracle, start, goal, resolution, max_distance=None, threshold=0.0, per_segment=16, polish=True):
    """Smallest clearance `min_distance - threshold` along the edge: `per_segment` equally spaced float64 samples inside
    every grid segment (ends included), then a bounded Brent search around the smallest one.  Returns (value, t) or
    (None, None) for a degenerate edge."""
    start, goal = np.asarray(start, np.float64), np.asarray(goal, np.float64)
    T, _ = grid(start, goal, resolution, max_distance)
    if T is None:
        return None, None
    ts = np.unique(np.concatenate([np.linspace(a, b, per_segment + 1) for a, b in zip(T[:-1], T[1:])]))
    pts = (1.0 - ts)[:, None] * start[None] + ts[:, None] * goal[None]
    c = scene_oracle.min_distance(pts) - threshold
    i = int(np.argmin(c))
    best, t_best = float(c[i]), float(ts[i])
    if polish:
        lo, hi = ts[max(i - 1, 0)], ts[min(i + 1, len(ts) - 1)]
        if hi > lo:
            f = lambda t: float(scene_oracle.min_distance(((1.0 - t) * start + t * goal)[None])[0] - threshold)
            r = minimize_scalar(f, bounds=(lo, hi), method="bounded", options={"xatol": 1e-7})
            if r.fun < best:
                best, t_best = float(r.fun), float(r.x)
    return best, t_best


def connect_slsqp(checker, start, goal, resolution, max_distance=None):
    """The reference's loop (connectors.py:126-145 / :169-182) over a scalar clearance `checker(q)`: True = the edge is
    returned, False = None."""
    start, goal = np.asarray(start, np.float64), np.asarray(goal, np.float64)
    T, _ = grid(start, goal, resolution, max_distance)
    if T is None:
        return False
    trajectory = lambda t: (1.0 - t) * start + t * goal
    for t0, t1 in zip(T[:-1], T[1:]):
        constraint = NonlinearConstraint(lambda t: checker(trajectory(t[0])), lb=-np.inf, ub=0.0)
        result = minimize(lambda t: t[0], x0=np.array([(t0 + t1) / 2.0]), bounds=[(t0, t1)], method="SLSQP",
                          constraints=[constraint])
        if result.success:
            return False
    return True
