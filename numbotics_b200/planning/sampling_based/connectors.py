"""Edge validity for sampling-based planners, fused on the GPU when the checker is an Arm.

Mirror of reference numbotics/planning/sampling_based/connectors.py:

* `ConnectorParams`  :12-29  (same fields, defaults and validation messages)
* `Connector`        :33-48
* `DiscreteConnector.connect / steer / is_valid`  :51-104

Semantics kept bit-for-bit: `distance <= float32 eps -> None`;
`T = arange(0, T_f, resolution / distance) U {T_f}` in float64; the edge is valid iff
every sample is valid; `connect` returns `copy(goal)`, `steer` returns
`copy(trajectory(T_f))`.

What changes is *where* the samples are evaluated:

* validity checker is an `ArmValidityChecker` (the `lambda q: not arm.in_collision(q)`
  of reference README.md:107 / tests/_test_rrt.py:91-94 made introspectable), default
  linear trajectory and default L2 distance -> ONE launch of the fused edge kernel
  (`nbx_edges_valid_f32`: sample count, interpolation, all-link FK and pair tests on
  the device).  `connect_batch` / `steer_batch` feed it E edges at a time.
* `ArmValidityChecker` with a custom trajectory or distance -> samples are generated
  on the host exactly like the reference, validity is one batched `in_collision` call.
* any other callable -> the reference's sequential loop over the user's callback.

* `ContinuousConnector.connect / steer / is_valid`  :108-185  -- the checker returns a signed clearance
  (<= 0 = invalid) and an edge is rejected when SOME point of a grid segment is invalid.  The reference asks one
  SLSQP solve per segment for such a point (`minimize t s.t. checker(trajectory(t)) <= 0`, success -> None).
  With an `ArmClearanceChecker`, the default trajectory and the default distance the whole edge set is decided on the
  GPU instead (`Arm.edges_clear`: signed minimum distances at the grid samples of all edges in one launch, segments
  certified free by the edge's bound on the clearance's rate of change, the others halved level by level);
  any other callable runs the reference's SLSQP loop statement for statement.
"""
from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import Callable

import numpy as np

from ..trajectories import unit_bspline, LinearTrajectory


def _default_trajectory(x, y):
    return unit_bspline(np.array([x, y]))


def _l2(x, y):
    return np.linalg.norm(x - y)


class ArmValidityChecker:
    """`q -> not arm.in_collision(q, threshold)`; accepts one configuration or a batch."""

    def __init__(self, arm, threshold: float = 0.0):
        self.arm = arm
        self.threshold = float(threshold)

    def __call__(self, q):
        hit = self.arm.in_collision(q, self.threshold)
        if isinstance(hit, bool):
            return not hit
        return ~hit


class ArmClearanceChecker:
    """`q -> arm.min_distance(q) - threshold`: the signed clearance a `ContinuousConnector` expects from its
    validity checker (reference connectors.py:131-134 constrains it to `<= 0`, `is_valid` is `> 0`, :184-185);
    negative values are penetration depths.  Accepts one configuration (-> float) or a batch."""

    def __init__(self, arm, threshold: float = 0.0):
        self.arm = arm
        self.threshold = float(threshold)

    def __call__(self, q):
        d = self.arm.min_distance(q) - self.threshold
        return float(d) if getattr(d, "ndim", 1) == 0 else d


@dataclass(frozen=True)
class ConnectorParams:
    resolution: float = 5e-2
    max_distance: float = 1.0
    trajectory_func: Callable = _default_trajectory
    validity_checker: Callable | None = None

    def __post_init__(self):
        if self.resolution <= 0:
            raise ValueError("Resolution must be positive")
        if self.resolution < 0.0 or self.resolution >= 1.0:
            raise ValueError("Resolution must be strictly between 0.0 and 1.0")
        if self.max_distance <= 0:
            raise ValueError("Max distance must be positive")
        if self.validity_checker is None:
            raise ValueError("Validity checker must be provided")
        if self.trajectory_func is None:
            raise ValueError("Trajectory conversion function must be provided")


class Connector(ABC):

    @abstractmethod
    def connect(self, start: np.ndarray, goal: np.ndarray):
        raise NotImplementedError

    @abstractmethod
    def steer(self, start: np.ndarray, goal: np.ndarray):
        raise NotImplementedError

    @abstractmethod
    def is_valid(self, state: np.ndarray):
        raise NotImplementedError


def sample_times(distance: float, resolution: float, t_f: float = 1.0) -> np.ndarray:
    """The reference's sample grid (connectors.py:68-69 / 92-93)."""
    T = np.arange(0.0, t_f, resolution / distance)
    return np.append(T, t_f)


class DiscreteConnector(Connector):

    def __init__(self, params: ConnectorParams):
        self._params = params

    # -- helpers -----------------------------------------------------------------------------
    def _fused(self, distance_func) -> bool:
        p = self._params
        return (isinstance(p.validity_checker, ArmValidityChecker)
                and p.trajectory_func is _default_trajectory and distance_func is _l2)

    def _edge(self, start, goal, distance_func, steer: bool):
        p = self._params
        distance = distance_func(start, goal)
        if distance <= np.finfo(np.float32).eps:
            return None
        t_f = 1.0
        if steer and distance > p.max_distance:
            t_f = p.max_distance / distance
        trajectory = p.trajectory_func(start, goal)
        chk = p.validity_checker
        if self._fused(distance_func):
            ok = chk.arm.edges_valid(np.asarray(start)[None], np.asarray(goal)[None], p.resolution,
                                     p.max_distance if steer else None, chk.threshold)
            if not bool(ok[0]):
                return None
        else:
            T = sample_times(distance, p.resolution, t_f)
            if isinstance(chk, ArmValidityChecker):
                pts = np.stack([np.asarray(trajectory(t), dtype=np.float64) for t in T])
                if not bool(np.all(chk(pts))):
                    return None
            else:
                for t in T:
                    if not self.is_valid(trajectory(t)):
                        return None
        return np.copy(goal) if not steer else np.copy(trajectory(t_f))

    # -- reference surface -------------------------------------------------------------------
    def connect(self, start: np.ndarray, goal: np.ndarray, distance_func: Callable = _l2):
        return self._edge(start, goal, distance_func, steer=False)

    def steer(self, start: np.ndarray, goal: np.ndarray, distance_func: Callable = _l2):
        return self._edge(start, goal, distance_func, steer=True)

    def is_valid(self, state: np.ndarray):
        return self._params.validity_checker(state)

    # -- batched surface (what a roadmap builder should call) -------------------------------------
    def _require_fused(self):
        if not self._fused(_l2):
            raise ValueError("batched edge validation needs an ArmValidityChecker, the default linear "
                             "trajectory and the default L2 distance")
        return self._params.validity_checker

    def connect_batch(self, starts, goals, return_samples: bool = False):
        """valid[e] == (connect(starts[e], goals[e]) is not None), all edges in one launch."""
        chk = self._require_fused()
        return chk.arm.edges_valid(starts, goals, self._params.resolution, None, chk.threshold, return_samples)

    def steer_batch(self, starts, goals, return_samples: bool = False):
        """valid[e] == (steer(starts[e], goals[e]) is not None); also returns the steered end states."""
        chk = self._require_fused()
        out = chk.arm.edges_valid(starts, goals, self._params.resolution, self._params.max_distance,
                                  chk.threshold, return_samples)
        valid = out[0] if return_samples else out
        try:
            import torch
            is_t = isinstance(starts, torch.Tensor)
        except Exception:  # pragma: no cover
            is_t = False
        if is_t:
            dist = (goals - starts).double().norm(dim=-1)
            t_f = torch.clamp(self._params.max_distance / dist, max=1.0).to(starts.dtype)
            ends = starts + t_f[..., None] * (goals - starts)
        else:
            s, g = np.asarray(starts, np.float64), np.asarray(goals, np.float64)
            dist = np.linalg.norm(g - s, axis=-1)
            with np.errstate(divide="ignore"):
                t_f = np.minimum(self._params.max_distance / dist, 1.0)
            ends = (1.0 - t_f)[..., None] * s + t_f[..., None] * g
        return (valid, ends, out[1]) if return_samples else (valid, ends)


class ContinuousConnector(Connector):
    """Reference connectors.py:108-185.  `connect` / `steer` keep the reference's signatures and return values
    (`None`, `copy(goal)`, `copy(trajectory(T_f))`); `connect_batch` / `steer_batch` decide E edges at once."""

    def __init__(self, params: ConnectorParams):
        self._params = params

    def _fused(self, distance_func) -> bool:
        p = self._params
        return (isinstance(p.validity_checker, ArmClearanceChecker)
                and p.trajectory_func is _default_trajectory and distance_func is _l2)

    def _edge(self, start, goal, distance_func, steer: bool):
        p = self._params
        distance = distance_func(start, goal)
        if distance <= np.finfo(np.float32).eps:
            return None
        t_f = 1.0
        if steer and distance > p.max_distance:
            t_f = p.max_distance / distance
        trajectory = p.trajectory_func(start, goal)
        chk = p.validity_checker
        if self._fused(distance_func):
            ok = chk.arm.edges_clear(np.asarray(start)[None], np.asarray(goal)[None], p.resolution,
                                     p.max_distance if steer else None, chk.threshold)
            if not bool(ok[0]):
                return None
        else:
            from scipy.optimize import NonlinearConstraint, minimize
            T = sample_times(distance, p.resolution, t_f)
            for t0, t1 in zip(T[:-1], T[1:]):
                constraint = NonlinearConstraint(lambda t: chk(trajectory(t[0])), lb=-np.inf, ub=0.0)
                result = minimize(lambda t: t[0], x0=np.array([(t0 + t1) / 2.0]), bounds=[(t0, t1)],
                                  method="SLSQP", constraints=[constraint])
                if result.success:            # a point of the segment with clearance <= 0 exists
                    return None
        return np.copy(goal) if not steer else np.copy(trajectory(t_f))

    def connect(self, start: np.ndarray, goal: np.ndarray, distance_func: Callable = _l2):
        return self._edge(start, goal, distance_func, steer=False)

    def steer(self, start: np.ndarray, goal: np.ndarray, distance_func: Callable = _l2):
        return self._edge(start, goal, distance_func, steer=True)

    def is_valid(self, state: np.ndarray):
        return self._params.validity_checker(state) > 0.0

    def _require_fused(self):
        if not self._fused(_l2):
            raise ValueError("batched continuous edge validation needs an ArmClearanceChecker, the default linear "
                             "trajectory and the default L2 distance")
        return self._params.validity_checker

    def connect_batch(self, starts, goals, return_detail: bool = False):
        """valid[e] == (connect(starts[e], goals[e]) is not None) for a solver that never misses a feasible point."""
        chk = self._require_fused()
        return chk.arm.edges_clear(starts, goals, self._params.resolution, None, chk.threshold, return_detail=return_detail)

    def steer_batch(self, starts, goals, return_detail: bool = False):
        chk = self._require_fused()
        return chk.arm.edges_clear(starts, goals, self._params.resolution, self._params.max_distance, chk.threshold,
                                   return_detail=return_detail)
