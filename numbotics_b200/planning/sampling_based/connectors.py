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

`ContinuousConnector` (scalar SLSQP over a signed distance) is out of scope
(SURVEY.md section 2 row 8).
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
