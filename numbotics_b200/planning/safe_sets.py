"""Sampling half of IRIS (reference numbotics/planning/safe_sets.py), batched on the device.

The reference's `IrisSolver.seperating_hyperplanes` (safe_sets.py:164-219) spends its time in two maps over a
thread pool of PyBullet world clones: `in_collision` of every sampled particle (`:186-191`) and, for the
"zoh" method, 15 dependent `collisions()` calls per colliding particle (`counter_ex_search_bisection`,
`:124-134`).  Here both are one batched call each (`Arm.in_collision`, `Arm.bisect_to_collision`); the
hyperplane bookkeeping around them follows the reference line by line on plain `(A, b)` / `(C, d)` arrays.

Out of scope: the convex-geometry package (`Polytope`, `Ellipse`, redundancy removal, the
maximum-volume inscribed ellipsoid -- cvxpy / MOSEK) and the "np2" counter-example search (scalar SLSQP).  The
caller supplies the polytope and the ellipsoid as arrays and owns the outer loop of `IrisSolver.solve`.
"""
from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class IrisParams:
    """Reference safe_sets.py:20-61 (the fields the sampling half uses; same defaults, same checks)."""
    configuration_margin: float = 1e-1
    admissible_collisions: float = 5e-3
    max_uncertainty: float = 5e-3
    max_iters: int = 100
    num_particles: int = 1000
    num_bisections: int = 15
    termination_tolerance: float = 1e-3
    collision_tolerance: float = 1e-6
    tau: float = 0.5
    hyperplane_method: str = "zoh"

    def __post_init__(self):
        # same conditions and messages as the reference's checks
        for name, noun in (("tau", "tau"), ("admissible_collisions", "admissible_collisions"), ("max_uncertainty", "max_uncertainty")):
            v = getattr(self, name)
            if not (0.0 <= v <= 1.0):
                raise ValueError(f"{noun} must be between 0 and 1")
        for name in ("max_iters", "num_particles", "num_bisections"):
            if getattr(self, name) < 1:
                raise ValueError(f"{name} must be greater than 0")
        if self.configuration_margin < 0.0:
            raise ValueError("configuration_margin must be greater than 0")
        if self.termination_tolerance <= 0.0:
            raise ValueError("termination_tolerance must be greater than 0")
        if self.hyperplane_method != "zoh":
            raise ValueError(f"hyperplane_method must be 'zoh' on this path, received {self.hyperplane_method}")


def hit_and_run(A, b, seed, samples: int, keep_ratio: float = 0.5, rng=None):
    """Reference `Polytope.sample` (math/geometry/polytope.py:271-309): a hit-and-run chain started at `seed`,
    shuffled, the first `samples` points kept.  Sequential by construction; it is the caller's side of the path."""
    if keep_ratio < 1e-2:
        raise ValueError("keep_ratio must be at least 0.01, otherwise performance could become extremely poor")
    if keep_ratio > 1.0:
        raise ValueError("keep_ratio must be less than 1")
    rng = np.random.default_rng() if rng is None else rng
    A = np.asarray(A, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    total = int(samples / keep_ratio)
    res = np.zeros((total, A.shape[1]))
    res[0] = seed
    for i in range(1, total):
        x = res[i - 1]
        u = rng.standard_normal(A.shape[1])
        u /= np.linalg.norm(u)
        lams = (b - A @ x) / (A @ u)
        lam = rng.uniform(np.max(lams[lams < 0]), np.min(lams[lams > 0]))
        res[i] = x + lam * u
    rng.shuffle(res)
    return res[:samples]


class IrisSolver:
    """`IrisSolver(subject, params)` of the reference with the polytope `P = {x : A x <= b}` and the ellipsoid
    `E = {C u + d}` passed as arrays."""

    def __init__(self, subject, params: IrisParams = IrisParams()):
        self._subject = subject
        self._params = params

    # reference safe_sets.py:78-81
    def new_separating_hyperplane(self, q, C, d):
        g = (C.T @ C) @ (q - d)
        a = g / np.linalg.norm(g)
        return a, float(a @ q) - self._params.configuration_margin

    def unadaptive_samples(self, outer_iters: int, k: int) -> int:
        """Sample count of iteration (i, k), reference safe_sets.py:176-182."""
        p = self._params
        delta = (36.0 * p.max_uncertainty) / ((np.pi ** 4) * ((outer_iters + 1) ** 2) * ((k + 1) ** 2))
        return max(int(np.ceil(2.0 * np.log(1.0 / delta) / (p.admissible_collisions * (p.tau ** 2)))), p.num_particles)

    # reference safe_sets.py:186-191, one batched call instead of thread_pool.map
    def colliding(self, points):
        return np.asarray(self._subject.in_collision(points, self._params.collision_tolerance), dtype=bool)

    # reference safe_sets.py:124-134 + :198-199, one batched call
    def counter_ex_search_bisection(self, S_col, d):
        return self._subject.bisect_to_collision(d, S_col, self._params.num_bisections, self._params.collision_tolerance)

    def seperating_hyperplanes(self, A, b, C, d, outer_iters: int = 0, sampler=None, rng=None):
        """Reference safe_sets.py:164-219 ("zoh" branch).  Returns the grown `(A, b)` and a per-iteration log
        `(samples, colliding fraction, hyperplanes)`.  `sampler(A, b, M)` defaults to hit-and-run from `d`."""
        p = self._params
        A = np.atleast_2d(np.asarray(A, dtype=np.float64))
        b = np.atleast_1d(np.asarray(b, dtype=np.float64))
        C = np.asarray(C, dtype=np.float64)
        d = np.asarray(d, dtype=np.float64)
        rng = np.random.default_rng() if rng is None else rng
        if sampler is None:
            sampler = lambda A_, b_, M_: hit_and_run(A_, b_, d, M_, 0.5, rng)
        log = []
        for k in range(p.max_iters):
            M = self.unadaptive_samples(outer_iters, k)
            points = sampler(A, b, M)
            hit = self.colliding(points)
            S_col = points[hit]
            frac = float(len(S_col)) / float(M)
            if frac < (1.0 - p.tau) * p.admissible_collisions:
                log.append((M, frac, A.shape[0]))
                break
            S_star = self.counter_ex_search_bisection(S_col, d)
            metric = np.linalg.norm((C @ (S_star - d[None])[..., None]).squeeze(axis=-1), axis=1)
            for q in S_star[np.argsort(metric)]:
                if np.all(A @ q <= b):                                  # P.contains(q)
                    a_new, b_new = self.new_separating_hyperplane(q, C, d)
                    A = np.vstack([A, a_new[None]])
                    b = np.append(b, b_new)
            log.append((M, frac, A.shape[0]))
        else:
            raise StopIteration("Iris solver exceeded max iterations")
        return A, b, log
