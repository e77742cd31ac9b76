"""Host logic of the IRIS sampling half (numbotics_b200/planning/safe_sets.py) on CPU, with a stand-in subject whose
"collision" is a ball in configuration space: parameter checks, sample counts, hyperplane formula and the zero-order-hold
loop of reference numbotics/planning/safe_sets.py:164-219."""
import numpy as np
import pytest

from numbotics_b200.planning.safe_sets import IrisParams, IrisSolver, hit_and_run


class BallWorld:
    """in_collision(q) = |q - centre| < radius; bisect_to_collision = the reference's loop, vectorised."""
    def __init__(self, centre, radius):
        self.centre, self.radius = np.asarray(centre, float), float(radius)
        self.calls = {"in_collision": 0, "bisect": 0}

    def in_collision(self, q, threshold=0.0):
        self.calls["in_collision"] += 1
        return np.linalg.norm(np.asarray(q) - self.centre, axis=-1) - self.radius < threshold

    def bisect_to_collision(self, inside, outside, num_bisections=15, threshold=1e-6):
        self.calls["bisect"] += 1
        lo = np.broadcast_to(np.asarray(inside, float), outside.shape).copy()
        hi = np.array(outside, float)
        for _ in range(num_bisections):
            mid = (lo + hi) / 2.0
            hit = self.in_collision(mid, threshold)
            hi[hit] = mid[hit]
            lo[~hit] = mid[~hit]
        return hi


def test_params_defaults_and_checks():
    p = IrisParams()
    assert (p.configuration_margin, p.admissible_collisions, p.max_uncertainty, p.max_iters) == (1e-1, 5e-3, 5e-3, 100)
    assert (p.num_particles, p.num_bisections, p.termination_tolerance, p.collision_tolerance, p.tau) == (1000, 15, 1e-3, 1e-6, 0.5)
    for kw, msg in (({"tau": -0.1}, "tau must be between 0 and 1"), ({"max_uncertainty": 2.0}, "max_uncertainty must be between 0 and 1"),
                    ({"max_iters": 0}, "max_iters must be greater than 0"), ({"num_particles": 0}, "num_particles must be greater than 0"),
                    ({"configuration_margin": -1.0}, "configuration_margin must be greater than 0"),
                    ({"hyperplane_method": "np2"}, "hyperplane_method must be 'zoh'")):
        with pytest.raises(ValueError, match=msg):
            IrisParams(**kw)


def test_sample_count_and_hyperplane_formula():
    s = IrisSolver(BallWorld([0, 0], 1.0), IrisParams())
    # reference safe_sets.py:176-182 for (i, k) = (0, 0) and (2, 3)
    for i, k in ((0, 0), (2, 3)):
        delta = 36.0 * 5e-3 / (np.pi ** 4 * (i + 1) ** 2 * (k + 1) ** 2)
        assert s.unadaptive_samples(i, k) == max(int(np.ceil(2.0 * np.log(1.0 / delta) / (5e-3 * 0.25))), 1000)
    C, d, q = np.diag([2.0, 1.0]), np.array([0.5, 0.0]), np.array([1.5, 1.0])
    a, b = s.new_separating_hyperplane(q, C, d)
    g = (C.T @ C) @ (q - d)
    assert np.allclose(a, g / np.linalg.norm(g)) and np.isclose(b, a @ q - 0.1)
    assert a @ d < b                     # the ellipsoid centre stays inside the new half space


def test_hit_and_run_stays_inside():
    A = np.vstack([np.eye(3), -np.eye(3), [[1.0, 1.0, 0.0]]])
    b = np.array([1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 0.5])
    pts = hit_and_run(A, b, np.zeros(3), 500, 0.5, np.random.default_rng(0))
    assert pts.shape == (500, 3) and np.all(pts @ A.T <= b + 1e-12)
    assert pts.std(axis=0).min() > 0.2   # it moves
    with pytest.raises(ValueError, match="keep_ratio must be at least 0.01"):
        hit_and_run(A, b, np.zeros(3), 10, 0.001)


def test_zero_order_hold_loop_carves_the_ball_out():
    world = BallWorld([1.2, 0.3], 0.6)
    params = IrisParams(num_particles=1500, admissible_collisions=0.02, configuration_margin=0.02, max_iters=60)
    solver = IrisSolver(world, params)
    A = np.vstack([np.eye(2), -np.eye(2)])
    b = np.array([2.0, 2.0, 2.0, 2.0])
    d = np.zeros(2)
    A2, b2, log = solver.seperating_hyperplanes(A, b, np.eye(2), d, outer_iters=0, rng=np.random.default_rng(3))
    assert log[-1][1] < (1 - params.tau) * params.admissible_collisions
    assert A2.shape[0] > 4 and np.all(A2 @ d <= b2)
    # one batched collision call per iteration and one batched bisection per non-final iteration
    assert world.calls["in_collision"] >= len(log) and world.calls["bisect"] == len(log) - 1
    # every added hyperplane separates the centre from (most of) the ball: its normal points towards the ball
    to_ball = world.centre / np.linalg.norm(world.centre)
    assert np.all(A2[4:] @ to_ball > 0.3)
    fresh = hit_and_run(A2, b2, d, 4000, 0.5, np.random.default_rng(4))
    assert world.in_collision(fresh, params.collision_tolerance).mean() < 2 * params.admissible_collisions
    with pytest.raises(StopIteration):
        # one iteration can not bring the colliding fraction under the bar: the reference's for-else raises
        IrisSolver(world, IrisParams(num_particles=200, max_iters=1)).seperating_hyperplanes(A, b, np.eye(2), d, rng=np.random.default_rng(5))
