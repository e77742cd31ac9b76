"""Batched IRIS sampling half: `Arm.bisect_to_collision` and the zero-order-hold
separating-hyperplane step against the float64 oracle restating reference numbotics/planning/safe_sets.py:124-219."""
import numpy as np
import pytest

from numbotics_b200 import workloads
from numbotics_b200.planning.safe_sets import IrisParams, IrisSolver, hit_and_run
from oracle import collision, iris

pytestmark = pytest.mark.gpu

EPS = 1e-4   # stated contact band (metres)


@pytest.fixture(scope="module")
def scene(cuda):
    arm, cubes = workloads.prm_scene(device=0)
    arm._test_keepalive = cubes          # obstacles leave the world when they are garbage collected
    return arm, collision.SceneOracle(arm)


def _colliding_samples(arm, so, n, seed):
    q = workloads.random_configs_np(arm, 40 * n, seed=seed)
    d = so.min_distance(q.astype(np.float32).astype(np.float64))
    q = q[d < -5e-3][:n]            # clearly colliding far ends
    assert q.shape[0] == n
    return q


@pytest.mark.parametrize("rounds", [6, 15])
def test_bisection_matches_oracle(scene, rounds):
    arm, so = scene
    tol = 1e-6
    centre = np.zeros(7)
    assert so.min_distance(centre[None])[0] > 0.01
    S_col = _colliding_samples(arm, so, 600, seed=3)
    ref, t_lo, t_hi, borderline = iris.bisect(so, centre, S_col, rounds, tol, band=EPS)
    out = arm.bisect_to_collision(centre, S_col, rounds, tol)
    assert out.dtype == np.float64 and out.shape == S_col.shape
    # samples whose every tested midpoint was outside the contact band: bit for bit (float64 interval arithmetic)
    clear = ~borderline
    if rounds == 6:
        assert clear.mean() > 0.8
    bad = np.where(clear & (np.abs(out - ref).max(axis=1) > 0))[0]
    assert bad.size == 0, (bad[:5], out[bad[:2]], ref[bad[:2]], np.isnan(out).sum(), np.isnan(ref).sum())
    # all samples: the far end is a dyadic point of the segment inside the oracle's interval at its first in-band round
    seg = S_col - centre
    t = (out - centre)[np.arange(len(out)), np.abs(seg).argmax(axis=1)] / seg[np.arange(len(out)), np.abs(seg).argmax(axis=1)]
    k = t * 2 ** rounds
    assert np.abs(k - np.round(k)).max() < 1e-6 and (np.round(k) >= 1).all()
    assert (t > t_lo - 1e-9).all() and (t <= t_hi + 1e-9).all()
    assert np.abs(out - (centre + t[:, None] * seg)).max() < 1e-12
    # and it does what it is for: the far end collides, the near end is free (both within the band)
    d_far = so.min_distance(out.astype(np.float32).astype(np.float64))
    near = out - seg / 2 ** rounds
    d_near = so.min_distance(near.astype(np.float32).astype(np.float64))
    assert (d_far < tol + EPS).all()
    assert (d_near > tol - EPS).all()


def test_bisection_dtypes_shapes_and_per_sample_inside(scene, cuda):
    import torch
    arm, so = scene
    S_col = _colliding_samples(arm, so, 64, seed=5)
    inside = np.zeros_like(S_col)
    inside[:, 0] = np.linspace(-0.2, 0.2, 64)
    a = arm.bisect_to_collision(inside, S_col, 10, 1e-6)
    b = np.stack([arm.bisect_to_collision(inside[i], S_col[i:i + 1], 10, 1e-6)[0] for i in range(0, 64, 8)])
    assert np.array_equal(a[::8], b)
    # float32 in -> float32 out; torch CUDA in -> torch CUDA out; batch dims kept
    f = arm.bisect_to_collision(np.zeros(7, np.float32), S_col.astype(np.float32), 10, 1e-6)
    assert f.dtype == np.float32 and np.abs(f - arm.bisect_to_collision(np.zeros(7), S_col, 10, 1e-6)).max() < 1e-3
    t = arm.bisect_to_collision(torch.zeros(7, device=cuda), torch.as_tensor(S_col, dtype=torch.float32, device=cuda).reshape(8, 8, 7), 10, 1e-6)
    assert t.is_cuda and t.shape == (8, 8, 7) and torch.equal(t.reshape(64, 7).cpu(), torch.from_numpy(f))
    # one bisection = the midpoint or the far end
    one = arm.bisect_to_collision(np.zeros(7), S_col, 1, 1e-6)
    mid = S_col / 2.0
    assert all(np.array_equal(o, m) or np.array_equal(o, s) for o, m, s in zip(one, mid, S_col))
    with pytest.raises(ValueError, match="num_bisections must be greater than 0"):
        arm.bisect_to_collision(np.zeros(7), S_col, 0)
    with pytest.raises(ValueError):
        arm.bisect_to_collision(np.zeros(6), S_col)
    assert arm.bisect_to_collision(np.zeros(7), np.zeros((0, 7))).shape == (0, 7)


def test_iris_params_mirror_reference_checks():
    p = IrisParams()
    assert (p.num_particles, p.num_bisections, p.tau, p.collision_tolerance) == (1000, 15, 0.5, 1e-6)
    for kw, msg in (({"tau": 1.5}, "tau must be between 0 and 1"), ({"num_bisections": 0}, "num_bisections must be greater than 0"),
                    ({"admissible_collisions": -0.1}, "admissible_collisions must be between 0 and 1"),
                    ({"termination_tolerance": 0.0}, "termination_tolerance must be greater than 0")):
        with pytest.raises(ValueError, match=msg):
            IrisParams(**kw)


def test_separating_hyperplanes_zoh(scene):
    """One `seperating_hyperplanes` call on the PRM scene: the region shrinks until the colliding fraction of fresh
    samples is below (1 - tau) * admissible_collisions, every added hyperplane keeps the centre inside."""
    arm, so = scene
    lo, hi = workloads.sample_limits(arm.joint_limits)
    A = np.vstack([np.eye(7), -np.eye(7)])
    b = np.concatenate([hi, -lo])
    d = np.zeros(7)
    C = np.eye(7)
    params = IrisParams(num_particles=2000, admissible_collisions=0.05, max_iters=40)
    solver = IrisSolver(arm, params)
    rng = np.random.default_rng(0)
    A2, b2, log = solver.seperating_hyperplanes(A, b, C, d, outer_iters=0, rng=rng)
    assert log[-1][1] < (1.0 - params.tau) * params.admissible_collisions
    assert A2.shape[0] > A.shape[0] and np.all(A2 @ d <= b2)
    assert np.array_equal(A2[:14], A) and np.array_equal(b2[:14], b)
    assert np.allclose(np.linalg.norm(A2[14:], axis=1), 1.0)
    # fresh samples of the final region, judged by the ORACLE: the colliding fraction is small
    pts = hit_and_run(A2, b2, d, 3000, 0.5, np.random.default_rng(1))
    frac = float((so.min_distance(pts.astype(np.float32).astype(np.float64)) < params.collision_tolerance).mean())
    assert frac < 2.0 * params.admissible_collisions
