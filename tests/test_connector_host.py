"""Host logic of the connector mirror: params validation, the reference's sample grid, and the generic
(user-callback) path, which is a plain Python loop exactly like the reference's (no GPU involved)."""
import numpy as np
import pytest

from numbotics_b200.planning import unit_bspline
from numbotics_b200.planning.trajectories import LinearTrajectory
from numbotics_b200.planning.sampling_based import ConnectorParams, DiscreteConnector, Connector
from numbotics_b200.planning.sampling_based.connectors import sample_times


def test_params_validation_messages():
    ok = lambda q: True
    with pytest.raises(ValueError, match="Resolution must be positive"):
        ConnectorParams(resolution=0.0, validity_checker=ok)
    with pytest.raises(ValueError, match="strictly between 0.0 and 1.0"):
        ConnectorParams(resolution=1.0, validity_checker=ok)
    with pytest.raises(ValueError, match="Max distance must be positive"):
        ConnectorParams(max_distance=0.0, validity_checker=ok)
    with pytest.raises(ValueError, match="Validity checker must be provided"):
        ConnectorParams()
    with pytest.raises(ValueError, match="Trajectory conversion function must be provided"):
        ConnectorParams(validity_checker=ok, trajectory_func=None)
    p = ConnectorParams(validity_checker=ok)
    assert p.resolution == 5e-2 and p.max_distance == 1.0
    with pytest.raises(Exception):
        p.resolution = 0.1          # frozen
    assert issubclass(DiscreteConnector, Connector)


def test_unit_bspline_is_linear_interpolation():
    a, b = np.array([0.0, 1.0, -2.0]), np.array([1.0, 3.0, 5.0])
    tr = unit_bspline(np.array([a, b]))
    assert isinstance(tr, LinearTrajectory)
    from scipy.interpolate import BSpline
    ref = BSpline(np.array([0.0, 0.0, 1.0, 1.0]), np.array([a, b]), 1)     # what the reference builds
    for t in [0.0, 0.1, 0.5, 0.999, 1.0]:
        assert np.allclose(tr(t), ref(t), atol=1e-15)
    assert np.allclose(tr(np.array([0.0, 0.25, 1.0])), ref(np.array([0.0, 0.25, 1.0])))
    with pytest.raises(ValueError, match="2D array"):
        unit_bspline(a)
    with pytest.raises(ValueError, match="Degree must be less"):
        unit_bspline(np.array([a, b]), degree=2)
    quad = unit_bspline(np.array([a, b, a]), degree=2)
    assert np.allclose(quad(0.0), a) and np.allclose(quad(1.0), a)


def test_sample_grid_is_numpy_arange():
    for dist, res in [(0.5, 0.1), (0.30000000000000004, 0.1), (1.234, 0.1), (0.05, 0.1), (2.0, 0.05)]:
        T = sample_times(dist, res)
        assert T[0] == 0.0 and T[-1] == 1.0 and np.all(np.diff(T[:-1]) > 0)
        assert len(T) == int(np.ceil(1.0 / (res / dist))) + 1
    T = sample_times(4.0, 0.1, 0.25)
    assert T[-1] == 0.25 and len(T) == int(np.ceil(0.25 / (0.1 / 4.0))) + 1


def test_generic_callback_path_is_sequential_with_early_exit():
    calls = []

    def checker(q):
        calls.append(np.array(q))
        return q[0] < 0.55

    conn = DiscreteConnector(ConnectorParams(resolution=0.1, max_distance=0.35, validity_checker=checker))
    a, b = np.zeros(2), np.array([1.0, 0.0])
    assert conn.connect(a, b) is None
    assert len(calls) == 7 and np.isclose(calls[-1][0], 0.6)       # stopped at the first invalid sample
    calls.clear()
    out = conn.connect(a, np.array([0.5, 0.0]))
    assert np.array_equal(out, [0.5, 0.0]) and len(calls) == 6
    assert conn.connect(a, a + 1e-9) is None                      # distance <= float32 eps
    st = conn.steer(a, b)
    assert np.allclose(st, [0.35, 0.0])
    assert bool(conn.is_valid(np.array([0.1, 0.0]))) is True
    # custom distance function is honoured (reference passes space.distance)
    calls.clear()
    conn.connect(a, np.array([0.5, 0.0]), distance_func=lambda x, y: 2.0 * np.linalg.norm(x - y))
    assert len(calls) == 11
    with pytest.raises(ValueError, match="ArmValidityChecker"):
        conn.connect_batch(np.zeros((2, 2)), np.ones((2, 2)))


def test_prm_star_radius_is_the_reference_expression():
    """`PRMStar.connection_radius` (reference planners/prm_star.py:20-26), scalar and vectorised; ln 1 = 0 gives radius 0."""
    import scipy.special
    from numbotics_b200.planning.sampling_based.roadmap import prm_star_radius
    for dim, n, V in ((7, 302, (2 * np.pi) ** 7), (3, 10, 8.0), (2, 2, 1.0)):
        d, k = float(dim), float(n)
        v_ball = (np.pi ** (d / 2.0)) / scipy.special.gamma((d / 2.0) + 1.0)
        gamma = 2.0 * (1.0 + (1.0 / d)) ** (1.0 / d) * (v_ball / V) ** (1.0 / d)
        want = gamma * (np.log(k) / k) ** (1 / d)
        assert abs(prm_star_radius(n, dim, V) - want) <= 1e-15 * max(want, 1.0)
    r = prm_star_radius(np.arange(1, 6), 7, 1.0)
    assert r.shape == (5,) and r[0] == 0.0 and np.all(np.diff(r[2:]) < 0) and r[1] > 0      # ln n / n peaks at n = e
