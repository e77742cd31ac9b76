"""Pin the collision oracle's narrow phase (float64 C GJK) without Bullet: closed-form known answers,
Monte-Carlo point sampling, an independent SLSQP solve, and checker-route vs pipeline-route agreement."""
import ctypes as C

import numpy as np
import pytest

from oracle import collision
from numbotics_b200 import workloads

S, B, CY = collision.PRIM_SPHERE, collision.PRIM_BOX, collision.PRIM_CYL


def pose(p=(0, 0, 0), R=None):
    T = np.eye(4)
    T[:3, 3] = p
    if R is not None:
        T[:3, :3] = R
    return T[:3, :4].reshape(12)


def rot(axis, ang):
    axis = np.asarray(axis, float) / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K


def dist(ta, pa, da, tb, pb, db, ma=0.0, mb=0.0, sharp=0):
    """Distance of one pair through oracle_pair_distances (both shapes static)."""
    sp = np.ascontiguousarray(np.stack([pa, pb]))
    types = np.array([ta, tb], np.int32)
    dims = np.ascontiguousarray(np.stack([np.asarray(da, float), np.asarray(db, float)]))
    margins = np.array([ma, mb], float)
    pairs = np.array([[0, 1]], np.int32)
    out = np.empty((1, 1))
    dummy = np.zeros((1, 1, 12))
    rc = collision.lib().oracle_pair_distances(
        C.c_int64(1), C.c_int(0), dummy.ctypes.data_as(C.POINTER(C.c_double)), C.c_int(2),
        sp.ctypes.data_as(C.POINTER(C.c_double)), types.ctypes.data_as(C.POINTER(C.c_int32)),
        dims.ctypes.data_as(C.POINTER(C.c_double)), margins.ctypes.data_as(C.POINTER(C.c_double)), C.c_int(1),
        pairs.ctypes.data_as(C.POINTER(C.c_int32)), np.array([sharp], np.uint8).ctypes.data_as(C.POINTER(C.c_uint8)),
        out.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == 0
    return float(out[0, 0])


def test_sphere_sphere():
    assert dist(S, pose((0, 0, 0)), [0.3, 0, 0], S, pose((1, 0, 0)), [0.2, 0, 0]) == pytest.approx(0.5, abs=1e-12)
    assert dist(S, pose((0, 0, 0)), [0.3, 0, 0], S, pose((0.4, 0, 0)), [0.2, 0, 0]) == pytest.approx(-0.1, abs=1e-12)


def test_sphere_box_face_edge_corner():
    he = [0.5, 0.4, 0.3]
    assert dist(S, pose((1.0, 0.1, 0.1)), [0.1, 0, 0], B, pose(), he) == pytest.approx(0.4, abs=1e-12)          # face
    assert dist(S, pose((1.0, 0.9, 0.0)), [0.1, 0, 0], B, pose(), he) == pytest.approx(np.hypot(0.5, 0.5) - 0.1, abs=1e-12)  # edge
    assert dist(S, pose((1.0, 0.9, 0.9)), [0.1, 0, 0], B, pose(), he) == pytest.approx(np.sqrt(0.25 + 0.25 + 0.36) - 0.1, abs=1e-12)  # corner
    # rotated box, same relative geometry
    R = rot([1, 2, 3], 0.7)
    assert dist(S, pose(R @ np.array([1.0, 0.9, 0.9])), [0.1, 0, 0], B, pose((0, 0, 0), R), he) == pytest.approx(np.sqrt(0.86) - 0.1, abs=1e-12)


def test_box_box_axis_aligned_and_rotated():
    assert dist(B, pose(), [0.5, 0.5, 0.5], B, pose((2.0, 0.2, -0.1)), [0.25, 0.25, 0.25]) == pytest.approx(1.25, abs=1e-12)
    # edge-edge: unit cube vs a cube rotated 45 deg about z whose vertical edge faces +x face
    R = rot([0, 0, 1], np.pi / 4)
    d = dist(B, pose(), [0.5, 0.5, 0.5], B, pose((2.0, 0, 0), R), [0.5, 0.5, 0.5])
    assert d == pytest.approx(2.0 - 0.5 - np.sqrt(0.5), abs=1e-12)
    # overlapping boxes: clamped to "at most zero"
    assert dist(B, pose(), [0.5, 0.5, 0.5], B, pose((0.6, 0, 0)), [0.5, 0.5, 0.5]) <= 0.0


def test_cylinder_known_answers():
    r, hh = 0.2, 0.5
    # parallel cylinders side by side
    assert dist(CY, pose(), [r, hh, 0], CY, pose((1.0, 0, 0.3)), [r, hh, 0]) == pytest.approx(1.0 - 2 * r, abs=1e-10)
    # stacked along the axis
    assert dist(CY, pose(), [r, hh, 0], CY, pose((0.05, 0, 1.5)), [r, hh, 0]) == pytest.approx(0.5, abs=1e-10)
    # perpendicular (second rotated about y): its side faces the first one's side
    R = rot([0, 1, 0], np.pi / 2)
    assert dist(CY, pose(), [r, hh, 0], CY, pose((0, 1.0, 0), R), [r, hh, 0]) == pytest.approx(1.0 - 2 * r, abs=1e-10)
    # cylinder rim vs sphere: closest feature is the rim circle
    d = dist(CY, pose(), [r, hh, 0], S, pose((0.6, 0, 0.9)), [0.05, 0, 0])
    assert d == pytest.approx(np.hypot(0.6 - r, 0.9 - hh) - 0.05, abs=1e-10)
    # cylinder cap vs box face
    assert dist(CY, pose(), [r, hh, 0], B, pose((0.05, -0.05, 1.0)), [0.3, 0.3, 0.1]) == pytest.approx(0.4, abs=1e-10)
    # cylinder side vs box face
    assert dist(CY, pose(), [r, hh, 0], B, pose((0.8, 0.0, 0.1)), [0.3, 0.3, 0.1]) == pytest.approx(0.3, abs=1e-10)


def test_margin_model_is_a_minkowski_sum():
    """Rounded primitive = shrunk core (+) sphere: face distances are unchanged, corner distances grow."""
    he = [0.5, 0.5, 0.5]
    m = 0.04
    assert dist(S, pose((1.0, 0, 0)), [0.1, 0, 0], B, pose(), he, mb=m) == pytest.approx(0.4, abs=1e-12)
    d_sharp = dist(S, pose((1, 1, 1)), [0.1, 0, 0], B, pose(), he, mb=0.0)
    d_round = dist(S, pose((1, 1, 1)), [0.1, 0, 0], B, pose(), he, mb=m)
    assert d_sharp == pytest.approx(np.sqrt(0.75) - 0.1, abs=1e-12)
    assert d_round == pytest.approx(np.sqrt(3 * 0.54 ** 2) - m - 0.1, abs=1e-12)
    assert d_round - d_sharp == pytest.approx((np.sqrt(3) - 1) * m, abs=1e-12)
    # sharp flag ignores the box margins (Bullet's box-box path)
    assert dist(B, pose((2, 2, 2)), he, B, pose(), he, ma=m, mb=m, sharp=1) == pytest.approx(np.sqrt(3.0), abs=1e-12)


def _sample_surface(t, p12, dims, n, rng):
    P = np.asarray(p12).reshape(3, 4)
    R, c = P[:, :3], P[:, 3]
    if t == S:
        v = rng.normal(size=(n, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True)
        return c + dims[0] * v
    if t == B:
        u = rng.uniform(-1, 1, size=(n, 3))
        k = rng.integers(0, 3, n)
        u[np.arange(n), k] = rng.choice([-1.0, 1.0], n)
        return c + (u * dims) @ R.T
    ang = rng.uniform(0, 2 * np.pi, n)
    z = rng.uniform(-1, 1, n)
    rad = np.ones(n)
    cap = rng.random(n) < 0.3
    z[cap] = rng.choice([-1.0, 1.0], cap.sum())
    rad[cap] = np.sqrt(rng.random(cap.sum()))
    loc = np.stack([dims[0] * rad * np.cos(ang), dims[0] * rad * np.sin(ang), dims[1] * z], axis=1)
    return c + loc @ R.T


@pytest.mark.parametrize("seed", range(6))
def test_monte_carlo_upper_bound_is_tight(seed):
    """The minimum over sampled surface point pairs bounds the true distance from above and approaches it."""
    rng = np.random.default_rng(seed)
    shapes = []
    for _ in range(2):
        t = int(rng.integers(0, 3))
        dims = np.array({S: [rng.uniform(0.05, 0.3), 0, 0], B: list(rng.uniform(0.05, 0.4, 3)),
                         CY: [rng.uniform(0.03, 0.2), rng.uniform(0.05, 0.4), 0]}[t])
        shapes.append((t, dims))
    R1, R2 = rot(rng.normal(size=3), rng.uniform(0, np.pi)), rot(rng.normal(size=3), rng.uniform(0, np.pi))
    sep = rng.normal(size=3); sep *= rng.uniform(0.9, 1.6) / np.linalg.norm(sep)
    pa, pb = pose((0, 0, 0), R1), pose(sep, R2)
    d = dist(shapes[0][0], pa, shapes[0][1], shapes[1][0], pb, shapes[1][1])
    assert d > 0
    A = _sample_surface(shapes[0][0], pa, shapes[0][1], 6000, rng)
    Bp = _sample_surface(shapes[1][0], pb, shapes[1][1], 6000, rng)
    mc = np.sqrt(((A[:, None, :] - Bp[None, :, :]) ** 2).sum(-1)).min()
    assert mc >= d - 1e-9
    assert mc - d < 0.05


@pytest.mark.parametrize("seed", range(8))
def test_against_independent_slsqp(seed):
    """min ||x - y|| s.t. x in A, y in B solved by SciPy SLSQP in local coordinates (no GJK involved)."""
    from scipy.optimize import minimize
    rng = np.random.default_rng(100 + seed)
    kinds = [(CY, B), (CY, CY), (B, B), (CY, S)][seed % 4]
    dims = [np.array({S: [rng.uniform(0.05, 0.3), 0, 0], B: list(rng.uniform(0.05, 0.4, 3)),
                      CY: [rng.uniform(0.03, 0.2), rng.uniform(0.05, 0.4), 0]}[t]) for t in kinds]
    Rs = [rot(rng.normal(size=3), rng.uniform(0, np.pi)) for _ in range(2)]
    sep = rng.normal(size=3); sep *= rng.uniform(0.8, 1.5) / np.linalg.norm(sep)
    cs = [np.zeros(3), sep]

    def cons(t, d, off):
        if t == B:
            return [{"type": "ineq", "fun": (lambda z, k=k, s=s: d[k] - s * z[off + k])} for k in range(3) for s in (1, -1)]
        if t == CY:
            return [{"type": "ineq", "fun": lambda z: d[0] ** 2 - z[off] ** 2 - z[off + 1] ** 2},
                    {"type": "ineq", "fun": lambda z: d[1] - z[off + 2]}, {"type": "ineq", "fun": lambda z: d[1] + z[off + 2]}]
        return [{"type": "ineq", "fun": lambda z: d[0] ** 2 - z[off] ** 2 - z[off + 1] ** 2 - z[off + 2] ** 2}]

    def obj(z):
        x = cs[0] + Rs[0] @ z[:3]
        y = cs[1] + Rs[1] @ z[3:]
        return float(((x - y) ** 2).sum())

    best = np.inf
    for _ in range(4):
        r = minimize(obj, rng.normal(scale=0.01, size=6), constraints=cons(kinds[0], dims[0], 0) + cons(kinds[1], dims[1], 3),
                     method="SLSQP", options={"ftol": 1e-15, "maxiter": 500})
        if r.success:
            best = min(best, np.sqrt(max(r.fun, 0.0)))
    d = dist(kinds[0], pose(cs[0], Rs[0]), dims[0], kinds[1], pose(cs[1], Rs[1]), dims[1])
    assert np.isfinite(best)
    assert d == pytest.approx(best, abs=2e-6)


def test_checker_route_matches_c_pipeline():
    """FK from the NumPy oracle + Python pair sets  ==  everything in C on the flattened descriptor."""
    arm, cubes = workloads.prm_scene()
    so = collision.SceneOracle(arm)
    desc = collision.CDesc(arm._arrays, arm.scene_arrays())
    q = workloads.random_configs_np(arm, 600, seed=2)
    d_chk = so.min_distance(q)
    hit, d_c = collision.c_in_collision(desc, q, arm.base_pose, 0.0, want_dist=True)
    assert np.abs(d_chk - d_c).max() < 1e-9
    assert np.array_equal(hit, d_chk < 0.0)
    assert 0 < hit.sum() < len(hit)
    # early-exit verdict path agrees with the full-distance path
    assert np.array_equal(collision.c_in_collision(desc, q, arm.base_pose, 0.0), hit)
    # edges
    s, g = q[:60], q[60:120]
    v_chk, ns_chk, _ = so.edges_valid(s, g, 0.1, return_detail=True)
    v_c, ns_c, nu_c = collision.c_edges_valid(desc, s, g, arm.base_pose, 0.1)
    assert np.array_equal(v_chk, v_c) and np.array_equal(ns_chk, ns_c)
    assert np.all(nu_c <= ns_c) and np.all(nu_c[v_c] == ns_c[v_c])
    v_chk2, ns2, _ = so.edges_valid(s, g, 0.1, max_distance=1.0, return_detail=True)
    v_c2, ns_c2, _ = collision.c_edges_valid(desc, s, g, arm.base_pose, 0.1, max_distance=1.0)
    assert np.array_equal(v_chk2, v_c2) and np.array_equal(ns2, ns_c2)


def test_margin_models_differ_only_near_contact():
    arm_b, _c1 = workloads.prm_scene(margin_model="bullet")
    arm_s, _c2 = workloads.prm_scene(margin_model="sharp")
    q = workloads.random_configs_np(arm_b, 400, seed=9)
    db = collision.SceneOracle(arm_b, "bullet").min_distance(q)
    ds = collision.SceneOracle(arm_s, "sharp").min_distance(q)
    sep = db > 0
    # rounding only ever removes material: distances grow, by at most (sqrt(3)-1) * (m_a + m_b)
    assert np.all(db[sep] >= ds[sep] - 1e-12)
    assert np.all(db[sep] - ds[sep] <= (np.sqrt(3) - 1) * 0.08 + 1e-12)
