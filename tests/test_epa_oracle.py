"""Pin the oracle's penetration depth (float64 EPA in oracle/nbx_oracle.c) without Bullet: closed forms (point in box, point
in cylinder, sphere-sphere), the separating-axis minimum for box pairs, and a brute-force minimisation of the support
function of the Minkowski difference for the curved pairs.  Signed distance of the rounded-core model below contact:
d = -(mA + mB) - depth(cores)."""
import numpy as np
import pytest
from scipy.optimize import minimize

from test_collision_oracle import S, B, CY, pose, rot, dist


def test_sphere_in_box_depth_is_distance_to_nearest_face():
    he = np.array([0.5, 0.4, 0.3])
    rng = np.random.default_rng(0)
    for _ in range(200):
        R = rot(rng.normal(size=3), rng.uniform(0, np.pi))
        x = rng.uniform(-1, 1, 3) * (he - 0.03) * 0.98      # inside the shrunk core: the nearest boundary is a flat face
        r = rng.uniform(0.01, 0.2)
        want = -np.min(he - np.abs(x)) - r
        for m in (0.0, 0.03):          # sharp box, and a core shrunk by 0.03 swept by a 0.03 sphere (same faces)
            got = dist(S, pose(R @ x + [0.3, -0.2, 0.1]), [r, 0, 0], B, pose((0.3, -0.2, 0.1), R), he, mb=m)
            assert got == pytest.approx(want, abs=1e-8), (x, r, m)


def test_symmetric_degenerate_seeds():
    """Centres coincide / sit on symmetry planes: the GJK simplex degenerates exactly (origin on an edge of it)."""
    assert dist(S, pose(), [0.1, 0, 0], B, pose(), [0.5, 0.4, 0.3]) == pytest.approx(-0.4, abs=1e-6)
    assert dist(B, pose(), [0.5, 0.5, 0.5], B, pose(), [0.5, 0.5, 0.5], sharp=1) == pytest.approx(-1.0, abs=1e-6)
    assert dist(B, pose(), [0.5, 0.5, 0.5], B, pose((0.6, 0, 0)), [0.5, 0.5, 0.5], sharp=1) == pytest.approx(-0.4, abs=1e-6)
    assert dist(CY, pose(), [0.2, 0.5, 0], CY, pose(), [0.2, 0.5, 0]) == pytest.approx(-0.4, abs=1e-5)
    assert dist(S, pose(), [0.3, 0, 0], S, pose((0.4, 0, 0)), [0.2, 0, 0]) == pytest.approx(-0.1, abs=1e-12)


def test_sphere_in_cylinder():
    r, hh = 0.3, 0.5
    rng = np.random.default_rng(1)
    for _ in range(200):
        R = rot(rng.normal(size=3), rng.uniform(0, np.pi))
        rho, phi, z = rng.uniform(0, 0.95 * r), rng.uniform(0, 2 * np.pi), rng.uniform(-0.95 * hh, 0.95 * hh)
        x = np.array([rho * np.cos(phi), rho * np.sin(phi), z])
        rs = rng.uniform(0.01, 0.1)
        want = -min(r - rho, hh - abs(z)) - rs
        got = dist(S, pose(R @ x), [rs, 0, 0], CY, pose((0, 0, 0), R), [r, hh, 0])
        assert got == pytest.approx(want, abs=2e-7), (rho, z)


def _sat_depth(Ra, ca, ha, Rb, cb, hb):
    """Minimum translation distance of two overlapping sharp boxes: the smallest overlap over the 15 SAT axes."""
    axes = [Ra[:, i] for i in range(3)] + [Rb[:, i] for i in range(3)]
    for i in range(3):
        for j in range(3):
            c = np.cross(Ra[:, i], Rb[:, j])
            if np.linalg.norm(c) > 1e-9:
                axes.append(c / np.linalg.norm(c))
    best = np.inf
    for n in axes:
        ra = np.sum(ha * np.abs(Ra.T @ n))
        rb = np.sum(hb * np.abs(Rb.T @ n))
        overlap = ra + rb - abs(n @ (cb - ca))
        best = min(best, overlap)
    return best


def test_box_box_matches_sat():
    rng = np.random.default_rng(2)
    n_checked = 0
    for _ in range(400):
        Ra, Rb = rot(rng.normal(size=3), rng.uniform(0, np.pi)), rot(rng.normal(size=3), rng.uniform(0, np.pi))
        ha, hb = rng.uniform(0.1, 0.5, 3), rng.uniform(0.1, 0.5, 3)
        cb = rng.normal(size=3) * 0.25
        depth = _sat_depth(Ra, np.zeros(3), ha, Rb, cb, hb)
        if depth <= 1e-3:
            continue
        got = dist(B, pose((0, 0, 0), Ra), ha, B, pose(cb, Rb), hb, sharp=1)
        assert got == pytest.approx(-depth, abs=1e-8)
        n_checked += 1
    assert n_checked > 200


def _support(kind, R, c, dims, n):
    if kind == S:
        return c.copy()
    loc = R.T @ n
    if kind == B:
        return c + R @ (np.sign(loc) * dims)
    r, hh = dims[0], dims[1]
    rad = np.hypot(loc[0], loc[1])
    p = np.array([0.0, 0.0, np.sign(loc[2]) * hh])
    if rad > 1e-14:
        p[:2] = r * loc[:2] / rad
    return c + R @ p


def _brute_depth(ka, Ra, ca, da, kb, Rb, cb, db):
    """min over the unit sphere of h_D(n) = n . (s_A(n) - s_B(-n)): dense sampling, then local refinement."""
    def h(n):
        n = n / np.linalg.norm(n)
        return n @ (_support(ka, Ra, ca, da, n) - _support(kb, Rb, cb, db, -n))
    k = np.arange(40000) + 0.5
    phi, th = np.arccos(1 - 2 * k / 40000), np.pi * (1 + 5 ** 0.5) * k
    N = np.stack([np.cos(th) * np.sin(phi), np.sin(th) * np.sin(phi), np.cos(phi)], 1)
    vals = np.array([h(n) for n in N])
    best = np.inf
    for i in np.argsort(vals)[:6]:
        res = minimize(h, N[i], method="Nelder-Mead", options={"xatol": 1e-10, "fatol": 1e-12, "maxiter": 4000})
        best = min(best, res.fun)
    return best


@pytest.mark.parametrize("kinds", [(CY, B), (CY, CY), (B, CY)])
def test_curved_pairs_match_brute_force(kinds):
    ka, kb = kinds
    rng = np.random.default_rng(3 + ka * 7 + kb)
    done = 0
    while done < 8:
        Ra, Rb = rot(rng.normal(size=3), rng.uniform(0, np.pi)), rot(rng.normal(size=3), rng.uniform(0, np.pi))
        da = np.array([0.15, 0.4, 0.0]) if ka == CY else rng.uniform(0.15, 0.4, 3)
        db = np.array([0.25, 0.3, 0.0]) if kb == CY else rng.uniform(0.15, 0.4, 3)
        cb = rng.normal(size=3) * 0.15
        got = dist(ka, pose((0, 0, 0), Ra), da, kb, pose(cb, Rb), db, sharp=1)
        if got > -5e-3:
            continue
        want = -_brute_depth(ka, Ra, np.zeros(3), da, kb, Rb, cb, db)
        assert got == pytest.approx(want, abs=2e-6), (got, want)
        done += 1


def test_signed_distance_is_continuous_across_core_contact():
    """Slide a rounded cylinder into a rounded box: d(t) is the same straight line before and after the cores touch."""
    he, r, hh, mb, mc = [0.5, 0.5, 0.5], 0.1, 0.3, 0.04, 0.01
    xs = np.linspace(0.75, 0.45, 61)
    d = np.array([dist(CY, pose((x, 0.05, 0.02)), [r, hh, 0], B, pose(), he, ma=mc, mb=mb) for x in xs])
    assert np.allclose(d, xs - 0.5 - r, atol=1e-7)          # cylinder side against the +x face
