"""Float64 NumPy restatement of the IRIS zero-order-hold bisection.  TEST INFRASTRUCTURE ONLY.

Follows reference numbotics/planning/safe_sets.py:124-134 (`IrisSolver.counter_ex_search_bisection`) statement by
statement, vectorised over the samples; the distance query is `oracle.collision.SceneOracle.min_distance` (parity
unpinned against PyBullet, like every collision distance here: DESIGN.md section 4).  Midpoints are rounded to
float32 before the distance query, which is what the device path evaluates.
"""
import numpy as np


def bisect(scene_oracle, inside, outside, num_bisections=15, threshold=1e-6, band=1e-4):
    """Returns (far_end (B,dof) float64, t_lo (B,), t_hi (B,), borderline (B,) bool).

    Bisection drives every sample onto the collision boundary, so late rounds test midpoints whose distance is inside
    the stated contact band, where the fp32 device path may legitimately take the other branch.  `t_lo, t_hi` are the
    segment parameters (inside + t (outside - inside), exact dyadic numbers) of the interval at the FIRST round whose
    midpoint had |min distance - threshold| <= band (the final interval if there was none): every admissible result
    has its far end in (t_lo, t_hi].  `borderline[b]` says whether such a round occurred."""
    outside = np.asarray(outside, np.float64)
    lo = np.broadcast_to(np.asarray(inside, np.float64), outside.shape).copy()
    hi = outside.copy()
    n = outside.shape[0]
    tl, th = np.zeros(n), np.ones(n)
    borderline = np.zeros(n, bool)
    t_lo, t_hi = tl.copy(), th.copy()
    for _ in range(num_bisections):
        mid = (lo + hi) / 2.0
        tm = (tl + th) / 2.0
        d = scene_oracle.min_distance(mid.astype(np.float32).astype(np.float64))
        newly = (np.abs(d - threshold) <= band) & ~borderline
        t_lo[newly], t_hi[newly] = tl[newly], th[newly]
        borderline |= newly
        hit = d < threshold
        hi[hit] = mid[hit]; th[hit] = tm[hit]
        lo[~hit] = mid[~hit]; tl[~hit] = tm[~hit]
    t_lo[~borderline], t_hi[~borderline] = tl[~borderline], th[~borderline]
    return hi, t_lo, t_hi, borderline
