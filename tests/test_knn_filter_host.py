"""The norm filter of the roadmap kNN (csrc/nbx_knn.cu) restated in NumPy float32, operation for operation except that the
products and sums are rounded separately instead of fused (more rounding, so a stricter check): a reference whose exact
float32 squared distance is below the bound must pass the filter, whatever the bound -- the filter only ever adds
candidates.  The tightest case is a bound one ulp above the distance itself."""
import numpy as np
import pytest

F = np.float32


def _filter_passes(a, b, worst):
    """a, b (n, 7) float32, worst (n,) float32 -> bool (n,): `sa + sb > thr` of the kernel."""
    n = a.shape[0]
    one = F(1.0)
    n2 = np.zeros(n, F)
    for d in range(7):
        n2 = (b[:, d] * b[:, d] + n2).astype(F)
    slot = (F(-0.5) * (one - F(1.6e-5)) * n2).astype(F)                  # knn_pack_kernel
    half_a2 = np.zeros(n, F)
    for d in range(7):
        half_a2 = (a[:, d] * a[:, d] + half_a2).astype(F)
    half_a2 = (half_a2 * F(0.5)).astype(F)
    fv = np.concatenate([a, np.ones((n, 1), F)], axis=1)                 # the query with 1 in the spare slot
    tv = np.concatenate([b, slot[:, None]], axis=1)
    sa, sb = np.zeros(n, F), np.zeros(n, F)
    for d in range(0, 8, 4):                                             # two partial sums, the kernel's order
        sa = (fv[:, d] * tv[:, d] + sa).astype(F); sb = (fv[:, d + 1] * tv[:, d + 1] + sb).astype(F)
        sa = (fv[:, d + 2] * tv[:, d + 2] + sa).astype(F); sb = (fv[:, d + 3] * tv[:, d + 3] + sb).astype(F)
    thr = ((half_a2 * (one - F(8e-6))).astype(F) + (F(-0.5) * worst).astype(F)).astype(F) - F(8e-6)
    return (sa + sb).astype(F) > thr


def _exact_d2(a, b):
    da, db = np.zeros(a.shape[0], F), np.zeros(a.shape[0], F)
    for d in range(0, 8, 4):
        e = [(a[:, d + i] - b[:, d + i]).astype(F) if d + i < 7 else np.zeros(a.shape[0], F) for i in range(4)]
        da = (e[0] * e[0] + da).astype(F); db = (e[1] * e[1] + db).astype(F)
        da = (e[2] * e[2] + da).astype(F); db = (e[3] * e[3] + db).astype(F)
    return (da + db).astype(F)


@pytest.mark.parametrize("scale,offset,spread", [(np.pi, 0.0, 1.0), (np.pi, 0.0, 0.01), (1.0, 50.0, 1.0), (1.0, -300.0, 0.05),
                                                 (1000.0, 0.0, 1.0), (1e-3, 0.0, 1.0)])
def test_filter_never_rejects_a_reference_below_the_bound(scale, offset, spread):
    rng = np.random.default_rng(7)
    n = 400_000
    a = (rng.uniform(-scale, scale, (n, 7)) + offset).astype(F)
    b = (a + rng.normal(size=(n, 7)) * scale * spread * rng.uniform(1e-4, 1.0, (n, 1))).astype(F)      # near and far neighbours
    d2 = _exact_d2(a, b)
    worst = np.nextafter(d2, F(np.inf))                                 # d2 < worst by one ulp: the tightest admissible bound
    assert _filter_passes(a, b, worst).all()
    # ... and it does filter: against a bound well below the distance (beyond its own tolerance) it rejects
    tol = 2.0 * 8e-6 * (0.5 * (a.astype(np.float64) ** 2).sum(1) + (b.astype(np.float64) ** 2).sum(1) + 1.0)
    far = d2 > 4.0 * tol + 1e-30
    low = (0.25 * d2).astype(F)
    rejected = ~_filter_passes(a[far], b[far], low[far])
    clearly = 0.75 * d2[far] > 2.0 * tol[far]
    assert rejected[clearly].all()
