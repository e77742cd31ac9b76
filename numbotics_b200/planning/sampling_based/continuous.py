"""The decision procedure behind `Arm.edges_clear` / `ContinuousConnector` (reference connectors.py:108-185), written over
torch tensors of ANY device and a clearance callback, so that the host logic can be tested without a GPU
(tests/test_continuous_host.py) while `Arm.edges_clear` plugs in the fused distance query.

`decide_edges` cuts every edge at the reference's grid `arange(0, T_f, resolution / distance) U {T_f}`
(connectors.py:123-124 / :166-167), evaluates the clearance at all grid samples in ONE callback, and then decides every
segment `[a, b]`: free when `(c_a + c_b - L (b - a)) / 2 > 0` with `L` the edge's bound on |d clearance / dt|, otherwise
halved -- one callback over the midpoints of all open segments of all edges per level -- until it is free, a midpoint with
clearance <= 0 turns up (the edge is invalid) or `max_depth` halvings are reached."""
import numpy as np
import torch


def decide_edges(starts, goals, resolution: float, max_distance, rate_rows, clearance, max_depth: int = 12):
    """starts, goals: (E, dof) float64 tensors on one device; rate_rows: (R, dof) array (descriptor.clearance_rate_rows);
    clearance(edge (n,) int64, t (n,) float64) -> (n,) float64 clearances at the points `start + t (goal - start)` of
    those edges.  Returns (clear (E,) bool, detail dict of tensors / ints)."""
    sd, gd = starts, goals
    dev = sd.device
    E = sd.shape[0]
    delta = gd - sd
    dist = delta.norm(dim=1)
    alive = dist > float(np.finfo(np.float32).eps)              # connectors.py:117-118
    safe = torch.where(alive, dist, torch.ones_like(dist))
    step = float(resolution) / safe
    tf = torch.ones_like(dist)
    if max_distance is not None:
        tf = torch.where(dist > float(max_distance), float(max_distance) / safe, tf)
    n_seg = torch.where(alive, torch.ceil(tf / step), torch.zeros_like(dist)).long()     # numpy.arange length rule
    n_pts = torch.where(alive, n_seg + 1, torch.zeros_like(n_seg))
    total = int(n_pts.sum())
    if total > (1 << 28):
        raise ValueError("too many samples for one call (%d): validate the edges in chunks" % total)
    M = torch.as_tensor(np.asarray(rate_rows), dtype=torch.float64, device=dev)
    rate = (delta.abs() @ M.T).amax(dim=1) if M.shape[0] else torch.zeros_like(dist)

    ids = torch.arange(E, device=dev)
    edge = torch.repeat_interleave(ids, n_pts, output_size=total)
    first = torch.cumsum(n_pts, 0) - n_pts
    k = torch.arange(total, device=dev) - first[edge]
    t = torch.where(k == n_seg[edge], tf[edge], k.double() * step[edge])
    c = clearance(edge, t)
    evals = total
    clear = alive.clone()
    cmin = torch.full((E,), float("inf"), dtype=torch.float64, device=dev)
    cmin.scatter_reduce_(0, edge, c, "amin")
    clear &= ~(cmin <= 0.0)
    # segments: sample i -> i + 1 inside one edge.  One compaction (= one host round trip on a CUDA device) per level.
    ia = torch.nonzero(k < n_seg[edge]).flatten()
    se, ta, tb, ca, cb = edge[ia], t[ia], t[ia + 1], c[ia], c[ia + 1]
    levels, unresolved = 0, 0
    while True:
        keep = torch.nonzero(clear[se] & ~((ca + cb - rate[se] * (tb - ta)) * 0.5 > 0.0)).flatten()
        if keep.numel() == 0:
            break
        if levels == max_depth:
            unresolved = int(keep.numel())
            break
        se, ta, tb, ca, cb = se[keep], ta[keep], tb[keep], ca[keep], cb[keep]
        levels += 1
        tm = 0.5 * (ta + tb)
        cm = clearance(se, tm)
        evals += int(se.numel())
        cmin.scatter_reduce_(0, se, cm, "amin")
        clear &= ~(cmin <= 0.0)
        se = torch.cat([se, se])
        ta, tb = torch.cat([ta, tm]), torch.cat([tm, tb])
        ca, cb = torch.cat([ca, cm]), torch.cat([cm, cb])
    return clear, {"distance_evaluations": evals, "grid_samples": total, "levels": levels,
                   "unresolved_segments": unresolved, "min_clearance": cmin, "rate_bound": rate}
