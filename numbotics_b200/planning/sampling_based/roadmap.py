"""Batched PRM roadmap construction: the caller of the edge kernel.

The reference builds a roadmap one sample at a time (numbotics/planning/sampling_based/planners/prm.py:19-47):
sample a state, add it as a vertex, ask faiss for its k nearest earlier vertices
(math/geometry/nearest_neighbors.py:6-85), and call `connector.connect(neighbor, new)` k times.  Here the same
roadmap is built inside out: all samples are drawn at once, the k nearest *earlier* vertices of every sample
come from one masked brute-force kNN on the device (plumbing: torch matmul + topk, like the reference's faiss
IndexFlatL2), and all N*k candidate edges go through ONE launch of the fused edge kernel
(`DiscreteConnector.connect_batch`).  Edge set and weights are those the sequential loop would produce for the
same sample sequence (up to kNN ties).
"""
from dataclasses import dataclass

import numpy as np

try:
    import torch
except Exception:  # pragma: no cover
    torch = None

from .connectors import DiscreteConnector, ArmValidityChecker


@dataclass
class Roadmap:
    nodes: "torch.Tensor"        # (N, dof) on the device
    edges: "torch.Tensor"        # (M, 2) int64, edge (earlier vertex, later vertex), undirected
    weights: "torch.Tensor"      # (M,) L2 length
    candidates: int              # candidate edges that were validated
    node_valid: "torch.Tensor"   # (N,) bool: the vertex itself is collision free

    def to_networkx(self):
        import networkx as nx
        G = nx.Graph()
        nodes = self.nodes.cpu().numpy()
        for i, s in enumerate(nodes):
            G.add_node(f"v_{i}", state=s)
        e = self.edges.cpu().numpy()
        w = self.weights.cpu().numpy()
        for (u, v), wt in zip(e, w):
            G.add_edge(f"v_{u}", f"v_{v}", weight=float(wt))
        return G

    def shortest_path(self, src: int, dst: int):
        """Vertex indices of the shortest path src -> dst (None if disconnected); Dijkstra on the host."""
        from scipy.sparse import coo_matrix
        from scipy.sparse.csgraph import dijkstra
        n = self.nodes.shape[0]
        e = self.edges.cpu().numpy()
        w = self.weights.cpu().numpy().astype(np.float64)
        if e.shape[0] == 0:
            return None
        A = coo_matrix((np.concatenate([w, w]), (np.concatenate([e[:, 0], e[:, 1]]), np.concatenate([e[:, 1], e[:, 0]]))), shape=(n, n)).tocsr()
        dist, pred = dijkstra(A, directed=False, indices=src, return_predecessors=True)
        if not np.isfinite(dist[dst]):
            return None
        path = [dst]
        while path[-1] != src:
            path.append(int(pred[path[-1]]))
        return path[::-1]


def knn_earlier(nodes, k: int, chunk: int = 4096, return_dist2: bool = False):
    """For every vertex i > 0 its (up to) k nearest vertices among 0..i-1 (the reference's incremental kNN:
    `ApproximateNearestNeighborsIndex.k_nearest` before `add_point`, planners/prm.py:35-40), exact L2, one C-ABI call
    (`nbx_knn_f32`, causal mode).  Returns (idx (N,k) int64 sorted by distance, valid (N,k) bool) and, with
    `return_dist2`, the squared distances (N,k) float32.  `chunk` is ignored (kept for the signature of the earlier torch
    implementation)."""
    from ... import _cabi
    if not nodes.is_cuda:
        raise ValueError("knn_earlier needs a CUDA tensor: numbotics_b200 has no CPU path")
    x = nodes.to(torch.float32).contiguous()
    N, dof = x.shape
    idx = torch.empty((N, k), dtype=torch.long, device=x.device)
    ok = torch.empty((N, k), dtype=torch.uint8, device=x.device)
    d2 = torch.empty((N, k), dtype=torch.float32, device=x.device) if return_dist2 else None
    lib = _cabi.load()
    with torch.cuda.device(x.device):
        _cabi.check(lib.nbx_knn_f32(x.data_ptr(), N, x.data_ptr(), N, dof, int(k), 1, idx.data_ptr(), ok.data_ptr(),
                                    d2.data_ptr() if return_dist2 else None,
                                    x.device.index, torch.cuda.current_stream(x.device).cuda_stream))
    if return_dist2:
        return idx, ok.to(torch.bool), d2
    return idx, ok.to(torch.bool)


def prm_star_radius(n_nodes, dimension: int, space_volume: float):
    """`PRMStar.connection_radius` (reference planners/prm_star.py:20-26) for a graph of `n_nodes` vertices:
    gamma * (ln n / n)^(1/d),  gamma = 2 (1 + 1/d)^(1/d) (V_ball / V_space)^(1/d).  `n_nodes` may be an array."""
    import math
    d = float(dimension)
    v_ball = math.pi ** (d / 2.0) / math.gamma(d / 2.0 + 1.0)
    gamma = 2.0 * (1.0 + 1.0 / d) ** (1.0 / d) * (v_ball / float(space_volume)) ** (1.0 / d)
    n = np.asarray(n_nodes, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        return gamma * (np.log(n) / n) ** (1.0 / d)


class BatchedPRM:
    """Roadmap over an Arm's configuration space with DiscreteConnector edge semantics."""

    def __init__(self, connector: DiscreteConnector, lower_bounds, upper_bounds, k_nearest: int = 10, star: bool = False,
                 reference_self_slot: bool = False):
        """`star`: PRM* (reference planners/prm_star.py:29-57) -- of the k nearest earlier vertices only those closer than
        `prm_star_radius(number of vertices at insertion, dimension, volume of the box)` are tried.
        `reference_self_slot`: the reference adds the new vertex to its index BEFORE it asks for the k nearest
        (graph.py:53-57, then prm.py:35-37 / prm_star.py:47-50), so one of the k answers is the vertex itself and
        `connect` rejects it (zero distance, connectors.py:64-65): effectively k - 1 earlier neighbours.  True reproduces
        that; the default keeps k real neighbours."""
        chk = connector._params.validity_checker
        if not isinstance(chk, ArmValidityChecker):
            raise ValueError("BatchedPRM needs a DiscreteConnector whose validity checker is an ArmValidityChecker")
        self.connector = connector
        self.arm = chk.arm
        self.lower = np.asarray(lower_bounds, dtype=np.float64)
        self.upper = np.asarray(upper_bounds, dtype=np.float64)
        if self.lower.shape != self.upper.shape or self.lower.ndim != 1:
            raise ValueError("Lower and upper bounds must be 1D arrays of the same shape")
        self.k = int(k_nearest)
        self.star = bool(star)
        self.k_eff = self.k - 1 if reference_self_slot else self.k
        if self.k_eff < 1:
            raise ValueError("k_nearest leaves no neighbour to connect to")

    def build(self, n_samples: int, start=None, goals=(), seed: int = 0, nodes=None) -> Roadmap:
        """Vertices = [start] + goals + uniform samples (or the given `nodes`); edges as described above."""
        dev = torch.device("cuda", self.arm.device_index)
        if nodes is None:
            g = torch.Generator(device=dev)
            g.manual_seed(seed)
            lo = torch.as_tensor(self.lower, dtype=torch.float32, device=dev)
            hi = torch.as_tensor(self.upper, dtype=torch.float32, device=dev)
            samples = lo + torch.rand((n_samples, lo.shape[0]), generator=g, dtype=torch.float32, device=dev) * (hi - lo)
            fixed = [np.asarray(s, dtype=np.float32) for s in ([start] if start is not None else []) + list(goals)]
            if fixed:
                samples = torch.cat([torch.as_tensor(np.stack(fixed), device=dev), samples], dim=0)
            nodes = samples
        else:
            nodes = torch.as_tensor(nodes, dtype=torch.float32, device=dev).contiguous()
        N = nodes.shape[0]
        node_valid = ~self.arm.in_collision(nodes, self.connector._params.validity_checker.threshold)
        k = self.k_eff
        if self.star:
            idx, ok, d2 = knn_earlier(nodes, k, return_dist2=True)
            # vertex i is inserted into a graph that then has i + 1 vertices (prm_star.py:44-47: add_vertex, then the radius)
            radius = prm_star_radius(np.arange(1, N + 1), nodes.shape[1], float(np.prod(self.upper - self.lower)))
            radius[0] = 0.0                                  # ln 1 = 0: the first vertex has no neighbours anyway
            r = torch.as_tensor(radius, dtype=torch.float32, device=dev)
            ok = ok & (d2.sqrt() < r[:, None])               # strict, graph.py:169
        else:
            idx, ok = knn_earlier(nodes, k)
        # every (vertex, neighbour slot) goes through the edge kernel -- the few empty slots of the first k vertices are
        # masked afterwards -- so that the only host round trip of the build is the final compaction of the edge list
        u = idx.reshape(-1)
        v = torch.arange(N, device=dev).repeat_interleave(k)
        okf = ok.reshape(-1)
        starts = nodes[u]                   # connect(neighbor.state, new_node.state)
        ends = nodes[v]
        valid = self.connector.connect_batch(starts, ends) & okf
        edges = torch.stack([u[valid], v[valid]], dim=1)
        weights = (nodes[edges[:, 0]] - nodes[edges[:, 1]]).norm(dim=1)
        n_candidates = int(okf.sum()) if self.star else int(min(N, k + 1) * (min(N, k + 1) - 1) // 2 + max(0, N - k - 1) * k)
        return Roadmap(nodes=nodes, edges=edges, weights=weights, candidates=n_candidates, node_valid=node_valid)
