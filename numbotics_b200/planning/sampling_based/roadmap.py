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


def knn_earlier(nodes, k: int, chunk: int = 4096):
    """For every vertex i > 0 its (up to) k nearest vertices among 0..i-1 (the reference's incremental kNN:
    `ApproximateNearestNeighborsIndex.k_nearest` before `add_point`, planners/prm.py:35-40), exact L2, one kernel launch
    (`nbx_knn_f32`, causal mode).  Returns (idx (N,k) int64 sorted by distance, valid (N,k) bool).  `chunk` is ignored
    (kept for the signature of the earlier torch implementation)."""
    from ... import _cabi
    if not nodes.is_cuda:
        raise ValueError("knn_earlier needs a CUDA tensor: numbotics_b200 has no CPU path")
    x = nodes.to(torch.float32).contiguous()
    N, dof = x.shape
    idx = torch.empty((N, k), dtype=torch.long, device=x.device)
    ok = torch.empty((N, k), dtype=torch.uint8, device=x.device)
    lib = _cabi.load()
    with torch.cuda.device(x.device):
        _cabi.check(lib.nbx_knn_f32(x.data_ptr(), N, x.data_ptr(), N, dof, int(k), 1, idx.data_ptr(), ok.data_ptr(), None,
                                    x.device.index, torch.cuda.current_stream(x.device).cuda_stream))
    return idx, ok.to(torch.bool)


class BatchedPRM:
    """Roadmap over an Arm's configuration space with DiscreteConnector edge semantics."""

    def __init__(self, connector: DiscreteConnector, lower_bounds, upper_bounds, k_nearest: int = 10):
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
        idx, ok = knn_earlier(nodes, self.k)
        # every (vertex, neighbour slot) goes through the edge kernel -- the few empty slots of the first k vertices are
        # masked afterwards -- so that the only host round trip of the build is the final compaction of the edge list
        u = idx.reshape(-1)
        v = torch.arange(N, device=dev).repeat_interleave(self.k)
        okf = ok.reshape(-1)
        starts = nodes[u]                   # connect(neighbor.state, new_node.state)
        ends = nodes[v]
        valid = self.connector.connect_batch(starts, ends) & okf
        edges = torch.stack([u[valid], v[valid]], dim=1)
        weights = (nodes[edges[:, 0]] - nodes[edges[:, 1]]).norm(dim=1)
        n_candidates = int(min(N, self.k + 1) * (min(N, self.k + 1) - 1) // 2 + max(0, N - self.k - 1) * self.k)
        return Roadmap(nodes=nodes, edges=edges, weights=weights, candidates=n_candidates, node_valid=node_valid)
