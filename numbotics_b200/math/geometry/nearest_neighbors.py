"""`ApproximateNearestNeighborsIndex` (reference numbotics/math/geometry/nearest_neighbors.py:6-85) over the exact
brute-force kNN kernel (`nbx_knn_f32`) instead of faiss `IndexFlatL2` -- which is an exact L2 index too, so the answers
are the same sets.

Same method names, arguments and return shapes.  Points live in one device tensor; `remove_points` compacts it (the
reference rebuilds its faiss index the same way).  Additions: `query` accepts `(M, dim)` batches when `flatten=False`
semantics are wanted for every row (`query_batch`).
"""
import numpy as np
import torch

from ... import _cabi


class ApproximateNearestNeighborsIndex:

    def __init__(self, dim: int, device: int = 0):
        self._dim = int(dim)
        self._device = torch.device("cuda", device)
        self._points = {}            # id -> point (host copy, as in the reference)
        self._ids = []               # row -> id
        self._next_id = 0
        self._dev = torch.empty((0, self._dim), dtype=torch.float32, device=self._device)

    def add_points(self, points: np.ndarray):
        points = np.asarray(points)
        if points.ndim == 1:
            points = points.reshape(1, -1)
        if points.shape[1] != self._dim:
            raise ValueError(f"Point dimension {points.shape[1]} does not match index dimension {self._dim}")
        ids = []
        for point in points:
            self._points[self._next_id] = point
            ids.append(self._next_id)
            self._next_id += 1
        self._ids.extend(ids)
        new = torch.as_tensor(points.astype(np.float32), device=self._device)
        self._dev = torch.cat([self._dev, new], dim=0)
        return ids

    def add_point(self, point: np.ndarray):
        if point.ndim != 1:
            raise ValueError(f"Point should be a 1D array, got shape {point.shape}")
        return self.add_points(point.reshape(1, -1))[0]

    def remove_points(self, ids):
        for point_id in ids:
            self._points.pop(point_id, None)
        self._ids = [i for i in self._ids if i in self._points]
        if self._points:
            self._dev = torch.as_tensor(np.vstack([self._points[i] for i in self._ids]).astype(np.float32), device=self._device)
        else:
            self._dev = torch.empty((0, self._dim), dtype=torch.float32, device=self._device)

    def remove_point(self, point_id: int):
        self.remove_points([point_id])

    def query_batch(self, queries: np.ndarray, k: int = 1):
        """(M, dim) queries -> (labels (M, k') int64 ids, squared distances (M, k')), k' = min(k, len(self))."""
        queries = np.asarray(queries)
        if queries.ndim == 1:
            queries = queries.reshape(1, -1)
        if queries.shape[1] != self._dim:
            raise ValueError(f"Query dimension {queries.shape[1]} does not match index dimension {self._dim}")
        kk = min(int(k), len(self._points))
        if kk == 0:
            return np.zeros((queries.shape[0], 0), np.int64), np.zeros((queries.shape[0], 0), np.float32)
        qd = torch.as_tensor(queries.astype(np.float32), device=self._device).contiguous()
        idx = torch.empty((qd.shape[0], kk), dtype=torch.long, device=self._device)
        d2 = torch.empty((qd.shape[0], kk), dtype=torch.float32, device=self._device)
        lib = _cabi.load()
        with torch.cuda.device(self._device):
            _cabi.check(lib.nbx_knn_f32(self._dev.data_ptr(), self._dev.shape[0], qd.data_ptr(), qd.shape[0], self._dim, kk, 0,
                                        idx.data_ptr(), None, d2.data_ptr(), self._device.index,
                                        torch.cuda.current_stream(self._device).cuda_stream))
        rows = idx.cpu().numpy()
        return np.asarray(self._ids, dtype=np.int64)[rows], d2.cpu().numpy()

    def query(self, query: np.ndarray, k: int = 1, return_labels: bool = False, flatten: bool = True):
        labels, _ = self.query_batch(query, k)
        labels = labels[0]
        items = np.array([self._points[idx] for idx in labels])
        if flatten and k == 1:
            items = items[0]
            labels = labels[0]
        if return_labels:
            return items, labels
        return items

    def nearest(self, query: np.ndarray, return_labels: bool = False):
        return self.query(query, k=1, return_labels=return_labels, flatten=True)

    def k_nearest(self, query: np.ndarray, k: int, return_labels: bool = False):
        return self.query(query, k=k, return_labels=return_labels)

    def __len__(self):
        return len(self._points)
