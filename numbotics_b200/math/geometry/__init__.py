"""Nearest-neighbour index of the reference's `math.geometry` package (the only part on the roadmap path)."""
from .nearest_neighbors import ApproximateNearestNeighborsIndex

__all__ = ["ApproximateNearestNeighborsIndex"]
