"""Small rigid-transform helpers the hot path's callers need to author chains.

Only `trans_mat` / `skew` / `rot_diff` are mirrored (reference
numbotics/math/spatial.py:157-212); the convex-geometry half of the reference's
`math` package is out of scope.
"""
from .spatial import trans_mat, skew, rot_diff, rpy_matrix

__all__ = ["trans_mat", "skew", "rot_diff", "rpy_matrix"]
