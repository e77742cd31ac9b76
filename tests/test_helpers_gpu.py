"""The nb_compute_* shims (reference helper signatures) vs golden vectors from the reference's raw helpers."""
import numpy as np
import pytest

from numbotics_b200.robots import nb_compute_transformation, nb_compute_jacobian

pytestmark = pytest.mark.gpu
TOL = 1e-12


def test_raw_helper_shims_match_reference(cuda, golden_raw):
    g = golden_raw
    q, T0 = g["q"], g["T0"]
    T = nb_compute_transformation(T0.copy(), np.ascontiguousarray(g["offsets"][:-1]), g["axes"], g["joint_types"], g["joint_idxs"], q)
    assert T.dtype == np.float64 and np.abs(T - g["T"]).max() < TOL
    B = q.shape[0]
    J = nb_compute_jacobian(T0.copy(), np.zeros((B, 8, 4, 4)), np.eye(4), g["offsets"], g["axes"], g["joint_types"],
                            g["joint_idxs"], q, None, None, False)
    assert J.shape == (B, 6, 7) and np.abs(J - g["J"]).max() < TOL
    # interior FIXED entries in the sequence
    Tm = nb_compute_transformation(T0.copy(), np.ascontiguousarray(g["mixed_offsets"][:-1]), g["mixed_axes"], g["mixed_types"],
                                   g["mixed_idxs"], q)
    assert np.abs(Tm - g["T_mixed"]).max() < TOL
    Jm = nb_compute_jacobian(T0.copy(), np.zeros((B, 9, 4, 4)), np.eye(4), g["mixed_offsets"], g["mixed_axes"][:8],
                             g["mixed_types"][:8], g["mixed_idxs"][:8], q, None, None, False)
    assert np.abs(Jm - g["J_mixed"]).max() < TOL


def test_helper_quirk_without_trailing_offset(cuda, golden_kinova):
    """offsets.shape[0] == axes.shape[0]: the reference's end frame overwrites the last joint's slot."""
    g = golden_kinova
    q = g["q"]
    B = q.shape[0]
    for frame in ["bracelet_link", "forearm_link"]:
        off, axes = g[f"{frame}/offsets"], g[f"{frame}/axes"]
        types, idxs = g[f"{frame}/joint_types"], g[f"{frame}/joint_idxs"]
        T0 = np.tile(g["base_pose"], (B, 1, 1))
        J = nb_compute_jacobian(T0, np.zeros((B, off.shape[0], 4, 4)), np.eye(4), off, axes, types, idxs, q, None, None, False)
        assert np.abs(J - g[f"{frame}/J_compat"]).max() < TOL
        T = nb_compute_transformation(T0, np.ascontiguousarray(off[:-1]), axes, types, idxs, q) @ off[-1]
        assert np.abs(T - g[f"{frame}/T_compat"]).max() < TOL


def test_helper_dtype_rules(cuda, golden_raw):
    import torch
    g = golden_raw
    with pytest.raises(TypeError):
        nb_compute_transformation(g["T0"], g["offsets"][:-1], g["axes"], g["joint_types"], g["joint_idxs"], g["q"].astype(np.float32))
    q = torch.as_tensor(g["q"], dtype=torch.float32, device=cuda)
    T0 = torch.as_tensor(g["T0"], dtype=torch.float32, device=cuda)
    T = nb_compute_transformation(T0, g["offsets"][:-1], g["axes"], g["joint_types"], g["joint_idxs"], q)
    assert T.is_cuda and T.dtype == torch.float32
    assert np.abs(T.cpu().numpy() - g["T"]).max() < 2e-5
    types = g["joint_types"].copy(); types[2] = 2
    with pytest.raises(NotImplementedError, match="spherical"):
        nb_compute_jacobian(g["T0"], None, np.eye(4), g["offsets"], g["axes"], types, g["joint_idxs"], g["q"], None, None, False)
