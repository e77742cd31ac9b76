"""Drop-in shims for the reference's numba kernels, backed by the sm_100a library.

Same names, argument order and array shapes as reference
numbotics/robots/helpers.py:91-92 (`nb_compute_transformation`) and :117-130
(`nb_compute_jacobian`).  The (offsets, axes, joint_types, joint_idxs) sequence is
flattened into a one-frame `nbx_chain` (cached per distinct sequence), the batch
goes through `nbx_fk_jac_f64` / `_f32`, and the result comes back in the caller's
container type (NumPy in -> NumPy out, torch CUDA in -> torch CUDA out).

Reference quirks kept on purpose so results match on identical inputs:

* `nb_compute_transformation` walks `offsets.shape[0]` entries and ignores any
  extra rows of `axes` / `joint_types` / `joint_idxs` (helpers.py:94-96) -- this is
  what `Arm.forward_kinematics` relies on when it passes `offsets[:-1]` (arm.py:397).
* `nb_compute_jacobian` walks `axes.shape[0]` joints and then ends the chain with
  `offsets[-1]` whatever `offsets.shape[0]` is (helpers.py:157-158).  When the caller
  gives no trailing offset (`offsets.shape[0] == axes.shape[0]`) the reference
  overwrites the last joint's frame with the end frame, so that joint's column
  becomes `[0; R_end a]`; the shim reproduces that (see `_compat_last_column`).
* `T_mats` is the reference's scratch for intermediate frames; the fused kernel has
  no such intermediate and leaves it untouched.
* SPHERICAL joints: `NotImplementedError` like helpers.py:144-145.
"""
from collections import OrderedDict

import numpy as np

from .. import _cabi
from ..physics import Constraint

try:
    import torch
except Exception:  # pragma: no cover
    torch = None

_FIXED = Constraint.FIXED.value
_SPHERICAL = Constraint.SPHERICAL.value
_cache: "OrderedDict[bytes, _cabi.ChainHandle]" = OrderedDict()
_CACHE_MAX = 64


def _np(x):
    if torch is not None and isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return np.asarray(x)


def _sequence_chain(offsets, axes, joint_types, joint_idxs, n_entries, tail, dof, device):
    """One-frame chain from the first `n_entries` (offset, axis, type, idx) rows + a tail offset."""
    offsets = _np(offsets).astype(np.float64)
    axes = _np(axes).astype(np.float64).reshape(-1, 3)
    joint_types = _np(joint_types).astype(np.int64).reshape(-1)
    joint_idxs = _np(joint_idxs).astype(np.int64).reshape(-1)
    key = b"|".join([offsets[:n_entries].tobytes(), axes[:n_entries].tobytes(), joint_types[:n_entries].tobytes(),
                     joint_idxs[:n_entries].tobytes(), np.asarray(tail, np.float64).tobytes(),
                     np.array([n_entries, dof, device], np.int64).tobytes()])
    h = _cache.get(key)
    if h is not None:
        _cache.move_to_end(key)
        return h
    run = np.eye(4)
    mov_offset, mov_axis, mov_type, mov_qidx = [], [], [], []
    for i in range(n_entries):
        t = int(joint_types[i])
        if t == _SPHERICAL:
            raise NotImplementedError("Jacobian for spherical joints not implemented")
        if t == _FIXED:
            run = run @ offsets[i]
            continue
        mov_offset.append((run @ offsets[i])[:3, :4].reshape(12))
        mov_axis.append(axes[i])
        mov_type.append(t)
        mov_qidx.append(int(joint_idxs[i]))
        run = np.eye(4)
    n_mov = len(mov_type)
    arrays = {
        "dof": dof,
        "mov_parent": np.arange(-1, n_mov - 1, dtype=np.int32),
        "mov_type": np.asarray(mov_type, np.int32), "mov_qidx": np.asarray(mov_qidx, np.int32),
        "mov_offset": np.asarray(mov_offset, np.float64).reshape(-1),
        "mov_axis": np.asarray(mov_axis, np.float64).reshape(-1),
        "frame_mov": np.array([n_mov - 1], np.int32),
        "frame_tail": (run @ np.asarray(tail, np.float64))[:3, :4].reshape(12),
        "frame_com": np.eye(4)[:3, :4].reshape(12),
        "prim_frame": np.zeros(0, np.int32), "prim_type": np.zeros(0, np.int32),
        "prim_dims": np.zeros(0), "prim_offset": np.zeros(0), "prim_margin": np.zeros(0),
    }
    h = _cabi.ChainHandle(arrays, device)
    _cache[key] = h
    if len(_cache) > _CACHE_MAX:
        _cache.popitem(last=False)
    return h


def _device_of(x):
    if torch is None or not torch.cuda.is_available():
        raise RuntimeError("numbotics_b200 needs a CUDA device: the batched kinematics path has no CPU fallback")
    if isinstance(x, torch.Tensor) and x.is_cuda:
        return x.device.index
    return torch.cuda.current_device()


def _check_float(x, name):
    if isinstance(x, np.ndarray):
        if x.dtype != np.float64:
            # the reference's eager numba signatures accept float64 only
            raise TypeError(f"{name} must be float64 (got {x.dtype})")
        return torch.float64
    if x.dtype not in (torch.float32, torch.float64):
        raise TypeError(f"{name} must be float32 or float64")
    return x.dtype


def _run(h, q, B, dof, tdtype, base, lp, gp, want_T, want_J, device):
    dev = torch.device("cuda", device)
    qd = (q if isinstance(q, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(q))).to(dev, tdtype).contiguous()
    lpd = None if lp is None else (lp if isinstance(lp, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(lp))).to(dev, tdtype).contiguous()
    gpd = None if gp is None else (gp if isinstance(gp, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(gp))).to(dev, tdtype).contiguous()
    lib = _cabi.load()
    keep, base_ptr = _cabi.base_pose_arg(base)
    with torch.cuda.device(device):
        Td = torch.empty((B, 4, 4), dtype=tdtype, device=dev) if want_T else None
        Jd = torch.empty((B, 6, dof), dtype=tdtype, device=dev) if want_J else None
        fn = lib.nbx_fk_jac_f64 if tdtype == torch.float64 else lib.nbx_fk_jac_f32
        _cabi.check(fn(h.ptr, 0, qd.data_ptr(), B, base_ptr, 0,
                       lpd.data_ptr() if lpd is not None else None, 16,
                       gpd.data_ptr() if gpd is not None else None, 16,
                       Td.data_ptr() if want_T else None, Jd.data_ptr() if want_J else None,
                       int(torch.cuda.current_stream(device).cuda_stream)))
    del keep
    return Td, Jd


def _split_base(T, tdtype, device):
    """(B,4,4) initial transforms -> (common 4x4 for the kernel, per-config left factor or None)."""
    Tn = T if isinstance(T, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(T))
    Tn = Tn.to(torch.device("cuda", device), tdtype)
    if bool((Tn == Tn[:1]).all()):
        return Tn[0].double().cpu().numpy(), None
    return np.eye(4), Tn


def nb_compute_transformation(T, offsets, axes, joint_types, joint_idxs, q):
    """T_b <- T_b . prod_i X_i(q_b[idx_i]) over the `offsets.shape[0]` entries (helpers.py:91-113)."""
    tdtype = _check_float(q, "q")
    B, dof = int(q.shape[0]), int(q.shape[1])
    device = _device_of(q)
    h = _sequence_chain(offsets, axes, joint_types, joint_idxs, int(offsets.shape[0]), np.eye(4), dof, device)
    base, left = _split_base(T, tdtype, device)
    Td, _ = _run(h, q, B, dof, tdtype, base, None, None, True, False, device)
    if left is not None:
        Td = left @ Td
    if isinstance(q, torch.Tensor):
        return Td.to(q.device)
    return Td.cpu().numpy()


def nb_compute_jacobian(T, T_mats, com_offset, offsets, axes, joint_types, joint_idxs, q,
                        local_pose=None, global_pose=None, use_com=False):
    """Geometric Jacobian (B,6,dof), rows [v; w] (helpers.py:117-187)."""
    tdtype = _check_float(q, "q")
    B, dof = int(q.shape[0]), int(q.shape[1])
    device = _device_of(q)
    n_joints = int(np.shape(axes)[0])
    off_np = _np(offsets).astype(np.float64)
    tail = off_np[-1]
    if use_com:
        tail = tail @ _np(com_offset).astype(np.float64)
    h = _sequence_chain(off_np, axes, joint_types, joint_idxs, n_joints, tail, dof, device)
    base, left = _split_base(T, tdtype, device)
    if left is not None and (local_pose is not None or global_pose is not None):
        raise NotImplementedError("per-configuration initial transforms together with local/global poses")
    # the reference applies local_pose, then lets global_pose REPLACE the end frame (helpers.py:164-170)
    lp = None if global_pose is not None else local_pose
    compat = off_np.shape[0] == n_joints and n_joints > 0
    Td, Jd = _run(h, q, B, dof, tdtype, base, lp, global_pose, compat, True, device)
    if left is not None:
        R = left[:, :3, :3]
        Jd = torch.cat([R @ Jd[:, :3, :], R @ Jd[:, 3:, :]], dim=1)
        if compat:
            Td = left @ Td
    if compat:
        _compat_last_column(Jd, Td, global_pose, _np(axes)[n_joints - 1], int(_np(joint_types)[n_joints - 1]),
                            int(_np(joint_idxs)[n_joints - 1]), tdtype)
    if isinstance(q, torch.Tensor):
        return Jd.to(q.device)
    return Jd.cpu().numpy()


def _compat_last_column(Jd, Td, global_pose, axis, jtype, idx, tdtype):
    """No trailing offset: reference helpers.py:157-158 stores the END frame in the last joint's
    slot, so that joint's column is computed from the end frame itself: v = w x 0 = 0 and
    w = R_end a (revolute) / v = R_end a (prismatic)."""
    if jtype == _FIXED:
        return
    if global_pose is not None:
        G = global_pose if isinstance(global_pose, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(global_pose))
        R_end = G.to(Td.device, tdtype)[:, :3, :3]
    else:
        R_end = Td[:, :3, :3]
    w = R_end @ torch.as_tensor(axis, dtype=tdtype, device=Td.device)
    if jtype == Constraint.REVOLUTE.value:
        Jd[:, :3, idx] = 0
        Jd[:, 3:, idx] = w
    else:
        Jd[:, :3, idx] = w
        Jd[:, 3:, idx] = 0
