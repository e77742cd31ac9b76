"""ctypes binding of include/nbx.h -- the only door from Python into the sm_100a kernels.

There is deliberately no fallback: if the shared library is missing or a call fails,
this module raises.  (The float64 NumPy/C restatement under `oracle/` is test
infrastructure and is never imported from here.)
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libnbx.so")

NBX_JOINT_REVOLUTE, NBX_JOINT_PRISMATIC, NBX_JOINT_FIXED = 0, 1, 4
NBX_PRIM_SPHERE, NBX_PRIM_BOX, NBX_PRIM_CYLINDER = 0, 1, 2
NBX_MAX_PATH, NBX_MAX_PRIMS, NBX_MAX_OBS, NBX_MAX_PAIRS = 64, 64, 64, 1024

_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)


class NbxError(RuntimeError):
    """A call into libnbx failed (message = nbx_last_error())."""


class ChainDesc(C.Structure):
    _fields_ = [
        ("dof", C.c_int32), ("n_mov", C.c_int32),
        ("mov_parent", _i32p), ("mov_type", _i32p), ("mov_qidx", _i32p),
        ("mov_offset", _f64p), ("mov_axis", _f64p),
        ("n_frames", C.c_int32),
        ("frame_mov", _i32p), ("frame_tail", _f64p), ("frame_com", _f64p),
        ("n_prims", C.c_int32),
        ("prim_frame", _i32p), ("prim_type", _i32p),
        ("prim_dims", _f64p), ("prim_offset", _f64p), ("prim_margin", _f64p),
    ]


class SceneDesc(C.Structure):
    _fields_ = [
        ("n_obs", C.c_int32),
        ("obs_type", _i32p), ("obs_dims", _f64p), ("obs_pose", _f64p), ("obs_margin", _f64p),
        ("n_self_pairs", C.c_int32), ("self_pairs", _i32p),
        ("n_env_pairs", C.c_int32), ("env_pairs", _i32p),
        ("self_sharp", _u8p), ("env_sharp", _u8p),
    ]


# name -> (restype, argtypes); the "does every symbol of nbx.h exist" test iterates this table
_vp = C.c_void_p
_i64 = C.c_int64
SIGNATURES = {
    "nbx_version": (C.c_int, []),
    "nbx_last_error": (C.c_char_p, []),
    "nbx_launch_count": (C.c_int64, []),
    "nbx_chain_create": (C.c_int, [C.POINTER(ChainDesc), C.c_int, C.POINTER(_vp)]),
    "nbx_chain_destroy": (None, [_vp]),
    "nbx_scene_create": (C.c_int, [_vp, C.POINTER(SceneDesc), C.POINTER(_vp)]),
    "nbx_scene_destroy": (None, [_vp]),
    "nbx_fk_jac_f32": (C.c_int, [_vp, C.c_int, _vp, _i64, _f64p, C.c_int, _vp, _i64, _vp, _i64, _vp, _vp, _vp]),
    "nbx_fk_jac_f64": (C.c_int, [_vp, C.c_int, _vp, _i64, _f64p, C.c_int, _vp, _i64, _vp, _i64, _vp, _vp, _vp]),
    "nbx_fk_jac_host_f32": (C.c_int, [_vp, C.c_int, _vp, _i64, _f64p, C.c_int, _vp, _vp]),
    "nbx_fk_jac_host_f64": (C.c_int, [_vp, C.c_int, _vp, _i64, _f64p, C.c_int, _vp, _vp]),
    "nbx_in_collision_f32": (C.c_int, [_vp, _vp, _vp, _i64, _f64p, C.c_float, _vp, _vp, _vp]),
    "nbx_in_collision_host_f32": (C.c_int, [_vp, _vp, _vp, _i64, _f64p, C.c_float, _vp]),
    "nbx_edges_valid_f32": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _f64p, C.c_double, C.c_double, C.c_float, _vp, _vp, _vp]),
    "nbx_edges_valid_detail_f32": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _f64p, C.c_double, C.c_double, C.c_float, _vp, _vp, _vp, _vp]),
    "nbx_edges_valid_host_f32": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _f64p, C.c_double, C.c_double, C.c_float, _vp]),
    "nbx_closest_f32": (C.c_int, [_vp, _vp, _vp, _i64, _f64p, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "nbx_pair_distances_f32": (C.c_int, [_vp, _vp, _vp, _i64, _f64p, _vp, _vp, _vp, _vp, _vp, _vp]),
    "nbx_pair_distances_grad_f32": (C.c_int, [_vp, _vp, _vp, _i64, _f64p, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "nbx_knn_f32": (C.c_int, [_vp, _i64, _vp, _i64, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int, _vp]),
    "nbx_bisect_f32": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _i64, _f64p, C.c_int, C.c_float, _vp, _vp]),
    "nbx_bisect_f64": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _i64, _f64p, C.c_int, C.c_float, _vp, _vp]),
    "nbx_ik_dls_f64": (C.c_int, [_vp, C.c_int, _vp, _i64, _vp, _i64, _f64p, C.c_int, C.c_int, _f64p, _f64p, C.c_double,
                                 C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "nbx_ik_dls_f32": (C.c_int, [_vp, C.c_int, _vp, _i64, _vp, _i64, _f64p, C.c_int, C.c_int, _f64p, _f64p, C.c_double,
                                 C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
}

_lib = None


def load():
    """Load libnbx.so (once).  Raises if it has not been built -- there is no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NbxError(
            f"{LIB_PATH} not found: the sm_100a extension has not been built. "
            "Run `python -m numbotics_b200.build` (or __graft_entry__.build()). "
            "numbotics_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError here = header and library disagree
        fn.restype = res
        fn.argtypes = args
    if lib.nbx_version() != 1:
        raise NbxError(f"libnbx version {lib.nbx_version()} does not match this binding (1)")
    _lib = lib
    return lib


def last_error() -> str:
    return load().nbx_last_error().decode("utf-8", "replace")


def check(rc: int):
    if rc != 0:
        raise NbxError(f"libnbx error {rc}: {last_error()}")


def launch_count() -> int:
    return int(load().nbx_launch_count())


def _as(arr, dtype):
    a = np.ascontiguousarray(arr, dtype=dtype)
    return a


def _ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def base_pose_arg(T):
    """4x4 float64 -> ctypes pointer (keeps the array alive through the returned tuple)."""
    a = np.ascontiguousarray(T, dtype=np.float64)
    if a.shape != (4, 4):
        raise ValueError("base_pose must be a 4x4 matrix")
    return a, _ptr(a, C.c_double)


class ChainHandle:
    """Owns an nbx_chain*.  `arrays` is the dict produced by robots.descriptor.flatten_chain."""

    def __init__(self, arrays: dict, device: int):
        lib = load()
        self._keep = {}
        d = ChainDesc()
        d.dof = int(arrays["dof"])
        d.n_mov = len(arrays["mov_type"])
        d.n_frames = len(arrays["frame_mov"])
        d.n_prims = len(arrays["prim_type"])
        for key, ct, dt in (
            ("mov_parent", C.c_int32, np.int32), ("mov_type", C.c_int32, np.int32), ("mov_qidx", C.c_int32, np.int32),
            ("mov_offset", C.c_double, np.float64), ("mov_axis", C.c_double, np.float64),
            ("frame_mov", C.c_int32, np.int32), ("frame_tail", C.c_double, np.float64), ("frame_com", C.c_double, np.float64),
            ("prim_frame", C.c_int32, np.int32), ("prim_type", C.c_int32, np.int32),
            ("prim_dims", C.c_double, np.float64), ("prim_offset", C.c_double, np.float64), ("prim_margin", C.c_double, np.float64),
        ):
            a = _as(arrays[key], dt)
            if a.size == 0:
                a = np.zeros(1, dtype=dt)
            self._keep[key] = a
            setattr(d, key, _ptr(a, ct))
        h = C.c_void_p()
        check(lib.nbx_chain_create(C.byref(d), int(device), C.byref(h)))
        self.ptr = h
        self.device = int(device)
        self.dof = d.dof

    def close(self):
        if getattr(self, "ptr", None):
            load().nbx_chain_destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SceneHandle:
    def __init__(self, chain: ChainHandle, arrays: dict):
        lib = load()
        self._chain = chain   # keep the chain alive as long as the scene
        self._keep = {}
        d = SceneDesc()
        d.n_obs = len(arrays["obs_type"])
        d.n_self_pairs = len(arrays["self_pairs"])
        d.n_env_pairs = len(arrays["env_pairs"])
        for key, ct, dt in (
            ("obs_type", C.c_int32, np.int32), ("obs_dims", C.c_double, np.float64),
            ("obs_pose", C.c_double, np.float64), ("obs_margin", C.c_double, np.float64),
            ("self_pairs", C.c_int32, np.int32), ("env_pairs", C.c_int32, np.int32),
            ("self_sharp", C.c_uint8, np.uint8), ("env_sharp", C.c_uint8, np.uint8),
        ):
            a = _as(arrays[key], dt)
            if a.size == 0:
                a = np.zeros(2, dtype=dt)
            self._keep[key] = a
            setattr(d, key, _ptr(a, ct))
        h = C.c_void_p()
        check(lib.nbx_scene_create(chain.ptr, C.byref(d), C.byref(h)))
        self.ptr = h

    def close(self):
        if getattr(self, "ptr", None):
            load().nbx_scene_destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
