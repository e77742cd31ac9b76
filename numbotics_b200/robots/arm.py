"""`Arm`: the reference's batched-kinematics / collision call surface over the sm_100a kernels.

Mirror of the hot-path half of reference numbotics/robots/arm.py:

* `__init__` cache build (`_link_joint_sequence`)                  arm.py:17-71
* `forward_kinematics(q, frame, use_com=False, local_pose=None)`   arm.py:369-410
* `jacobian(q, frame, use_com, local_pose, global_pose)`           arm.py:413-461
* `self_collision_pairs` / `collision_pairs` / add / remove        arm.py:190-366
* `in_collision(q, threshold=0.0)`                                 arm.py:603-604

Same names, argument order, defaults, shapes and error texts.  Differences, all
deliberate and documented in DESIGN.md:

* every call is batched: `in_collision` accepts `(..., dof)` and returns a bool
  array (the reference accepts one configuration only, arm.py:555-557);
* dtype rule "same in, same out": float64/float32 NumPy in -> same dtype NumPy out,
  torch tensors in -> torch tensors out on the same device;
* frames whose path ends in a movable joint get the *correct* link frame (the
  reference drops the last joint's rotation there);
* `global_pose=False` (the reference's default, which crashes on `.shape`) is
  treated as `None`;
* there is no CPU fallback: without a CUDA device every compute call raises.
"""
import ctypes as C

from dataclasses import dataclass

import numpy as np
import networkx as nx

from ..physics import GraphChain, Constraint, PhysicsObject, Chain, Link, Shape
from .. import _cabi
from . import descriptor

try:  # torch is plumbing (device memory, streams); the package still imports without it
    import torch
except Exception:  # pragma: no cover
    torch = None


def _is_torch(x):
    return torch is not None and isinstance(x, torch.Tensor)


def _cuda_stream_ptr(device_index: int) -> int:
    return int(torch.cuda.current_stream(device_index).cuda_stream)


@dataclass
class Proximity:
    """The reference's record (numbotics/physics/collision.py:26-32)."""
    subject: object
    target: object
    position_on_subject: np.ndarray
    position_on_target: np.ndarray
    normal_target_to_subject: np.ndarray
    distance: float


@dataclass
class Proximities:
    """Struct-of-arrays form of `Proximity` for a batch of configurations.

    `subject_index` / `target_index` (int32, device-resident for CUDA inputs) index `links` (link primitives in
    descriptor order) and, where `target_is_obstacle`, `obstacles`; `.subject` / `.target` materialise the object
    arrays on demand (a host round trip)."""
    position_on_subject: object
    position_on_target: object
    normal_target_to_subject: object
    distance: object
    valid: object
    subject_index: object
    target_index: object
    target_is_obstacle: object
    links: np.ndarray
    obstacles: np.ndarray
    gradient: object = None

    @staticmethod
    def _host(x):
        return x.cpu().numpy() if _is_torch(x) else np.asarray(x)

    @property
    def subject(self):
        i = self._host(self.subject_index).astype(np.int64)
        return self.links[np.where(i >= 0, i, len(self.links) - 1)]

    @property
    def target(self):
        i = self._host(self.target_index).astype(np.int64)
        env = self._host(self.target_is_obstacle).astype(bool)
        none_l, none_o = len(self.links) - 1, len(self.obstacles) - 1
        return np.where(env, self.obstacles[np.where(env & (i >= 0), i, none_o)], self.links[np.where(~env & (i >= 0), i, none_l)])


class LinkTarget:
    """A link of ANOTHER chain in the world as a collision target (reference robots/arm.py:239-241): the link's own
    collision shape, posed by that chain's current `configuration`.  Quacks like a `PhysicsObject` for the scene flattening."""

    def __init__(self, link, pose):
        self.link = link
        self._collision_shape = link._collision_shape
        self.pose = np.asarray(pose, dtype=np.float64)
        self.name = link.name

    def __hash__(self):
        return hash(self.link)

    def __eq__(self, other):
        return isinstance(other, LinkTarget) and other.link is self.link


def _public(obj):
    """The object a caller knows: the foreign Link behind a LinkTarget, else the object itself."""
    return obj.link if isinstance(obj, LinkTarget) else obj


class Arm:

    def __init__(self, chain: GraphChain, device: int | None = None, margin_model: str = "bullet"):
        if not isinstance(chain, GraphChain):
            raise ValueError("Arm needs a GraphChain (it reads the chain's tree)")
        self._chain = chain
        G = chain._G
        order = list(nx.topological_sort(G))
        root_node = order[0]

        # --- the reference's per-frame cache, reproduced verbatim in structure (arm.py:23-62) ---
        self._link_joint_sequence = {root_node: tuple()}
        self._links_from_nodes = {root_node: G.nodes[root_node]["link"]}
        for node in order[1:]:
            path = nx.shortest_path(G, root_node, node)
            fixed_run, axes, offsets, types, idxs = [], [], [], [], []
            for (u, v) in zip(path[:-1], path[1:]):
                joint = G.edges[(u, v)]["joint"]
                if joint.type == Constraint.FIXED:
                    fixed_run.append(joint)
                    continue
                axes.append(joint.axis)
                types.append(joint.type.value)
                idxs.append(chain._joint_to_index.get(joint))
                T = np.eye(4)
                for fj in fixed_run:
                    T = T @ fj.offset
                fixed_run = []
                offsets.append(T @ joint.offset)
            if fixed_run:
                T = np.eye(4)
                for fj in fixed_run:
                    T = T @ fj.offset
                offsets.append(T)
            self._link_joint_sequence[node] = (
                np.array(offsets).reshape(-1, 4, 4), np.array(axes).reshape(-1, 3),
                np.array(types, dtype=np.int64), np.array(idxs, dtype=np.int64))
            self._links_from_nodes[node] = G.nodes[node]["link"]

        self._additional_self_collision_pairs = set()
        self._void_self_collision_pairs = set()
        self._additional_collision_pairs = set()
        self._void_collision_pairs = set()
        self._pairs_revision = 0
        self.self_collision_pairs()

        # --- device side -----------------------------------------------------------------
        self._margin_model = margin_model
        self._arrays = descriptor.flatten_chain(chain, margin_model)
        self._device = device
        self._chain_handle = None
        self._scenes = {}
        self._scene_arrays = None

    # ------------------------------------------------------------------ properties
    @property
    def base_pose(self):
        return self._chain.base_pose

    @property
    def joint_limits(self):
        return self._chain.joint_limits

    @property
    def dof(self):
        return self._chain.dof

    @property
    def device_index(self) -> int:
        if self._device is None:
            if torch is None or not torch.cuda.is_available():
                raise RuntimeError(
                    "numbotics_b200 needs a CUDA device: the batched kinematics / collision path has "
                    "no CPU fallback")
            self._device = torch.cuda.current_device()
        return self._device

    # ------------------------------------------------------------------ handles
    def _handle(self) -> _cabi.ChainHandle:
        if self._chain_handle is None:
            self._chain_handle = _cabi.ChainHandle(self._arrays, self.device_index)
        return self._chain_handle

    def _scene(self, threshold: float = 0.0, exact_distance: bool = False) -> _cabi.SceneHandle:
        """Device scene for the current world / pair sets.  Verdict queries use a scene whose obstacle pairs are
        statically pruned for thresholds up to `static_thr` (rebuilt when a larger threshold is asked for);
        minimum-distance queries use the unpruned scene."""
        world = self._chain.world
        base = self.base_pose
        which = "dist" if exact_distance else "verdict"
        cur = self._scenes.get(which)
        key = (world.revision, self._pairs_revision, base.tobytes())
        if cur is None or cur["key"] != key or (not exact_distance and threshold > cur["static_thr"]):
            static_thr = None if exact_distance else max(0.05, 2.0 * float(threshold))
            arrays = self.scene_arrays(static_thr)
            cur = {"key": key, "static_thr": static_thr, "arrays": arrays,
                   "handle": _cabi.SceneHandle(self._handle(), arrays)}
            self._scenes[which] = cur
        self._scene_arrays = cur["arrays"]
        return cur["handle"]

    def scene_arrays(self, static_threshold: float | None = None) -> dict:
        """Current obstacles + allowed pairs as nbx_scene_desc arrays (also used by the tests).
        `static_threshold` not None -> obstacle pairs that can never come within it are dropped."""
        self_pairs = self.self_collision_pairs()
        env_pairs = []
        targets = {}
        for a, b in self.collision_pairs().difference(self_pairs):
            if isinstance(b, PhysicsObject):
                env_pairs.append((a, b))
            elif isinstance(b, Link) and b not in self._chain._links and b._collision_shape.shape != Shape.EMPTY:
                if b not in targets:
                    targets[b] = LinkTarget(b, self._foreign_link_pose(b))
                env_pairs.append((a, targets[b]))
        prune = None if static_threshold is None else (self.base_pose, self.joint_limits, static_threshold)
        return descriptor.flatten_scene(self._arrays, self_pairs, env_pairs, self._margin_model, prune)

    # ------------------------------------------------------------------ pair bookkeeping
    def self_collision_pairs(self):
        """Effective rule of reference arm.py:190-223.

        The reference's "at least two non-fixed joints apart" filter never runs: it asks
        networkx for a path between `link.name` strings ("world:body:link") that are not graph
        nodes, `NodeNotFound` is swallowed (arm.py:202-214), and what remains is: every unordered
        pair of non-EMPTY links that are not joined by a direct edge (arm.py:216-217).
        """
        if not hasattr(self, "_self_collision_pairs"):
            self._self_collision_pairs = set()
            Gu = self._chain._G.to_undirected()
            for link_a in self._chain._links:
                if link_a._collision_shape.shape == Shape.EMPTY:
                    continue
                for link_b in self._chain._links:
                    if link_a is link_b or link_b._collision_shape.shape == Shape.EMPTY:
                        continue
                    if not Gu.has_edge(link_a._name, link_b._name) and (link_b, link_a) not in self._self_collision_pairs:
                        self._self_collision_pairs.add((link_a, link_b))
        return self._self_collision_pairs.union(
            self._additional_self_collision_pairs).difference(self._void_self_collision_pairs)

    def _foreign_link_pose(self, link):
        """World pose of a link of another chain at that chain's current configuration -- through that chain's own
        batched-FK path (a lazily built `Arm` on the same device), not a CPU restatement."""
        for obj in self._chain.world.objects():
            if isinstance(obj, Chain) and obj is not self._chain and link in obj._links:
                other = getattr(obj, "_nbx_arm", None)
                if other is None:
                    other = Arm(obj, device=self.device_index, margin_model=self._margin_model)
                    obj._nbx_arm = other
                return np.asarray(other.forward_kinematics(np.asarray(obj.configuration, dtype=np.float64), link._name))
        raise ValueError(f"Object name: {link._name} must be a valid object in the world")

    def collision_pairs(self):
        """reference arm.py:226-250, including the links of the OTHER chains in the world (:239-241)."""
        pairs = set()
        for link in self._chain._links:
            if link._collision_shape.shape == Shape.EMPTY:
                continue
            for obj in self._chain.world.objects():
                if obj is self._chain:
                    continue
                if isinstance(obj, PhysicsObject):
                    if obj._collision_shape.shape == Shape.EMPTY:
                        continue
                    pairs.add((link, obj))
                elif isinstance(obj, Chain):
                    for obj_link in obj._links:
                        pairs.add((link, obj_link))
        return pairs.union(self.self_collision_pairs()).union(
            self._additional_collision_pairs).difference(self._void_collision_pairs)

    def _resolve(self, ref):
        if isinstance(ref, str):
            obj = self._links_from_nodes.get(ref)
            if obj is None:
                obj = self._chain.world.get_object(ref)
                if obj is None:
                    raise ValueError(f"Object name: {ref} must be a valid object in the world")
            return obj
        return ref

    def add_collision_pair(self, link_a, link_b):
        """reference arm.py:253-308."""
        link_a, link_b = self._resolve(link_a), self._resolve(link_b)
        a_in, b_in = link_a in self._chain._links, link_b in self._chain._links
        if not a_in and not b_in:
            return
        if not a_in and b_in:
            link_a, link_b = link_b, link_a
            a_in, b_in = b_in, a_in
        if b_in:
            cur = self.self_collision_pairs()
            if (link_a, link_b) not in cur and (link_b, link_a) not in cur:
                if (link_a, link_b) not in self._additional_self_collision_pairs and \
                        (link_b, link_a) not in self._additional_self_collision_pairs:
                    # the reference files this under `_additional_collision_pairs` (arm.py:290)
                    self._additional_collision_pairs.add((link_a, link_b))
            self._void_self_collision_pairs.discard((link_a, link_b))
            self._void_self_collision_pairs.discard((link_b, link_a))
        else:
            cur = self.collision_pairs()
            if (link_a, link_b) not in cur and (link_b, link_a) not in cur:
                if (link_a, link_b) not in self._additional_collision_pairs and \
                        (link_b, link_a) not in self._additional_collision_pairs:
                    self._additional_collision_pairs.add((link_a, link_b))
            self._void_collision_pairs.discard((link_a, link_b))
            self._void_collision_pairs.discard((link_b, link_a))
        self._pairs_revision += 1

    def remove_collision_pair(self, link_a, link_b):
        """reference arm.py:311-366."""
        link_a, link_b = self._resolve(link_a), self._resolve(link_b)
        a_in, b_in = link_a in self._chain._links, link_b in self._chain._links
        if not a_in and not b_in:
            return
        if not a_in and b_in:
            link_a, link_b = link_b, link_a
            a_in, b_in = b_in, a_in
        if b_in:
            self._additional_self_collision_pairs.discard((link_a, link_b))
            self._additional_self_collision_pairs.discard((link_b, link_a))
            # pairs added through add_collision_pair live in the other set (see above)
            self._additional_collision_pairs.discard((link_a, link_b))
            self._additional_collision_pairs.discard((link_b, link_a))
            if (link_a, link_b) not in self._void_self_collision_pairs and \
                    (link_b, link_a) not in self._void_self_collision_pairs:
                cur = self.self_collision_pairs()
                if (link_a, link_b) in cur:
                    self._void_self_collision_pairs.add((link_a, link_b))
                elif (link_b, link_a) in cur:
                    self._void_self_collision_pairs.add((link_b, link_a))
        else:
            self._additional_collision_pairs.discard((link_a, link_b))
            self._additional_collision_pairs.discard((link_b, link_a))
            if (link_a, link_b) not in self._void_collision_pairs and \
                    (link_b, link_a) not in self._void_collision_pairs:
                cur = self.collision_pairs()
                if (link_a, link_b) in cur:
                    self._void_collision_pairs.add((link_a, link_b))
                elif (link_b, link_a) in cur:
                    self._void_collision_pairs.add((link_b, link_a))
        self._pairs_revision += 1

    # ------------------------------------------------------------------ argument handling
    def _frame_id(self, frame):
        if self._link_joint_sequence.get(frame) is None:
            raise ValueError(f"Frame {frame} not found in chain")
        return self._arrays["frame_id"][frame]

    def _check_q(self, q, what="q"):
        if q.shape[-1] != self.dof:
            raise ValueError(f"{what} must have {self.dof} elements")

    @staticmethod
    def _check_pose_arg(pose, name, q_shape):
        """Shape rules of arm.py:384-392 / 432-447.  Returns (array-like (n,4,4) or (4,4), stride)."""
        if tuple(pose.shape[-2:]) != (4, 4):
            raise ValueError(f"{name} must be a 4x4 matrix")
        if pose.ndim == 2:
            return pose, 0
        if tuple(pose.shape[:-2]) != tuple(q_shape[:-1]):
            raise ValueError(f"{name} must have the same batch dimensions as q")
        return pose.reshape(-1, 4, 4), 16

    def _to_device(self, x, dtype):
        """NumPy / torch (any device) -> contiguous torch tensor on this arm's CUDA device."""
        dev = torch.device("cuda", self.device_index)
        if _is_torch(x):
            return x.to(device=dev, dtype=dtype).contiguous()
        return torch.from_numpy(np.ascontiguousarray(x)).to(device=dev, dtype=dtype)

    def _fkjac(self, q, frame, use_com, local_pose, global_pose, want_T, want_J):
        """Shared driver of forward_kinematics / jacobian / fk_jacobian."""
        fid = self._frame_id(frame)
        self._check_q(q)
        in_torch = _is_torch(q)
        q_shape = tuple(q.shape)
        single = len(q_shape) == 1
        if local_pose is not None and global_pose is not None:
            raise ValueError("local_pose and global_pose cannot both be provided")
        lp = gp = None
        lps = gps = 0
        if local_pose is not None:
            lp, lps = self._check_pose_arg(local_pose, "local_pose", q_shape)
        if global_pose is not None:
            gp, gps = self._check_pose_arg(global_pose, "global_pose", q_shape)

        if in_torch:
            if q.dtype not in (torch.float32, torch.float64):
                raise TypeError("q must be float32 or float64")
            tdtype, out_np = q.dtype, None
        else:
            q = np.asarray(q)
            if q.dtype not in (np.float32, np.float64):
                q = q.astype(np.float64)
            tdtype, out_np = (torch.float32 if q.dtype == np.float32 else torch.float64), q.dtype
        f64 = tdtype == torch.float64
        B = int(np.prod(q_shape[:-1])) if not single else 1
        base_keep, base_ptr = _cabi.base_pose_arg(self.base_pose)
        lib = _cabi.load()
        h = self._handle()
        dof = self.dof

        host_path = (not in_torch or q.device.type == "cpu") and lp is None and gp is None
        if host_path:
            # HOST buffers straight through the C-ABI: chunked H2D / kernel / D2H pipeline in C++
            qn = q.detach().numpy() if in_torch else q
            qn = np.ascontiguousarray(qn.reshape(B, dof))
            if B >= 32768:
                # large results land in page-locked host memory (torch's caching pinned allocator): the D2H copies
                # then run at PCIe rate instead of the ~5 GB/s of pageable first-touch memory.  The ndarray
                # keeps the pinned block alive through its base.
                Tn = torch.empty((B, 4, 4), dtype=tdtype, pin_memory=True).numpy() if want_T else None
                Jn = torch.empty((B, 6, dof), dtype=tdtype, pin_memory=True).numpy() if want_J else None
            else:
                Tn = np.empty((B, 4, 4), qn.dtype) if want_T else None
                Jn = np.empty((B, 6, dof), qn.dtype) if want_J else None
            fn = lib.nbx_fk_jac_host_f64 if f64 else lib.nbx_fk_jac_host_f32
            _cabi.check(fn(h.ptr, fid, qn.ctypes.data, B, base_ptr, int(bool(use_com)),
                           Tn.ctypes.data if want_T else None, Jn.ctypes.data if want_J else None))
            if in_torch:
                Tn = torch.from_numpy(Tn) if want_T else None
                Jn = torch.from_numpy(Jn) if want_J else None
            T_out, J_out = Tn, Jn
        else:
            dev_index = self.device_index
            qd = self._to_device(q, tdtype).reshape(B, dof)
            lpd = self._to_device(lp, tdtype) if lp is not None else None
            gpd = self._to_device(gp, tdtype) if gp is not None else None
            with torch.cuda.device(dev_index):
                Td = torch.empty((B, 4, 4), dtype=tdtype, device=qd.device) if want_T else None
                Jd = torch.empty((B, 6, dof), dtype=tdtype, device=qd.device) if want_J else None
                fn = lib.nbx_fk_jac_f64 if f64 else lib.nbx_fk_jac_f32
                _cabi.check(fn(h.ptr, fid, qd.data_ptr(), B, base_ptr, int(bool(use_com)),
                               lpd.data_ptr() if lpd is not None else None, lps,
                               gpd.data_ptr() if gpd is not None else None, gps,
                               Td.data_ptr() if want_T else None, Jd.data_ptr() if want_J else None,
                               _cuda_stream_ptr(dev_index)))
            if in_torch and q.device.type == "cuda":
                T_out, J_out = Td, Jd
            elif in_torch:
                T_out = Td.cpu() if want_T else None
                J_out = Jd.cpu() if want_J else None
            else:
                T_out = Td.cpu().numpy().astype(out_np, copy=False) if want_T else None
                J_out = Jd.cpu().numpy().astype(out_np, copy=False) if want_J else None
        del base_keep
        if want_T:
            T_out = T_out[0] if single else T_out.reshape(*q_shape[:-1], 4, 4)
        if want_J:
            J_out = J_out[0] if single else J_out.reshape(*q_shape[:-1], 6, dof)
        return T_out, J_out

    # ------------------------------------------------------------------ kinematics
    def forward_kinematics(self, q, frame: str, use_com: bool = False, local_pose=None):
        return self._fkjac(q, frame, use_com, local_pose, None, True, False)[0]

    def jacobian(self, q, frame: str, use_com: bool = False, local_pose=None, global_pose=False):
        if global_pose is False:   # the reference's default value; it cannot be used as a pose
            global_pose = None
        return self._fkjac(q, frame, use_com, local_pose, global_pose, False, True)[1]

    def fk_jacobian(self, q, frame: str, use_com: bool = False, local_pose=None):
        """Pose and Jacobian from ONE pass over q (what the fused kernel natively produces)."""
        return self._fkjac(q, frame, use_com, local_pose, None, True, True)

    def inverse_kinematics(self, pose, q0, frame: str, use_com: bool = False, use_limits: bool = False,
                           tol: float = 1e-6, max_iter: int = 100, max_failures: int = 15, return_iters: bool = False):
        """Batched damped-least-squares IK (reference arm.py:464-552), the whole loop on the device.

        Returns `(converged, q)` like the reference: `converged = |error twist| < tol`; shapes follow the
        reference's batch rules (pose `(...,4,4)` / `(4,4)`, q0 `(...,dof)` / `(dof,)`).
        """
        fid = self._frame_id(frame)
        self._check_q(q0, "q0")
        if max_iter < 1:
            raise ValueError("max_iter must be greater than 0")
        in_torch = _is_torch(q0) or _is_torch(pose)
        pose_shape, q_shape = tuple(pose.shape), tuple(q0.shape)
        if pose_shape[-2:] != (4, 4):
            raise ValueError("pose must be a 4x4 matrix")
        batch_dims = None
        if len(pose_shape) > 2:
            batch_dims = pose_shape[:-2]
        if len(q_shape) > 1:
            if batch_dims is not None:
                if q_shape[:-1] != batch_dims:
                    raise ValueError("pose and q0 must have the same batch dimensions")
            else:
                batch_dims = q_shape[:-1]
        B = int(np.prod(batch_dims)) if batch_dims is not None else 1
        f64 = (q0.dtype in (np.float64,)) if not _is_torch(q0) else (q0.dtype == torch.float64)
        if not _is_torch(q0) and q0.dtype not in (np.float32, np.float64):
            q0 = np.asarray(q0, dtype=np.float64)
            f64 = True
        tdtype = torch.float64 if f64 else torch.float32
        dev_index = self.device_index
        qd = self._to_device(q0, tdtype).reshape(-1, self.dof)
        if qd.shape[0] != B:
            qd = qd.expand(B, self.dof).contiguous()
        pd = self._to_device(pose, tdtype).reshape(-1, 4, 4)
        stride = 16 if pd.shape[0] == B and B > 1 else (0 if pd.shape[0] == 1 else 16)
        lim = np.ascontiguousarray(self.joint_limits, dtype=np.float64)
        lo = np.ascontiguousarray(lim[:, 0]) if lim.size else np.zeros(1)
        hi = np.ascontiguousarray(lim[:, 1]) if lim.size else np.zeros(1)
        base_keep, base_ptr = _cabi.base_pose_arg(self.base_pose)
        lib = _cabi.load()
        with torch.cuda.device(dev_index):
            q_out = torch.empty_like(qd)
            ok = torch.empty((B,), dtype=torch.uint8, device=qd.device)
            iters = torch.empty((B,), dtype=torch.int32, device=qd.device) if return_iters else None
            fn = lib.nbx_ik_dls_f64 if f64 else lib.nbx_ik_dls_f32
            _cabi.check(fn(self._handle().ptr, fid, pd.data_ptr(), stride, qd.data_ptr(), B, base_ptr, int(bool(use_com)),
                           int(bool(use_limits)), lo.ctypes.data_as(C.POINTER(C.c_double)), hi.ctypes.data_as(C.POINTER(C.c_double)),
                           float(tol), int(max_iter), int(max_failures), q_out.data_ptr(), ok.data_ptr(),
                           iters.data_ptr() if return_iters else None, _cuda_stream_ptr(dev_index)))
        del base_keep
        ok = ok.to(torch.bool)
        if batch_dims is None:
            q_res, ok_res = q_out[0], ok           # reference: q[0] and a (1,) bool array (arm.py:550-551)
        else:
            q_res, ok_res = q_out.reshape(*batch_dims, self.dof), ok.reshape(batch_dims)
        it_res = None if iters is None else (iters if batch_dims is None else iters.reshape(batch_dims))
        if not in_torch:
            q_res, ok_res = q_res.cpu().numpy(), ok_res.cpu().numpy()
            it_res = None if it_res is None else it_res.cpu().numpy()
        elif _is_torch(q0) and q0.device.type != "cuda":
            q_res, ok_res = q_res.cpu(), ok_res.cpu()
            it_res = None if it_res is None else it_res.cpu()
        return (ok_res, q_res, it_res) if return_iters else (ok_res, q_res)

    # ------------------------------------------------------------------ collision validity
    def in_collision(self, q, threshold: float = 0.0):
        """`min over allowed pairs of distance < threshold` (reference arm.py:603-604), batched.

        q `(dof,)` -> Python bool (reference behaviour); `(..., dof)` -> bool array / tensor.
        """
        verdict = self._collide(q, threshold, want_dist=False)[0]
        if verdict.ndim == 0:
            return bool(verdict)
        return verdict

    def min_distance(self, q, threshold: float = 0.0):
        """Signed minimum distance over the allowed pairs: negative = penetration depth (EPA), as the reference's
        `closest_to(q).distance` (arm.py:599-604)."""
        return self._collide(q, threshold, want_dist=True)[1]

    def _collide(self, q, threshold, want_dist):
        self._check_q(q)
        in_torch = _is_torch(q)
        q_shape = tuple(q.shape)
        single = len(q_shape) == 1
        B = int(np.prod(q_shape[:-1])) if not single else 1
        scene = self._scene(threshold, exact_distance=want_dist)
        if self._scene_arrays["self_pairs"].shape[0] + self._scene_arrays["n_env_pairs_before_pruning"] == 0:
            # reference: min([]) -> ValueError (arm.py:600)
            raise ValueError("min() arg is an empty sequence")
        dev_index = self.device_index
        base_keep, base_ptr = _cabi.base_pose_arg(self.base_pose)
        qd = self._to_device(q, torch.float32).reshape(B, self.dof)
        lib = _cabi.load()
        with torch.cuda.device(dev_index):
            vd = torch.empty((B,), dtype=torch.uint8, device=qd.device)
            dd = torch.empty((B,), dtype=torch.float32, device=qd.device) if want_dist else None
            _cabi.check(lib.nbx_in_collision_f32(
                self._handle().ptr, scene.ptr, qd.data_ptr(), B, base_ptr, float(threshold),
                vd.data_ptr(), dd.data_ptr() if want_dist else None, _cuda_stream_ptr(dev_index)))
        del base_keep
        verdict = vd.to(torch.bool)
        if in_torch:
            verdict = verdict.to(q.device)
            dist = dd.to(q.device) if want_dist else None
        else:
            verdict = verdict.cpu().numpy()
            dist = dd.cpu().numpy() if want_dist else None
        shape = q_shape[:-1]
        verdict = verdict.reshape(shape) if not single else verdict.reshape(())
        if want_dist:
            dist = dist.reshape(shape) if not single else dist.reshape(())
        return verdict, dist

    # ------------------------------------------------------------------ every allowed pair
    def pair_distances(self, q, points: bool = False, gradient: bool = False):
        """Batched `Arm.collisions` (reference arm.py:555-580): one proximity per allowed pair and configuration.

        Returns `(pairs, distance)` with `pairs` the list of `(subject, target)` objects (self pairs first) and
        `distance (..., P)`; with `points=True` a `Proximities`-like dict of arrays
        `{distance, position_on_subject, position_on_target, normal_target_to_subject, valid}` of shape `(..., P[, 3])`;
        `gradient=True` (implies the dict form) adds `gradient (..., P, dof)`: d distance / d q of every pair, i.e.
        `jacobian_proximity` (reference arm.py:621-632) for all allowed pairs at once.
        """
        points = points or gradient
        self._check_q(q)
        in_torch = _is_torch(q)
        q_shape = tuple(q.shape)
        single = len(q_shape) == 1
        B = int(np.prod(q_shape[:-1])) if not single else 1
        scene = self._scene(0.0, exact_distance=True)
        arrays = self._scene_arrays
        link_of_prim = {p: l for l, p in self._arrays["prim_of_link"].items()}
        obstacles = [_public(o) for o in arrays["obstacles"]]
        pairs = [(link_of_prim[int(a)], link_of_prim[int(b)]) for a, b in arrays["self_pairs"]]
        pairs += [(link_of_prim[int(a)], obstacles[int(o)]) for a, o in arrays["env_pairs"]]
        P = len(pairs)
        if P == 0:
            raise ValueError("min() arg is an empty sequence")
        dev_index = self.device_index
        base_keep, base_ptr = _cabi.base_pose_arg(self.base_pose)
        qd = self._to_device(q, torch.float32).reshape(B, self.dof)
        lib = _cabi.load()
        with torch.cuda.device(dev_index):
            f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=qd.device)
            dist = f(B, P)
            ps, pt, nrm = (f(B, P, 3), f(B, P, 3), f(B, P, 3)) if points else (None, None, None)
            ok = torch.empty((B, P), dtype=torch.uint8, device=qd.device) if points else None
            grad = f(B, P, self.dof) if gradient else None
            if gradient:
                _cabi.check(lib.nbx_pair_distances_grad_f32(
                    self._handle().ptr, scene.ptr, qd.data_ptr(), B, base_ptr, dist.data_ptr(), ps.data_ptr(), pt.data_ptr(),
                    nrm.data_ptr(), ok.data_ptr(), grad.data_ptr(), _cuda_stream_ptr(dev_index)))
            else:
                _cabi.check(lib.nbx_pair_distances_f32(
                    self._handle().ptr, scene.ptr, qd.data_ptr(), B, base_ptr, dist.data_ptr(),
                    ps.data_ptr() if points else None, pt.data_ptr() if points else None, nrm.data_ptr() if points else None,
                    ok.data_ptr() if points else None, _cuda_stream_ptr(dev_index)))
        del base_keep
        if in_torch:
            conv = lambda x: x.to(q.device)
        else:
            out_dt = q.dtype if getattr(q, "dtype", None) in (np.float32, np.float64) else np.float64
            conv = lambda x: x.cpu().numpy().astype(out_dt, copy=False) if x.dtype == torch.float32 else x.cpu().numpy()
        shape = q_shape[:-1]
        dist_o = conv(dist).reshape(*shape, P)
        if not points:
            return pairs, dist_o
        rec = {"distance": dist_o, "position_on_subject": conv(ps).reshape(*shape, P, 3),
               "position_on_target": conv(pt).reshape(*shape, P, 3),
               "normal_target_to_subject": conv(nrm).reshape(*shape, P, 3), "valid": conv(ok.to(torch.bool)).reshape(*shape, P)}
        if gradient:
            rec["gradient"] = conv(grad).reshape(*shape, P, self.dof)
        return pairs, rec

    def collisions(self, q):
        """reference arm.py:555-580 for one configuration: the list of `Proximity` records of every allowed pair."""
        if tuple(q.shape) != (self.dof,):
            raise ValueError(f"q must be a 1D array with {self.dof} elements")
        pairs, rec = self.pair_distances(q, points=True)
        host = lambda x: x.cpu().numpy() if _is_torch(x) else x
        d, ps, pt, n = (host(rec[k]) for k in ("distance", "position_on_subject", "position_on_target", "normal_target_to_subject"))
        return [Proximity(subject=s_, target=t_, position_on_subject=ps[i], position_on_target=pt[i],
                          normal_target_to_subject=n[i], distance=float(d[i])) for i, (s_, t_) in enumerate(pairs)]

    def self_collisions(self, q):
        """reference arm.py:583-596: the proximities of the allowed link-link pairs, one configuration."""
        links = set(self._arrays["prim_of_link"].keys())
        return [p for p in self.collisions(q) if p.target in links]

    @staticmethod
    def _involves(subject, target, obj):
        """`obj` as the reference's distance_to / jacobian_proximity accept it: a PhysicsObject, a Link, or a whole Chain."""
        if isinstance(obj, Chain):
            return target in obj._links or subject in obj._links
        return target is obj or subject is obj

    def jacobian_proximity(self, q, obj, link=None):
        """`Arm.jacobian_proximity(q, obj, link)` (reference arm.py:621-632), batched: for every allowed pair between the arm
        (or its one `link`) and `obj` the row `n . J_v(subject, p_subject) - n . J_v(target, p_target)` = d distance / d q.

        `q (dof,)` -> `(n_pairs, dof)`, or `(dof,)` when there is exactly one pair (reference :630-632); `q (..., dof)` ->
        `(..., n_pairs, dof)` / `(..., dof)`.  Pairs come in the order of `pair_distances`."""
        pairs, rec = self.pair_distances(q, gradient=True)
        cols = [i for i, (s_, t_) in enumerate(pairs) if self._involves(s_, t_, obj)]
        if link is not None:
            cols = [i for i in cols if pairs[i][0] is link or pairs[i][1] is link]
            if not cols:
                raise ValueError(f"Collision pair ({link.name}, {obj.name}) not valid")
        J = rec["gradient"][..., cols, :]
        return J[..., 0, :] if len(cols) == 1 else J

    def distance_to(self, q, obj, link=None):
        """reference arm.py:607-618: the proximities between the arm (or one `link`) and `obj`, one configuration."""
        prox = [p for p in self.collisions(q) if self._involves(p.subject, p.target, obj)]
        if link is None:
            return prox
        out = [p for p in prox if p.subject is link or p.target is link]
        if not out:
            raise ValueError(f"Collision pair ({link.name}, {obj.name}) not valid")
        return out

    # ------------------------------------------------------------------ closest pair, witness points, distance gradient
    def closest_to(self, q, gradient: bool = False):
        """Batched `Arm.closest_to` (reference arm.py:599-600): the allowed pair with the smallest distance.

        `q (dof,)` -> one `Proximity` (the reference's record: subject, target, position_on_subject,
        position_on_target, normal_target_to_subject, distance); `q (..., dof)` -> a `Proximities` record whose
        fields are arrays with the batch dimensions of q (`subject` / `target` are object arrays of Link /
        PhysicsObject, `valid` is False only where no allowed pair exists; overlapping pairs report the signed
        penetration distance, push-out normal and witness points of the EPA).
        `gradient=True` adds `.gradient (..., dof)`: d distance / d q, i.e. `jacobian_proximity` of the closest pair.
        """
        self._check_q(q)
        in_torch = _is_torch(q)
        q_shape = tuple(q.shape)
        single = len(q_shape) == 1
        B = int(np.prod(q_shape[:-1])) if not single else 1
        scene = self._scene(0.0, exact_distance=True)
        arrays = self._scene_arrays
        if arrays["self_pairs"].shape[0] + arrays["n_env_pairs_before_pruning"] == 0:
            raise ValueError("min() arg is an empty sequence")
        dev_index = self.device_index
        base_keep, base_ptr = _cabi.base_pose_arg(self.base_pose)
        qd = self._to_device(q, torch.float32).reshape(B, self.dof)
        lib = _cabi.load()
        with torch.cuda.device(dev_index):
            f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=qd.device)
            dist, ps, pt, nrm = f(B), f(B, 3), f(B, 3), f(B, 3)
            grad = f(B, self.dof) if gradient else None
            sub = torch.empty((B,), dtype=torch.int32, device=qd.device)
            tgt = torch.empty((B,), dtype=torch.int32, device=qd.device)
            ok = torch.empty((B,), dtype=torch.uint8, device=qd.device)
            _cabi.check(lib.nbx_closest_f32(self._handle().ptr, scene.ptr, qd.data_ptr(), B, base_ptr, dist.data_ptr(),
                                            ps.data_ptr(), pt.data_ptr(), nrm.data_ptr(), sub.data_ptr(), tgt.data_ptr(),
                                            ok.data_ptr(), grad.data_ptr() if gradient else None, _cuda_stream_ptr(dev_index)))
        del base_keep
        # subject / target as indices (kept on the device for CUDA inputs) + lookup tables; object arrays on demand
        prim_of_link = self._arrays["prim_of_link"]
        n_prims = max(prim_of_link.values()) + 1 if prim_of_link else 0
        obstacles = [_public(o) for o in arrays["obstacles"]]
        link_tab = np.empty(n_prims + 1, dtype=object)            # last entry: None (no pair)
        for l, p in prim_of_link.items():
            link_tab[p] = l
        obs_tab = np.empty(len(obstacles) + 1, dtype=object)
        for i, o in enumerate(obstacles):
            obs_tab[i] = o
        is_env = (tgt >= 0) & ((tgt & 0x40000000) != 0)
        tgt_idx = torch.where(tgt >= 0, tgt & 0x3fffffff, tgt)
        if in_torch:
            conv = lambda x: x.to(q.device)
        else:
            out_dt = q.dtype if getattr(q, "dtype", None) in (np.float32, np.float64) else np.float64
            conv = lambda x: x.cpu().numpy().astype(out_dt, copy=False) if x.dtype == torch.float32 else x.cpu().numpy()
        shape = q_shape[:-1]
        rec = Proximities(
            position_on_subject=conv(ps).reshape(*shape, 3), position_on_target=conv(pt).reshape(*shape, 3),
            normal_target_to_subject=conv(nrm).reshape(*shape, 3), distance=conv(dist).reshape(shape),
            valid=conv(ok.to(torch.bool)).reshape(shape), subject_index=conv(sub).reshape(shape),
            target_index=conv(tgt_idx).reshape(shape), target_is_obstacle=conv(is_env).reshape(shape),
            links=link_tab, obstacles=obs_tab,
            gradient=conv(grad).reshape(*shape, self.dof) if gradient else None)
        if single and not gradient:
            si, ti, env = int(sub[0]), int(tgt_idx[0]), bool(is_env[0])
            return Proximity(subject=link_tab[si if si >= 0 else n_prims],
                             target=(obs_tab if env else link_tab)[ti if ti >= 0 else -1],
                             position_on_subject=rec.position_on_subject, position_on_target=rec.position_on_target,
                             normal_target_to_subject=rec.normal_target_to_subject, distance=float(rec.distance))
        return rec

    def jacobian_proximity_closest(self, q):
        """`Arm.jacobian_proximity` (reference arm.py:621-632) for the closest pair of each configuration:
        `n . J_v(subject, p_subject) - n . J_v(target, p_target)`, shape `(..., dof)`."""
        return self.closest_to(q, gradient=True).gradient

    # ------------------------------------------------------------------ bisection to the collision boundary
    def bisect_to_collision(self, inside, outside, num_bisections: int = 15, threshold: float = 1e-6):
        """Batched `IrisSolver.counter_ex_search_bisection` (reference safe_sets.py:124-134).

        `inside` (dof,) (shared, e.g. the ellipsoid centre) or (..., dof); `outside` (..., dof), colliding samples.
        Per sample the interval [inside, outside] is halved `num_bisections` times, keeping the half whose far end
        has `min distance < threshold`; returns the far (colliding) ends, same type / dtype / shape as `outside`.
        """
        self._check_q(outside, "outside")
        self._check_q(inside, "inside")
        if num_bisections < 1:
            raise ValueError("num_bisections must be greater than 0")
        in_torch = _is_torch(outside)
        o_shape = tuple(outside.shape)
        i_shape = tuple(inside.shape)
        if len(i_shape) > 1 and i_shape != o_shape:
            raise ValueError("inside must be one configuration or have the same shape as outside")
        B = int(np.prod(o_shape[:-1])) if len(o_shape) > 1 else 1
        if in_torch:
            if outside.dtype not in (torch.float32, torch.float64):
                raise TypeError("outside must be float32 or float64")
            tdtype, out_np = outside.dtype, None
        else:
            outside = np.asarray(outside)
            if outside.dtype not in (np.float32, np.float64):
                outside = outside.astype(np.float64)
            tdtype, out_np = (torch.float32 if outside.dtype == np.float32 else torch.float64), outside.dtype
        scene = self._scene(threshold)
        if self._scene_arrays["self_pairs"].shape[0] + self._scene_arrays["n_env_pairs_before_pruning"] == 0:
            raise ValueError("min() arg is an empty sequence")
        dev_index = self.device_index
        base_keep, base_ptr = _cabi.base_pose_arg(self.base_pose)
        od = self._to_device(outside, tdtype).reshape(B, self.dof)
        idv = self._to_device(inside if _is_torch(inside) else np.asarray(inside), tdtype)
        idv = idv.reshape(B, self.dof) if len(i_shape) > 1 else idv.reshape(self.dof)
        lib = _cabi.load()
        fn = lib.nbx_bisect_f64 if tdtype == torch.float64 else lib.nbx_bisect_f32
        with torch.cuda.device(dev_index):
            out = torch.empty_like(od)
            _cabi.check(fn(self._handle().ptr, scene.ptr, idv.data_ptr(), self.dof if len(i_shape) > 1 else 0,
                           od.data_ptr(), B, base_ptr, int(num_bisections), float(threshold), out.data_ptr(),
                           _cuda_stream_ptr(dev_index)))
        del base_keep
        out = out.reshape(o_shape)
        if in_torch:
            return out.to(outside.device)
        return out.cpu().numpy().astype(out_np, copy=False)

    # ------------------------------------------------------------------ edge validation
    def edges_valid(self, starts, goals, resolution: float, max_distance: float | None = None,
                    threshold: float = 0.0, return_samples: bool = False, return_useful: bool = False):
        """Batched DiscreteConnector verdicts for edges start[e] -> goal[e] (see planning.sampling_based).

        `max_distance=None` = `connect` semantics, otherwise `steer` semantics.
        `return_samples` adds the number of interpolation samples per edge; `return_useful` adds the number of samples
        the sequential reference would have evaluated (up to and including the first invalid one, connectors.py:71-73).
        """
        self._check_q(starts, "start")
        self._check_q(goals, "goal")
        if tuple(starts.shape) != tuple(goals.shape):
            raise ValueError("starts and goals must have the same shape")
        if resolution <= 0:
            raise ValueError("Resolution must be positive")
        in_torch = _is_torch(starts)
        shape = tuple(starts.shape[:-1])
        E = int(np.prod(shape)) if len(shape) else 1
        scene = self._scene(threshold)
        if self._scene_arrays["self_pairs"].shape[0] + self._scene_arrays["n_env_pairs_before_pruning"] == 0:
            raise ValueError("min() arg is an empty sequence")
        dev_index = self.device_index
        base_keep, base_ptr = _cabi.base_pose_arg(self.base_pose)
        sd = self._to_device(starts, torch.float32).reshape(E, self.dof)
        gd = self._to_device(goals, torch.float32).reshape(E, self.dof)
        lib = _cabi.load()
        with torch.cuda.device(dev_index):
            vd = torch.empty((E,), dtype=torch.uint8, device=sd.device)
            want_n = return_samples or return_useful
            nd = torch.empty((E,), dtype=torch.int32, device=sd.device) if want_n else None
            fd = torch.empty((E,), dtype=torch.int32, device=sd.device) if return_useful else None
            if return_useful:
                _cabi.check(lib.nbx_edges_valid_detail_f32(
                    self._handle().ptr, scene.ptr, sd.data_ptr(), gd.data_ptr(), E, base_ptr,
                    float(resolution), float(max_distance) if max_distance is not None else 0.0,
                    float(threshold), vd.data_ptr(), nd.data_ptr(), fd.data_ptr(), _cuda_stream_ptr(dev_index)))
            else:
                _cabi.check(lib.nbx_edges_valid_f32(
                    self._handle().ptr, scene.ptr, sd.data_ptr(), gd.data_ptr(), E, base_ptr,
                    float(resolution), float(max_distance) if max_distance is not None else 0.0,
                    float(threshold), vd.data_ptr(), nd.data_ptr() if want_n else None,
                    _cuda_stream_ptr(dev_index)))
        del base_keep
        valid = vd.to(torch.bool)
        useful = torch.where(valid, nd, torch.clamp(fd, max=0x7ffffffe) + 1) if return_useful else None
        if in_torch:
            conv = lambda x: x.to(starts.device).reshape(shape)
        else:
            conv = lambda x: x.cpu().numpy().reshape(shape)
        out = [conv(valid)]
        if return_samples:
            out.append(conv(nd))
        if return_useful:
            out.append(conv(useful))
        return tuple(out) if len(out) > 1 else out[0]

    # ------------------------------------------------------------------ continuous edge validation
    def edges_clear(self, starts, goals, resolution: float, max_distance: float | None = None,
                    threshold: float = 0.0, max_depth: int = 12, return_detail: bool = False):
        """Batched `ContinuousConnector` verdicts (reference connectors.py:108-185): `clear[e]` is False iff SOME point of
        the straight motion `start[e] -> goal[e]` (up to `T_f` under steer semantics) has clearance
        `min_distance(q) - threshold <= 0` -- the feasibility problem the reference hands to one SLSQP solve per segment of
        the sample grid `arange(0, T_f, resolution / distance) U {T_f}`.

        The reference's solver is a local search from each segment's midpoint; here every segment is DECIDED: the signed
        minimum distance (fused distance query, `nbx_in_collision_f32` with `min_dist`) is evaluated at the grid samples of
        all edges in one launch, and a segment `[a, b]` with positive clearances `c_a, c_b` at its ends is free when
        `(c_a + c_b - L (b - a)) / 2 > 0`, `L = max_r sum_c M[r, c] |goal_c - start_c|` being the edge's bound on
        `|d clearance / dt|` (`descriptor.clearance_rate_rows`).  Segments that fail the test are halved -- one more launch
        over all midpoints of all edges per level -- until they pass, a midpoint with clearance <= 0 is found (the edge is
        invalid), or `max_depth` halvings are reached (clearance below ~`L * width / 2` along the whole piece: inside the
        contact band; reported in the detail and counted as free, which is what a failed SLSQP solve amounts to).

        `return_detail` adds a dict: distance evaluations issued, levels used, unresolved segments, the smallest clearance
        seen per edge."""
        self._check_q(starts, "start")
        self._check_q(goals, "goal")
        if tuple(starts.shape) != tuple(goals.shape):
            raise ValueError("starts and goals must have the same shape")
        if resolution <= 0:
            raise ValueError("Resolution must be positive")
        in_torch = _is_torch(starts)
        shape = tuple(starts.shape[:-1])
        E = int(np.prod(shape)) if len(shape) else 1
        scene = self._scene(threshold, exact_distance=True)
        arrays = self._scenes["dist"]["arrays"]
        if arrays["self_pairs"].shape[0] + arrays["n_env_pairs_before_pruning"] == 0:
            raise ValueError("min() arg is an empty sequence")
        M = descriptor.clearance_rate_rows(self._arrays, arrays)
        if not np.all(np.isfinite(M)):
            raise NotImplementedError("continuous edge validation needs bounded lever arms: a prismatic joint below a "
                                      "revolute one makes the rate of change of the clearance unbounded")
        dev_index = self.device_index
        base_keep, base_ptr = _cabi.base_pose_arg(self.base_pose)
        lib = _cabi.load()
        stream = _cuda_stream_ptr(dev_index)
        hptr = self._handle().ptr
        sd = self._to_device(starts, torch.float32).reshape(E, self.dof).double()
        gd = self._to_device(goals, torch.float32).reshape(E, self.dof).double()
        dev = sd.device

        def clearance(edge, t):
            """signed minimum distance - threshold at the points t of the edges `edge` (one launch)"""
            q = (sd[edge] + t[:, None] * (gd[edge] - sd[edge])).float().contiguous()
            n = q.shape[0]
            out = torch.empty((n,), dtype=torch.float32, device=dev)
            if n:
                vd = torch.empty((n,), dtype=torch.uint8, device=dev)
                _cabi.check(lib.nbx_in_collision_f32(hptr, scene.ptr, q.data_ptr(), n, base_ptr, float(threshold),
                                                     vd.data_ptr(), out.data_ptr(), stream))
            return out.double() - float(threshold)

        from ..planning.sampling_based.continuous import decide_edges
        with torch.cuda.device(dev_index):
            clear, det = decide_edges(sd, gd, resolution, max_distance, M, clearance, max_depth)
        del base_keep
        conv = (lambda x: x.to(starts.device).reshape(shape)) if in_torch else (lambda x: x.cpu().numpy().reshape(shape))
        if not return_detail:
            return conv(clear)
        det["min_clearance"], det["rate_bound"] = conv(det["min_clearance"]), conv(det["rate_bound"])
        return conv(clear), det
