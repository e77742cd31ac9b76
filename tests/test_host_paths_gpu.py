"""The *_host_* entry points of the C-ABI (HOST buffers in, HOST buffers out): chunked pipelines over pooled streams,
concurrent host threads (the IRIS thread-pool caller, reference numbotics/utils/threading.py:33-37), bit-equality with the
device-pointer entry points; the conservative verdict of a GJK that runs out of iterations; NCCL gather when the box has
two GPUs."""
import ctypes as C
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

from numbotics_b200 import _cabi, workloads
from oracle import collision

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def scene(cuda):
    arm, cubes = workloads.prm_scene(device=0)
    return arm, cubes


def _host_fk(arm, q, want=("T", "J")):
    lib = _cabi.load()
    keep, base = _cabi.base_pose_arg(arm.base_pose)
    B = q.shape[0]
    T = np.empty((B, 4, 4), q.dtype)
    J = np.empty((B, 6, arm.dof), q.dtype)
    fn = lib.nbx_fk_jac_host_f64 if q.dtype == np.float64 else lib.nbx_fk_jac_host_f32
    _cabi.check(fn(arm._handle().ptr, arm._arrays["frame_id"]["tool_frame"], q.ctypes.data, B, base, 0, T.ctypes.data, J.ctypes.data))
    return T, J


@pytest.mark.parametrize("B", [1, 4095, 4096 * 3 + 17, (1 << 19) * 3 + 5])
def test_fk_host_chunking_is_bitwise_the_device_path(scene, B):
    """Every chunking regime of the host pipeline (one small chunk, three short chunks, full 2^19 chunks + a ragged tail)."""
    import torch
    arm, _ = scene
    q = workloads.random_configs_np(arm, B, seed=3, dtype=np.float32)
    T, J = _host_fk(arm, q)
    Td, Jd = arm.fk_jacobian(torch.as_tensor(q, device="cuda:0"), "tool_frame")
    assert np.array_equal(T, Td.cpu().numpy()) and np.array_equal(J, Jd.cpu().numpy())


def test_fk_host_concurrent_threads(scene):
    """Eight host threads inside nbx_fk_jac_host_f32 at once: each borrows its own pipe, results are bit-equal to a serial call."""
    arm, _ = scene
    qs = [workloads.random_configs_np(arm, 50000 + 1000 * i, seed=10 + i, dtype=np.float32) for i in range(8)]
    want = [_host_fk(arm, q) for q in qs]
    got = [None] * 8
    errs = []

    def run(i):
        try:
            for _ in range(3):
                got[i] = _host_fk(arm, qs[i])
        except Exception as exc:   # pragma: no cover
            errs.append(exc)

    th = [threading.Thread(target=run, args=(i,)) for i in range(8)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs
    for (T, J), (Tw, Jw) in zip(got, want):
        assert np.array_equal(T, Tw) and np.array_equal(J, Jw)


@pytest.mark.parametrize("B", [1, 5000, (1 << 19) * 2 + 77])
def test_in_collision_host_equals_device(scene, B):
    import torch
    arm, _ = scene
    lib = _cabi.load()
    q = workloads.random_configs_np(arm, B, seed=4, dtype=np.float32)
    v = np.empty(B, np.uint8)
    keep, base = _cabi.base_pose_arg(arm.base_pose)
    sc = arm._scene(0.0)
    _cabi.check(lib.nbx_in_collision_host_f32(arm._handle().ptr, sc.ptr, q.ctypes.data, B, base, 0.0, v.ctypes.data))
    vd = arm.in_collision(torch.as_tensor(q, device="cuda:0")).cpu().numpy()
    assert np.array_equal(v.astype(bool), vd)


@pytest.mark.parametrize("E", [3, 40000, (1 << 17) * 2 + 9])
def test_edges_host_equals_device(scene, E):
    import torch
    arm, _ = scene
    lib = _cabi.load()
    s = workloads.random_configs_np(arm, E, seed=5, dtype=np.float32)
    g = (s + np.random.default_rng(6).normal(scale=0.3, size=s.shape)).astype(np.float32)
    v = np.empty(E, np.uint8)
    keep, base = _cabi.base_pose_arg(arm.base_pose)
    sc = arm._scene(0.0)
    _cabi.check(lib.nbx_edges_valid_host_f32(arm._handle().ptr, sc.ptr, s.ctypes.data, g.ctypes.data, E, base, 0.1, 0.0, 0.0, v.ctypes.data))
    vd = arm.edges_valid(torch.as_tensor(s, device="cuda:0"), torch.as_tensor(g, device="cuda:0"), 0.1).cpu().numpy()
    assert np.array_equal(v.astype(bool), vd)


_GJK_CAP_SCRIPT = r"""
import sys, numpy as np
sys.path.insert(0, %r)
from numbotics_b200 import workloads
from oracle import collision
arm, cubes = workloads.prm_scene(device=0)
so = collision.SceneOracle(arm)
rng = np.random.default_rng(7)
sgn = lambda n: rng.choice([-1.0, 1.0], n)
q = rng.normal(scale=0.5, size=(6000, 7))
q[:, 1] = rng.uniform(1.2, 2.41, 6000) * sgn(6000)
q[:, 3] = rng.uniform(2.2, 2.66, 6000) * sgn(6000)
q[:, 5] = rng.uniform(1.5, 2.23, 6000) * sgn(6000)
q = q.astype(np.float32)
hit = arm.in_collision(q)
d = so.min_distance(q.astype(np.float64))
colliding = d < -1e-4
free = d > 1e-4
print("RESULT", int(np.sum(colliding & ~hit)), int(np.sum(free & hit)), int(colliding.sum()), int(free.sum()))
"""


@pytest.mark.parametrize("iters", [2, 4])
def test_gjk_iteration_cap_never_yields_a_false_free(cuda, iters):
    """a verdict GJK that stops with its bracket still open must not call the pair free.  With the iteration cap
    forced down to 2 / 4 (NBX_GJK_ITERS, tests only) many grazing cylinder / box pairs stop undecided: every configuration
    the oracle calls colliding must still be reported as a hit (false hits are the allowed, conservative, direction)."""
    env = dict(os.environ, NBX_GJK_ITERS=str(iters))
    r = subprocess.run([sys.executable, "-c", _GJK_CAP_SCRIPT % ROOT], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")][0].split()
    false_free, false_hit, n_col, n_free = map(int, line[1:])
    assert n_col > 500 and n_free > 500
    assert false_free == 0
    assert false_hit > 0          # the cap really was in force (with 48 iterations there are none, see test_collision_gpu)


def test_nccl_gather_two_gpus(cuda):
    """dist.gather_shards over NCCL (north_star: NCCL only gathers verdicts / poses).  Needs two visible GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU visible: the NCCL gather is covered by bench.py --gpus N and the gloo tests")
    script = os.path.join(ROOT, "tests", "_nccl_gather_worker.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29561", script], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    assert "GATHER_OK" in r.stdout
