"""Round-2 probes on one GPU (not the contract bench): fp32 FK/J error against chain depth, bisection timing,
concurrent host callers.  `python scripts/r02_probe.py [depth|bisect|threads|all]`."""
import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from numbotics_b200 import workloads, _cabi

what = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = torch.device("cuda", 0)


def best_ms(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)


if what in ("depth", "all"):
    # fp32 kernels against the float64 kernels (themselves at 1e-11 of the reference's numba output) on 2^20 configurations
    print("depth  max_rel_err_T  max_rel_err_J  p99.9_T  p99.9_J   (rel = max|f32 - f64| / max(max|f64|, 1) per matrix)")
    for n in (7, 8, 12, 16, 20, 24, 28, 30, 32, 48):
        arm = workloads.serial_arm(n, device=0) if n != 7 else workloads.kinova_arm(device=0)
        frame = "tool" if n != 7 else "tool_frame"
        B = 1 << 20
        q = workloads.random_configs_cuda(arm, B, seed=5)
        T32, J32 = arm.fk_jacobian(q, frame)
        T64, J64 = arm.fk_jacobian(q.double(), frame)
        def rel(a, b):
            e = (a.double() - b).abs().reshape(B, -1).amax(1) / b.abs().reshape(B, -1).amax(1).clamp(min=1.0)
            return float(e.max()), float(torch.quantile(e[: 1 << 18], 0.999))
        (eT, pT), (eJ, pJ) = rel(T32, T64), rel(J32, J64)
        print(f"{n:5d}  {eT:.3e}  {eJ:.3e}  {pT:.3e}  {pJ:.3e}")
        del T32, J32, T64, J64, q

if what in ("bisect", "all"):
    arm, cubes = workloads.prm_scene(device=0)
    q = workloads.random_configs_cuda(arm, 1 << 22, seed=0)
    hit = arm.in_collision(q, 1e-6)
    S = q[hit][: 120000].contiguous()
    c = torch.zeros(7, device=dev)
    ms = best_ms(lambda: arm.bisect_to_collision(c, S, 15, 1e-6), n=5, warm=2)
    print(f"bisection: {S.shape[0]} samples x 15 rounds {ms:.3f} ms -> {S.shape[0]*15/ms*1e3:.3e} midpoint checks/s")
    t0 = time.perf_counter()
    for _ in range(20):
        arm.bisect_to_collision(c, S, 15, 1e-6)
    torch.cuda.synchronize()
    print(f"bisection wall per call (20 back to back): {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms")

if what in ("threads", "all"):
    arm, cubes = workloads.prm_scene(device=0)
    lib = _cabi.load()
    keep, base = _cabi.base_pose_arg(arm.base_pose)
    fid = arm._arrays["frame_id"]["tool_frame"]
    h = arm._handle()
    n_thr, B = 8, 1 << 18
    bufs = []
    for i in range(n_thr):
        qh = torch.empty((B, 7), dtype=torch.float32).pin_memory()
        qh.copy_(workloads.random_configs_cuda(arm, B, seed=i))
        bufs.append((qh, torch.empty((B, 4, 4), dtype=torch.float32).pin_memory(), torch.empty((B, 6, 7), dtype=torch.float32).pin_memory()))

    def call(i):
        qh, Th, Jh = bufs[i]
        _cabi.check(lib.nbx_fk_jac_host_f32(h.ptr, fid, qh.data_ptr(), B, base, 0, Th.data_ptr(), Jh.data_ptr()))

    for i in range(n_thr): call(i)
    reps = 6
    t0 = time.perf_counter()
    for _ in range(reps):
        for i in range(n_thr): call(i)
    serial = time.perf_counter() - t0

    def worker(i):
        for _ in range(reps): call(i)
    th = [threading.Thread(target=worker, args=(i,)) for i in range(n_thr)]
    t0 = time.perf_counter()
    [t.start() for t in th]; [t.join() for t in th]
    conc = time.perf_counter() - t0
    tot = n_thr * reps * B
    print(f"nbx_fk_jac_host_f32, {n_thr} callers x {B} cfgs: one after the other {tot/serial:.3e} cfg/s, {n_thr} host threads {tot/conc:.3e} cfg/s "
          f"(x{serial/conc:.2f})")
    # small calls: where the per-call latency, not PCIe, is the cost
    Bs = 4096
    def call_small(i):
        qh, Th, Jh = bufs[i]
        _cabi.check(lib.nbx_fk_jac_host_f32(h.ptr, fid, qh.data_ptr(), Bs, base, 0, Th.data_ptr(), Jh.data_ptr()))
    reps = 200
    t0 = time.perf_counter()
    for _ in range(reps):
        for i in range(n_thr): call_small(i)
    serial = time.perf_counter() - t0
    def worker_s(i):
        for _ in range(reps): call_small(i)
    th = [threading.Thread(target=worker_s, args=(i,)) for i in range(n_thr)]
    t0 = time.perf_counter()
    [t.start() for t in th]; [t.join() for t in th]
    conc = time.perf_counter() - t0
    print(f"nbx_fk_jac_host_f32, {n_thr} callers x {Bs} cfgs: one after the other {serial/(reps*n_thr)*1e6:.1f} us/call, {n_thr} host threads "
          f"{conc/(reps*n_thr)*1e6:.1f} us/call aggregate (x{serial/conc:.2f})")
