"""Timing of the SURVEY 8(f) "next" rows on one GPU (not part of the contract bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from numbotics_b200 import workloads

def best_ms(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)

arm, cubes = workloads.prm_scene(device=0)
dev = torch.device("cuda", 0)
# f.1 IK: targets = FK of random configurations, start = target configuration + noise
for dt in (torch.float32, torch.float64):
    B = 1 << 18
    qt = workloads.random_configs_cuda(arm, B, seed=3, dtype=dt)
    pose = arm.forward_kinematics(qt, "tool_frame")
    q0 = qt + 0.3 * torch.randn_like(qt)
    ok, q, it = arm.inverse_kinematics(pose, q0, "tool_frame", return_iters=True)
    ms = best_ms(lambda: arm.inverse_kinematics(pose, q0, "tool_frame"), n=3, warm=1)
    print(f"IK {dt}: {B} problems {ms:.2f} ms -> {B/ms*1e3:.3e} problems/s, converged {float(ok.float().mean()):.3f}, mean iters {float(it.float().mean()):.1f}")
# f.3 bisection: 2^17 colliding samples, 15 rounds
q = workloads.random_configs_cuda(arm, 1 << 22, seed=0)
hit = arm.in_collision(q, 1e-6)
S = q[hit][: 1 << 17].contiguous()
c = torch.zeros(7, device=dev)
ms = best_ms(lambda: arm.bisect_to_collision(c, S, 15, 1e-6), n=3, warm=1)
print(f"bisection: {S.shape[0]} samples x 15 rounds {ms:.2f} ms -> {S.shape[0]/ms*1e3:.3e} samples/s ({S.shape[0]*15/ms*1e3:.3e} midpoint checks/s)")
# f.4 closest pair + gradient
Bc = 1 << 20
ms0 = best_ms(lambda: arm.min_distance(q[:Bc]), n=3, warm=1)
ms1 = best_ms(lambda: arm.closest_to(q[:Bc]), n=3, warm=1)
ms2 = best_ms(lambda: arm.closest_to(q[:Bc], gradient=True), n=3, warm=1)
print(f"min_distance: {Bc} cfgs {ms0:.2f} ms -> {Bc/ms0*1e3:.3e} cfg/s; closest_to {ms1:.2f} ms -> {Bc/ms1*1e3:.3e} cfg/s; with gradient {ms2:.2f} ms -> {Bc/ms2*1e3:.3e} cfg/s  (host-side subject/target object mapping included)")
