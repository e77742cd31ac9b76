"""Scratch timing of the resident-in-HBM kernels (not the contract bench; see bench.py)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from numbotics_b200 import workloads, _cabi

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n+1)]
    ev[0].record()
    for i in range(n):
        fn(); ev[i+1].record()
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i+1]) for i in range(n)]
    return min(ts), sorted(ts)[len(ts)//2]

arm, cubes = workloads.prm_scene(device=0)
B = 1 << 24
q = workloads.random_configs_cuda(arm, B, seed=0)
best, med = timeit(lambda: arm.fk_jacobian(q, "tool_frame"))
print(f"fk+jac f32 7dof B={B}: best {best:.3f} ms med {med:.3f} ms -> {B/best*1e3:.3e} cfg/s, {260*B/best/1e6:.1f} GB/s")
best, med = timeit(lambda: arm.forward_kinematics(q, "tool_frame"))
print(f"fk only: best {best:.3f} ms -> {B/best*1e3:.3e} cfg/s")
Bc = 1 << 22
best, med = timeit(lambda: arm.in_collision(q[:Bc]), n=5)
print(f"in_collision B={Bc}: best {best:.3f} ms -> {Bc/best*1e3:.3e} cfg/s; hit frac {float(arm.in_collision(q[:Bc]).float().mean()):.4f}")
nodes = q[:100000]
s, g = workloads.knn_edges_cuda(nodes, 10)
best, med = timeit(lambda: arm.edges_valid(s, g, 0.1), n=5)
v, ns = arm.edges_valid(s, g, 0.1, return_samples=True)
print(f"edges E={s.shape[0]}: best {best:.3f} ms -> {s.shape[0]/best*1e3:.3e} edges/s; valid frac {float(v.float().mean()):.4f}; mean samples {float(ns.float().mean()):.2f}")
arm30 = workloads.serial_arm(30, device=0)
B30 = 1 << 22
q30 = workloads.random_configs_cuda(arm30, B30, seed=1)
best, med = timeit(lambda: arm30.fk_jacobian(q30, "tool"), n=5)
print(f"fk+jac f32 30dof B={B30}: best {best:.3f} ms -> {B30/best*1e3:.3e} cfg/s, {904*B30/best/1e6:.1f} GB/s")
