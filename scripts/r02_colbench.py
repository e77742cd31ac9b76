"""A/B timing of the collision queries on resident buffers (scratch; the contract bench is bench.py):
in_collision 2^22, edges 10^6 and the 125 k shard of an 8-GPU run, IRIS bisection 1.2e5 x 15 rounds."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from numbotics_b200 import workloads


def timeit(fn, n=8, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        fn(); ev[i + 1].record()
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(n)]
    return min(ts), sorted(ts)[len(ts) // 2]


arm, cubes = workloads.prm_scene(device=0)
q = workloads.random_configs_cuda(arm, 1 << 22, seed=0)
best, med = timeit(lambda: arm.in_collision(q))
hit = arm.in_collision(q)
print(f"in_collision 2^22: best {best:.4f} ms med {med:.4f} -> {(1 << 22) / best * 1e3:.3e} cfg/s hits {int(hit.sum())}")
for nb in (1 << 20, 1 << 19):
    best, med = timeit(lambda: arm.in_collision(q[:nb]))
    print(f"in_collision {nb}: best {best:.4f} ms med {med:.4f}")
s, g = workloads.knn_edges_cuda(q[:100000], 10)
for E in (1000000, 500000, 250000, 125000):
    ss, gg = s[:E].contiguous(), g[:E].contiguous()
    best, med = timeit(lambda: arm.edges_valid(ss, gg, 0.1))
    v = arm.edges_valid(ss, gg, 0.1)
    print(f"edges {E}: best {best:.4f} ms med {med:.4f} -> {E / best * 1e3:.3e} edges/s valid {int(v.sum())}")
if hasattr(arm, "bisect_to_collision"):
    hits = q[hit][:120000]
    free = q[~hit][:hits.shape[0]]
    n = min(hits.shape[0], free.shape[0])
    try:
        best, med = timeit(lambda: arm.bisect_to_collision(free[:n], hits[:n], 15))
        print(f"bisect {n} x 15: best {best:.4f} ms med {med:.4f}")
    except Exception as ex:  # signature drift: this is a scratch script
        print("bisect skipped:", ex)
