"""Timing / ncu target: every allowed pair of every configuration (Arm.pair_distances = batched Arm.collisions), 2^18 PRM
configurations x 58 pairs, with and without the distance gradients."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from numbotics_b200 import workloads
arm, cubes = workloads.prm_scene(device=0)
q = workloads.random_configs_cuda(arm, 1 << 18, seed=0)
def best(fn, n=3):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
pairs, d = arm.pair_distances(q)
print("pairs", len(pairs), "distance only", best(lambda: arm.pair_distances(q)), "ms; with points", best(lambda: arm.pair_distances(q, points=True)),
      "ms; with gradient", best(lambda: arm.pair_distances(q, gradient=True)), "ms;", q.shape[0], "configurations; min", float(d.min()))
