"""ncu target: the fused distance query (Arm.min_distance) on 2^20 PRM configurations, two calls (profile the second)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from numbotics_b200 import workloads
arm, cubes = workloads.prm_scene(device=0)
q = workloads.random_configs_cuda(arm, 1 << 20, seed=0)
for _ in range(2):
    d = arm.min_distance(q)
torch.cuda.synchronize()
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
a.record(); d = arm.min_distance(q); b.record(); torch.cuda.synchronize()
print("ok", float(d.min()), float(d.mean()), "ms", a.elapsed_time(b))
