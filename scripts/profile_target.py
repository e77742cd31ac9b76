"""Short program for ncu: two launches of each hot kernel at its BASELINE size (one warm, one measured)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from numbotics_b200 import workloads

arm, cubes = workloads.prm_scene(device=0)
q = workloads.random_configs_cuda(arm, 1 << 24, seed=0)
for _ in range(2):
    T, J = arm.fk_jacobian(q, "tool_frame")
torch.cuda.synchronize()
for _ in range(2):
    hit = arm.in_collision(q[: 1 << 22])
torch.cuda.synchronize()
s, g = workloads.knn_edges_cuda(q[:100000], 10)
for _ in range(2):
    v = arm.edges_valid(s, g, 0.1)
torch.cuda.synchronize()
del T, J
arm30 = workloads.serial_arm(30, device=0)
q30 = workloads.random_configs_cuda(arm30, 1 << 22, seed=1)
for _ in range(2):
    T30, J30 = arm30.fk_jacobian(q30, "tool")
torch.cuda.synchronize()
print("ok", float(hit.float().mean()), float(v.float().mean()))
