"""ncu target: the FK + Jacobian kernels at their bench sizes, two launches each (profile the second):
7-DoF float32 2^24 (headline), 7-DoF float64 2^24, 30-DoF float32 2^22."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from numbotics_b200 import workloads
arm, cubes = workloads.prm_scene(device=0)
q = workloads.random_configs_cuda(arm, 1 << 24, seed=0)
for _ in range(2):
    T, J = arm.fk_jacobian(q, "tool_frame")
torch.cuda.synchronize()
del T, J
qd = q.double()
for _ in range(2):
    T, J = arm.fk_jacobian(qd, "tool_frame")
torch.cuda.synchronize()
del T, J, qd
arm30 = workloads.serial_arm(30, device=0)
q30 = workloads.random_configs_cuda(arm30, 1 << 22, seed=1)
for _ in range(2):
    T30, J30 = arm30.fk_jacobian(q30, "tool")
torch.cuda.synchronize()
print("ok")
