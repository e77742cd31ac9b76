import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from numbotics_b200 import workloads
arm30 = workloads.serial_arm(30, device=0)
q30 = workloads.random_configs_cuda(arm30, 1 << 20, seed=1)
for _ in range(2):
    T30, J30 = arm30.fk_jacobian(q30, "tool")
torch.cuda.synchronize()
print("ok")
