import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from numbotics_b200 import workloads
arm, cubes = workloads.prm_scene(device=0)
q = workloads.random_configs_cuda(arm, 1 << 22, seed=0)
for _ in range(2):
    hit = arm.in_collision(q)
torch.cuda.synchronize()
print("ok", float(hit.float().mean()))
