import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from numbotics_b200 import workloads
arm, cubes = workloads.prm_scene(device=0)
q = workloads.random_configs_cuda(arm, 100000, seed=0)
s, g = workloads.knn_edges_cuda(q, 10)
for _ in range(2):
    v = arm.edges_valid(s, g, 0.1)
torch.cuda.synchronize()
print("ok", float(v.float().mean()))
