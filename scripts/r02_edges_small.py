import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from numbotics_b200 import workloads
arm, cubes = workloads.prm_scene(device=0)
nodes = workloads.random_configs_cuda(arm, 100000, seed=0)
s, g = workloads.knn_edges_cuda(nodes, 10)
for E in (1000, 10000, 30000, 60000, 125000):
    ss, gg = s[:E].contiguous(), g[:E].contiguous()
    for _ in range(4):
        arm.edges_valid(ss, gg, 0.1)
    torch.cuda.synchronize()
# in_collision at small sizes for comparison
q = workloads.random_configs_cuda(arm, 1 << 22, seed=0)
for B in (4096, 32768, 131072, 1 << 20):
    for _ in range(4):
        arm.in_collision(q[:B])
    torch.cuda.synchronize()
