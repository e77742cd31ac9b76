import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from numbotics_b200 import workloads
from numbotics_b200.planning.sampling_based.roadmap import knn_earlier
arm, cubes = workloads.prm_scene(device=0)
x = workloads.random_configs_cuda(arm, 100000, seed=0)
for _ in range(2):
    idx, ok = knn_earlier(x, 10)
torch.cuda.synchronize()
print("ok")
