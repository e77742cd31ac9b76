"""ncu target: two calls of the IRIS bisection (1.2e5 colliding samples x 15 rounds); profile the second one
(`--launch-skip 48 -c 46`: 2 launches of the in_collision call that picks the samples + 46 of the warm-up call)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from numbotics_b200 import workloads
arm, cubes = workloads.prm_scene(device=0)
q = workloads.random_configs_cuda(arm, 1 << 22, seed=0)
hit = arm.in_collision(q, 1e-6)
S = q[hit][:120000].contiguous()
c = torch.zeros(7, device=S.device)
for _ in range(2):
    out = arm.bisect_to_collision(c, S, 15, 1e-6)
torch.cuda.synchronize()
print("ok", tuple(out.shape))
