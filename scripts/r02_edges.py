"""Edge-validation timing at the per-GPU sizes of BASELINE configs[3] on 1 / 2 / 4 / 8 GPUs (one GPU, the shard sizes)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from numbotics_b200 import workloads
arm, cubes = workloads.prm_scene(device=0)
nodes = workloads.random_configs_cuda(arm, 100000, seed=0)
s, g = workloads.knn_edges_cuda(nodes, 10)
def best(fn, n=7, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
base = None
for G in (1, 2, 4, 8):
    E = s.shape[0] // G
    ss, gg = s[:E].contiguous(), g[:E].contiguous()
    ms = best(lambda: arm.edges_valid(ss, gg, 0.1))
    base = base or ms
    print(f"shard of {G}: {E} edges {ms:.3f} ms -> {E/ms*1e3:.3e} edges/s per GPU, x{G} = {G*E/ms*1e3:.3e}; strong-scaling efficiency {base/(G*ms):.3f}")
