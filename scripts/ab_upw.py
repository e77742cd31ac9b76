"""A/B of the collision kernels' range length (NBX_COLLIDE_UPW = tiles of 32 units per warp; 0 = library default)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from numbotics_b200 import workloads

arm, cubes = workloads.prm_scene(device=0)
q = workloads.random_configs_cuda(arm, 1 << 22, seed=0)


def best(fn, n=12, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[0], ts[len(ts) // 2]


b, m = best(lambda: arm.in_collision(q))
print(f"NBX_COLLIDE_UPW={os.environ.get('NBX_COLLIDE_UPW', 'default')}: in_collision 2^22 best {b:.4f} ms median {m:.4f} ms", flush=True)
