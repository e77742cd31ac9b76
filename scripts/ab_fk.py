import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from numbotics_b200 import workloads
def timeit(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts=[]
    for i in range(n):
        a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts), sorted(ts)[len(ts)//2]
arm, cubes = workloads.prm_scene(device=0)
B = 1 << 24
q = workloads.random_configs_cuda(arm, B, seed=0)
tag = os.environ.get("TAG", "")
b, m = timeit(lambda: arm.fk_jacobian(q, "tool_frame"))
print(f"{tag} fk+jac: best {b:.3f} med {m:.3f} ms  {260*B/b/1e6:.0f} GB/s")
b, m = timeit(lambda: arm.forward_kinematics(q, "tool_frame"))
print(f"{tag} fk only: best {b:.3f} ms")
b, m = timeit(lambda: arm.jacobian(q, "tool_frame"))
print(f"{tag} jac only: best {b:.3f} ms")
# practical HBM ceilings on this GPU for context
x = torch.empty(1 << 30, dtype=torch.float32, device="cuda")
y = torch.empty(1 << 30, dtype=torch.float32, device="cuda")
b, m = timeit(lambda: x.fill_(1.0), n=10)
print(f"{tag} write-only 4 GiB fill: {4.295/b*1e3:.0f} GB/s")
b, m = timeit(lambda: y.copy_(x), n=10)
print(f"{tag} copy 4+4 GiB: {8.59/b*1e3:.0f} GB/s")
