import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from numbotics_b200 import workloads
def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts=[]
    for i in range(n):
        a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
tag = os.environ.get("TAG","")
arm30 = workloads.serial_arm(30, device=0)
B30 = 1 << 22
q30 = workloads.random_configs_cuda(arm30, B30, seed=1)
b = timeit(lambda: arm30.fk_jacobian(q30, "tool"))
print(f"{tag} 30dof fk+jac: {b:.3f} ms -> {B30/b*1e3:.3e} cfg/s, {904*B30/b/1e6:.0f} GB/s")
b = timeit(lambda: arm30.forward_kinematics(q30, "tool"))
print(f"{tag} 30dof fk only: {b:.3f} ms")
arm, cubes = workloads.prm_scene(device=0)
q = workloads.random_configs_cuda(arm, 1 << 22, seed=0)
b = timeit(lambda: arm.fk_jacobian(q, "forearm_link"))
print(f"{tag} kinova forearm_link (sparse, generic kernel) fk+jac 4M: {b:.3f} ms")
