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
import numpy as np
q64 = q30[: 1 << 20].double()
b = timeit(lambda: arm30.fk_jacobian(q64, "tool"), n=5)
print(f"{tag} 30dof f64 fk+jac 1M: {b:.3f} ms -> {q64.shape[0]/b*1e3:.3e} cfg/s")
arm48 = workloads.serial_arm(48, device=0)
q48 = workloads.random_configs_cuda(arm48, 1 << 21, seed=1)
b = timeit(lambda: arm48.fk_jacobian(q48, "tool"), n=5)
print(f"{tag} 48dof f32 fk+jac 2M: {b:.3f} ms -> {q48.shape[0]/b*1e3:.3e} cfg/s, {(48*28+64)*q48.shape[0]/b/1e6:.0f} GB/s")
b = timeit(lambda: arm48.fk_jacobian(q48, "link_24"), n=5)
print(f"{tag} 48dof f32 link_24 (sparse) fk+jac 2M: {b:.3f} ms")
