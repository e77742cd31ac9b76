"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from numbotics_b200 import workloads
arm, cubes = workloads.prm_scene(device=0)
dev = torch.device("cuda", 0)
for B in (1, 37, 300):
    q = workloads.random_configs_cuda(arm, B, seed=B)
    arm.fk_jacobian(q, "tool_frame"); arm.fk_jacobian(q, "forearm_link"); arm.fk_jacobian(q.double(), "tool_frame")
    arm.in_collision(q); arm.min_distance(q); arm.closest_to(q, gradient=True)
    if B > 1:
        arm.edges_valid(q[: B // 2], q[B // 2: 2 * (B // 2)], 0.1)
q = workloads.random_configs_cuda(arm, 2000, seed=1)
hit = arm.in_collision(q, 1e-6)
if int(hit.sum()) > 0:
    arm.bisect_to_collision(torch.zeros(7, device=dev), q[hit], 5, 1e-6)
pose = arm.forward_kinematics(q[:100], "tool_frame")
arm.inverse_kinematics(pose, q[:100] + 0.1, "tool_frame")
qn = q[:500].cpu().numpy()
arm.fk_jacobian(qn, "tool_frame"); arm.fk_jacobian(qn.astype(np.float64), "tool_frame")
for n in (12, 30, 40):
    a = workloads.serial_arm(n, device=0)
    qq = workloads.random_configs_cuda(a, 100, seed=n)
    a.fk_jacobian(qq, "tool"); a.fk_jacobian(qq.double(), "tool"); a.fk_jacobian(qq, f"link_{n // 2}")
torch.cuda.synchronize()
print("sanitize target ok")
