import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from numbotics_b200 import workloads
arm, cubes = workloads.prm_scene(device=0)
for dt in (np.float64, np.float32):
    for B in (4096, 1 << 20, 1 << 22):
        q = workloads.random_configs_np(arm, B, seed=0, dtype=dt)
        arm.fk_jacobian(q[:1024], "tool_frame")
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter(); T, J = arm.fk_jacobian(q, "tool_frame"); best = min(best, time.perf_counter() - t0)
        bytes_ = q.nbytes + T.nbytes + J.nbytes
        print(f"numpy {np.dtype(dt).name} B={B}: {best*1e3:.2f} ms -> {B/best:.3e} cfg/s, {bytes_/best/1e9:.1f} GB/s host<->device")
B = 1 << 20
q = workloads.random_configs_np(arm, B, seed=0, dtype=np.float32)
t0 = time.perf_counter(); h = arm.in_collision(q); print(f"numpy in_collision B={B}: {(time.perf_counter()-t0)*1e3:.2f} ms")
t0 = time.perf_counter(); h = arm.in_collision(q); print(f"numpy in_collision B={B}: {(time.perf_counter()-t0)*1e3:.2f} ms")
