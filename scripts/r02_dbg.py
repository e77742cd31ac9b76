import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from numbotics_b200 import workloads, _cabi
arm, cubes = workloads.prm_scene(device=0)
q = workloads.random_configs_np(arm, 2000, seed=6).astype(np.float32)
def tryit(name, fn):
    try:
        r = fn(); torch.cuda.synchronize(); print(name, "ok", getattr(r, "shape", None), flush=True)
    except Exception as e:
        print(name, "FAILED", str(e)[:200], flush=True)
tryit("thr 0", lambda: arm.in_collision(q))
tryit("thr 0.02", lambda: arm.in_collision(q, 0.02))
tryit("thr -0.003", lambda: arm.in_collision(q, -0.003))
tryit("thr 0 again", lambda: arm.in_collision(q))
tryit("thr -0.02", lambda: arm.in_collision(q, -0.02))
tryit("thr 0 again", lambda: arm.in_collision(q))
tryit("min_distance", lambda: arm.min_distance(q))
tryit("thr 0 again", lambda: arm.in_collision(q))
tryit("pair_distances", lambda: arm.pair_distances(q)[1])
tryit("thr 0 again", lambda: arm.in_collision(q))
arm2, cubes2 = workloads.prm_scene(device=0)
tryit("new scene thr 0", lambda: arm2.in_collision(q))
tryit("edges -0.003", lambda: arm.edges_valid(q[:1000], q[1000:], 0.1, threshold=-0.003))
tryit("edges 0", lambda: arm.edges_valid(q[:1000], q[1000:], 0.1))
tryit("closest", lambda: arm.closest_to(q, gradient=True).distance)
tryit("bisect", lambda: arm.bisect_to_collision(np.zeros(7, np.float32), q[:500], 6, 1e-6))
tryit("thr 0 again", lambda: arm.in_collision(q))
