"""Continuous edge validation (`Arm.edges_clear`, batched ContinuousConnector) next to the discrete query on the PRM
workload of BASELINE configs[3]: 10^5 of the 10^6 candidate edges, resolution 0.1.  CUDA events around the whole call
(a handful of launches of the fused distance query + torch index arithmetic between them), best of 3."""
import sys
import torch

sys.path.insert(0, ".")
from numbotics_b200 import workloads

arm, cubes = workloads.prm_scene(device=0)
nodes = workloads.random_configs_cuda(arm, 100000, seed=0)
s_all, g_all = workloads.knn_edges_cuda(nodes, 10)
E = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
s, g = s_all[:E].contiguous(), g_all[:E].contiguous()


def best_ms(fn, n=3):
    fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        out.append(a.elapsed_time(b))
    return min(out)


clear, det = arm.edges_clear(s, g, 0.1, return_detail=True)
valid = arm.edges_valid(s, g, 0.1)
t_c = best_ms(lambda: arm.edges_clear(s, g, 0.1))
t_d = best_ms(lambda: arm.edges_valid(s, g, 0.1))
print(f"edges {E}: continuous {t_c:.3f} ms = {E / t_c * 1e3:.3e} edges/s ({det['distance_evaluations']} distance evaluations for "
      f"{det['grid_samples']} grid samples, {det['levels']} levels, {det['unresolved_segments']} unresolved); "
      f"discrete {t_d:.3f} ms; clear {float(clear.float().mean()):.4f} valid {float(valid.float().mean()):.4f} "
      f"clear-but-not-valid {int((clear & ~valid).sum())} valid-but-not-clear {int((valid & ~clear).sum())}")
