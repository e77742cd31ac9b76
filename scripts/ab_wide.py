"""30-DoF wide-kernel timing at two batch sizes (A/B runs: NBX_WIDE_KT is read once per process)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from numbotics_b200 import workloads


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[0], ts[len(ts) // 2]


tag = os.environ.get("TAG", "") + " kt=" + os.environ.get("NBX_WIDE_KT", "auto")
arm = workloads.serial_arm(30, device=0)
for lg in (16, 20, 22, 25):
    B = 1 << lg
    q = workloads.random_configs_cuda(arm, B, seed=1)
    T = torch.empty((B, 4, 4), device=q.device, dtype=q.dtype)
    best, med = timeit(lambda: arm.fk_jacobian(q, "tool"))
    print(f"{tag} 30dof 2^{lg} fk+jac: best {best:.4f} ms median {med:.4f} ms -> {B / best * 1e3:.3e} cfg/s, {904 * B / best / 1e6:.0f} GB/s")
    del q, T
    torch.cuda.empty_cache()
