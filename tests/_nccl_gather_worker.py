"""Worker of tests/test_host_paths_gpu.py::test_nccl_gather_two_gpus (launched with torchrun, one rank per GPU): every rank
validates its contiguous shard of the edges, rank 0 gathers the verdict bytes over NCCL and checks them against one
single-GPU evaluation of all edges."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from numbotics_b200 import workloads                    # noqa: E402
from numbotics_b200.dist import init_from_env, shard_bounds, gather_shards   # noqa: E402


def main():
    rank, local_rank, world = init_from_env("nccl")
    torch.cuda.set_device(local_rank)
    arm, cubes = workloads.prm_scene(device=local_rank)
    nodes = workloads.random_configs_cuda(arm, 3000, seed=0)
    s_all, g_all = workloads.knn_edges_cuda(nodes, 5)
    E = s_all.shape[0] - 3                               # shards of unequal size
    s_all, g_all = s_all[:E], g_all[:E]
    lo, hi = shard_bounds(E, rank, world)
    v = arm.edges_valid(s_all[lo:hi].contiguous(), g_all[lo:hi].contiguous(), 0.1)
    full = gather_shards(v, E, dst=0)
    every = gather_shards(v, E)
    if rank == 0:
        want = arm.edges_valid(s_all, g_all, 0.1)
        assert full.shape[0] == E and torch.equal(full, want) and torch.equal(every, want)
        print("GATHER_OK", int(want.sum()), E, flush=True)
    else:
        assert full is None and every.shape[0] == E
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
