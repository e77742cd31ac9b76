"""BASELINE-size runs checked through size-independent properties (the oracle finishes only small samples in seconds):
independence from the batch layout (a batch gives the bits of its contiguous shards, which is what the multi-GPU
sharding relies on), rigid-body invariants, and oracle spot checks on random rows."""
import numpy as np
import pytest

from numbotics_b200 import workloads
from numbotics_b200.dist import shard_bounds
from oracle import collision, fkjac

pytestmark = pytest.mark.gpu

EPS = 1e-4


@pytest.fixture(scope="module")
def scene(cuda):
    arm, cubes = workloads.prm_scene(device=0)
    arm._test_keepalive = cubes
    return arm, collision.SceneOracle(arm)


def test_in_collision_4M_configs(scene, cuda):
    """BASELINE configs[2]: 2^22 configurations."""
    import torch
    arm, so = scene
    B = 1 << 22
    q = workloads.random_configs_cuda(arm, B, seed=0)
    hit = arm.in_collision(q)
    assert hit.shape == (B,) and hit.dtype == torch.bool
    assert 0.02 < float(hit.float().mean()) < 0.04
    # shards of uneven, non-tile-aligned sizes reproduce the batch bit for bit (3 and 7 "GPUs")
    for world in (3, 7):
        parts = [arm.in_collision(q[lo:hi]) for lo, hi in (shard_bounds(B, r, world) for r in range(world))]
        assert torch.equal(torch.cat(parts), hit)
    assert torch.equal(arm.in_collision(q), hit)                  # and the call is deterministic
    # distance query agrees with the verdict query away from the threshold
    d = arm.min_distance(q[: 1 << 20])
    clear = (d.abs() > EPS)
    assert torch.equal((d < 0)[clear], hit[: 1 << 20][clear])
    # oracle spot check
    idx = torch.randint(0, B, (4096,), device=cuda)
    qs = q[idx].cpu().numpy()
    do = so.min_distance(qs.astype(np.float64))
    ok = np.abs(do) > EPS
    assert np.array_equal(hit[idx].cpu().numpy()[ok], (do < 0)[ok])


def test_edges_1M(scene, cuda):
    """BASELINE configs[3]: 100k nodes x 10 nearest neighbours, resolution 0.1."""
    import torch
    arm, so = scene
    nodes = workloads.random_configs_cuda(arm, 100000, seed=0)
    s, g = workloads.knn_edges_cuda(nodes, 10)
    E = s.shape[0]
    assert E == 1000000
    valid, ns = arm.edges_valid(s, g, 0.1, return_samples=True)
    assert 0.9 < float(valid.float().mean()) < 0.99
    # sample counts: NumPy's arange length rule + the end point, in float64 (connectors.py:62-69)
    dist = (g.double() - s.double()).norm(dim=1)
    expect = torch.ceil(1.0 / (0.1 / dist)).to(torch.int32) + 1
    assert torch.equal(ns, expect)
    # contiguous E/G slices reproduce the batch bit for bit (the multi-GPU partition of the bench)
    for world in (2, 8):
        parts = [arm.edges_valid(s[lo:hi], g[lo:hi], 0.1) for lo, hi in (shard_bounds(E, r, world) for r in range(world))]
        assert torch.equal(torch.cat(parts), valid)
    # an edge whose two end points collide can not be valid; a valid edge has collision-free end points
    hs, hg = arm.in_collision(s), arm.in_collision(g)
    assert not bool((valid & (hs | hg)).any())
    # oracle on a random subset (sequential reference semantics), edges with an in-band sample excluded
    idx = torch.randperm(E, device=cuda)[:1500]
    vo, ons, md = so.edges_valid(s[idx].cpu().numpy().astype(np.float64), g[idx].cpu().numpy().astype(np.float64), 0.1, return_detail=True)
    assert np.array_equal(ns[idx].cpu().numpy(), ons)
    # an invalid edge's minimum is only known up to its first invalid sample on the device side; the oracle's edge minimum
    # decides which edges are inside the contact band
    clear = np.abs(md) > EPS
    assert 1.0 - clear.mean() <= 0.005, f"the contact band excludes {1.0 - clear.mean():.2e} of the samples"
    assert np.array_equal(valid[idx].cpu().numpy()[clear], np.asarray(vo, bool)[clear])


def test_serial30_4M_configs(cuda):
    """BASELINE configs[4] slice: 2^22 configurations of the 30-DoF chain."""
    import torch
    arm = workloads.serial_arm(30, device=0)
    B = 1 << 22
    q = workloads.random_configs_cuda(arm, B, seed=1)
    T, J = arm.fk_jacobian(q, "tool")
    R = T[:, :3, :3]
    assert float((R @ R.transpose(1, 2) - torch.eye(3, device=cuda)).abs().max()) < 5e-5
    assert float((T[:, 3] - torch.tensor([0.0, 0.0, 0.0, 1.0], device=cuda)).abs().max()) == 0.0
    assert float((J[:, 3:, :].norm(dim=1) - 1).abs().max()) < 5e-5
    # the last joint's column: omega x (p_e - p_30) with |p_e - p_30| = 0.05 (the fixed tool offset)
    assert float((J[:, :3, 29].norm(dim=1) - 0.05).abs().max()) < 1e-5
    for world in (3,):
        parts = [arm.fk_jacobian(q[lo:hi], "tool") for lo, hi in (shard_bounds(B, r, world) for r in range(world))]
        assert torch.equal(torch.cat([p[0] for p in parts]), T) and torch.equal(torch.cat([p[1] for p in parts]), J)
    idx = torch.randint(0, B, (2048,), device=cuda)
    ora = fkjac.ArmOracle(arm._link_joint_sequence, 30, arm.base_pose)
    qs = q[idx].double().cpu().numpy()
    rel = lambda a, b: float(np.abs(a - b).max())
    assert rel(T[idx].cpu().numpy(), ora.forward_kinematics(qs, "tool")) < 1e-5
    assert rel(J[idx].cpu().numpy(), ora.jacobian(qs, "tool")) < 1e-5
