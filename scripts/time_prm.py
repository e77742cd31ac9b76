import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from numbotics_b200 import workloads
from numbotics_b200.planning.sampling_based import ConnectorParams, DiscreteConnector, ArmValidityChecker
from numbotics_b200.planning.sampling_based.roadmap import BatchedPRM, knn_earlier
arm, cubes = workloads.prm_scene(device=0)
conn = DiscreteConnector(ConnectorParams(resolution=0.1, max_distance=np.pi, validity_checker=ArmValidityChecker(arm)))
lo, hi = workloads.sample_limits(arm.joint_limits)
prm = BatchedPRM(conn, lo, hi, k_nearest=10)
for N in (100000,):
    prm.build(N, seed=0)
    torch.cuda.synchronize(); t0 = time.perf_counter(); rm = prm.build(N, seed=1); torch.cuda.synchronize(); t1 = time.perf_counter()
    nodes = rm.nodes
    torch.cuda.synchronize(); t2 = time.perf_counter(); idx, ok = knn_earlier(nodes, 10); torch.cuda.synchronize(); t3 = time.perf_counter()
    print(f"N={N}: build {1e3*(t1-t0):.1f} ms (candidates {rm.candidates}, edges {rm.edges.shape[0]}), of which knn_earlier {1e3*(t3-t2):.1f} ms")
