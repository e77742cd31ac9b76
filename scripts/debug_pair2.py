import sys, os, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from test_gjk_gpu import flyer, SHAPES, make_obstacle
from oracle import collision
link_kind, obs_kind, mm = "box", "cyl", "sharp"
arm, w = flyer(SHAPES[link_kind](), margin_model=mm)
obs = make_obstacle(obs_kind, w)
rng = np.random.default_rng(zlib.crc32(f"{link_kind}-{obs_kind}".encode()))
n = 12000
q = np.zeros((n, 6))
direction = rng.normal(size=(n, 3)); direction /= np.linalg.norm(direction, axis=1, keepdims=True)
q[:, :3] = direction * rng.uniform(0.0, 0.9, (n, 1)); q[:, 3:] = rng.uniform(-3.0, 3.0, (n, 3))
q = q.astype(np.float32)
so = collision.SceneOracle(arm, mm)
d = so.min_distance(q.astype(np.float64))
hit = arm.in_collision(q, 0.0)
dist = arm.min_distance(q)
bad = np.where((np.abs(d) > 1e-4) & (hit != (d < 0)))[0]
print("bad", bad, d[bad], hit[bad], dist[bad])
for i in bad:
    qi = q[i].astype(np.float64)
    # shrink test: how deep is the overlap? move the flyer along +-each axis by small steps and watch the oracle
    for delta in (1e-5, 1e-4, 3e-4, 1e-3, 3e-3):
        best = -1
        for ax in range(3):
            for sgn in (-1, 1):
                qq = qi.copy(); qq[ax] += sgn * delta
                dd = so.min_distance(qq[None])[0]
                best = max(best, dd)
        print("  delta", delta, "max oracle distance over +-xyz shifts", best)
