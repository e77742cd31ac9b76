import sys, os, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from test_gjk_gpu import flyer, SHAPES, make_obstacle
from oracle import collision
for link_kind, obs_kind, mm in [("box","cyl","sharp"),("sphere","box","sharp"),("cyl","cyl","bullet"),("disc","cyl","sharp")]:
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
    for thr in (0.0, 0.03):
        hit = arm.in_collision(q, thr)
        bad = np.where((np.abs(d - thr) > 1e-4) & (hit != (d < thr)))[0]
        print(link_kind, obs_kind, mm, "thr", thr, "bad verdicts", len(bad), [(float(d[i]), bool(hit[i])) for i in bad[:6]])
    dist = arm.min_distance(q)
    err = np.abs(dist - d); err[d <= 0] = 0
    print("   max dist err", err.max(), "at d", d[np.argmax(err)], "n>5e-5:", int((err > 5e-5).sum()), "nonfinite/huge:", int((~np.isfinite(dist) | (np.abs(dist) > 1e3)).sum()))
