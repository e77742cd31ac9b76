"""The float32 device code of csrc/nbx_gjk.cuh (support functions, witness GJK, EPA penetration depth) compiled for the
HOST with g++ and compared with the float64 oracle, pair by pair: 10^4 random overlapping box / cylinder / sphere pairs and
the exactly symmetric poses that degenerate the GJK seed.  Runs without a GPU; the same header is what the kernels compile.
(This cross-check is also what caught the oracle's own cylinder-support bug on cap-plane directions, see nbx_oracle.c.)"""
import os
import re
import subprocess

import pytest

from oracle import build as oracle_build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H = os.path.join(ROOT, "tests", "host_harness")


def _build(src, out):
    lib = oracle_build.build()
    cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=fast", "-I", os.path.join(H, "fake"), "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(ROOT, "numbotics_b200", "csrc"), os.path.join(H, src), "-o", out, lib,
           "-Wl,-rpath," + os.path.dirname(lib), "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_f32_penetration_depth_matches_the_oracle_on_random_pairs(tmp_path, seed):
    exe = str(tmp_path / "epa_host")
    _build("epa_host.cpp", exe)
    r = subprocess.run([exe, str(seed), "60000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    m = re.search(r"tested (\d+) bad (\d+) worst ([0-9.e+-]+)", r.stdout)
    assert m, r.stdout[-2000:]
    tested, bad, worst = int(m.group(1)), int(m.group(2)), float(m.group(3))
    # penetrating pairs from 2e-5 m down to decimetres.  Beyond 1e-4: only near-contact pairs with a big cylinder whose float32
    # witness GJK stops 1e-4 ABOVE contact while the cores overlap by a few 1e-4 (about 1 pair in 30 000)
    assert tested > 15000 and bad <= 3 and worst < 1e-3, r.stdout[-2000:]
    m = re.search(r"near-contact separated: (\d+) bad (\d+) worst ([0-9.e+-]+)", r.stdout)
    n_near, bad_near, worst_near = int(m.group(1)), int(m.group(2)), float(m.group(3))
    assert n_near > 400 and bad_near <= max(3, n_near // 300) and worst_near < 1e-3, r.stdout[-2000:]


def test_f32_penetration_depth_on_symmetric_poses(tmp_path):
    exe = str(tmp_path / "epa_sym")
    _build("epa_sym.cpp", exe)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    worst = float(re.search(r"^worst ([0-9.e+-]+)", r.stdout, re.M).group(1))
    assert worst < 1e-4, r.stdout[-3000:]


def test_f32_penetration_witnesses_are_points_of_the_shapes(tmp_path):
    """Penetrating pairs of the PRM scene (contact-rich configurations, poses from the float64 FK oracle): the host-compiled
    float32 witness GJK / EPA against the oracle pair by pair -- signed distance, and the two witness points moved onto the
    rounded surfaces must be points of their own shapes (oracle point-to-shape distance <= 3e-4 outside, sagitta inside)."""
    import ctypes as C
    import numpy as np
    from numbotics_b200 import workloads
    from oracle import collision
    lib_path = str(tmp_path / "libgjkhost.so")
    cmd = ["g++", "-O2", "-std=c++17", "-ffp-contract=fast", "-shared", "-fPIC", "-I", os.path.join(H, "fake"),
           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "numbotics_b200", "csrc"),
           os.path.join(H, "gjk_host_lib.cpp"), "-o", lib_path]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    lib = C.CDLL(lib_path)
    arm, cubes = workloads.prm_scene(device=0)          # model only: no device call is made
    so = collision.SceneOracle(arm)
    rng = np.random.default_rng(31)
    n = 1200
    sgn = lambda k: rng.choice([-1.0, 1.0], k)
    q = rng.normal(scale=0.5, size=(n, 7))
    q[:, 1] = rng.uniform(1.2, 2.41, n) * sgn(n)
    q[:, 3] = rng.uniform(2.2, 2.66, n) * sgn(n)
    q[:, 5] = rng.uniform(1.5, 2.23, n) * sgn(n)
    q = q.astype(np.float32).astype(np.float64)
    pd = so.pair_distances(q)
    poses = so.link_poses(q)
    dp = C.POINTER(C.c_double)

    def core(k, m):
        t, d = so.types[k], so.dims[k].copy()
        if t == collision.PRIM_BOX:
            d = d - m
        elif t == collision.PRIM_CYL:
            d[0] -= m; d[1] -= m
        else:
            d[:] = 0
        return np.ascontiguousarray(d)

    bi, pi = np.where(pd < -1e-3)
    assert len(bi) > 1000
    err, off = [], []
    for i, c in zip(bi, pi):
        a, b = so.pairs[c]
        pa = np.ascontiguousarray(poses[i, a])
        pb = np.ascontiguousarray(poses[i, b] if b < so.n_link else so.static_poses[b - so.n_link])
        ma, mb = so.margins[a], so.margins[b]
        out = np.zeros(10)
        lib.host_core_signed_distance(int(so.types[a]), pa.ctypes.data_as(dp), core(a, ma).ctypes.data_as(dp), int(so.types[b]),
                                      pb.ctypes.data_as(dp), core(b, mb).ctypes.data_as(dp), out.ctypes.data_as(dp))
        nrm, A, B = out[1:4], out[4:7], out[7:10]
        err.append(abs(out[0] - ma - mb - pd[i, c]))
        ps, pt = A - ma * nrm, B + mb * nrm
        off.append(max(so.point_to_shape(a, pa[None], ps[None])[0], so.point_to_shape(b, pb[None], pt[None])[0]))
        assert abs(np.linalg.norm(nrm) - 1) < 1e-4 and np.abs((A - B) - out[0] * nrm).max() < 1e-4
    err, off = np.asarray(err), np.asarray(off)
    assert err.max() < 2e-4, err.max()
    assert off.max() < 3e-4, (off.max(), np.argmax(off))
