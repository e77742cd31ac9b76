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
