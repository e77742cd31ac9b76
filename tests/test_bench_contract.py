"""bench.py contract, CPU side: the reference arm (`--impl reference`) runs without a GPU (it times the reference's own
numba kernels staged under oracle/_ref/ -- or the oracle port when numba is missing -- on the host cores) and prints one
JSON line with the keys the driver reads; under a 2-rank launch only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def _json_lines(text):
    return [json.loads(l) for l in text.splitlines() if l.startswith("{")]


def test_reference_arm_line():
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "2", "--warmup", "3"],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = _json_lines(r.stdout)
    assert len(lines) == 1
    d = lines[0]
    assert KEYS <= set(d)
    assert d["impl"] == "reference" and d["metric"] == "fk_jacobian_configs_per_sec" and d["unit"] == "configs/s"
    assert d["steps"] == 2 and d["warmup"] >= 3 and d["higher_is_better"] is True
    assert d["value"] > 1e4
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert cb["port"]["kind"] == "port" and cb["port"]["value"] > 1e5
    # the config is the GPU arm's own (the driver compares the two dicts): the bounded sample lives in cpu_baseline.sample
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.bench_config(1)
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and "workload" in d["config"]


def test_reference_arm_two_ranks_print_once():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29547", "bench.py", "--impl", "reference",
                        "--gpus", "2", "--steps", "1", "--warmup", "3"],
                       cwd=ROOT, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = _json_lines(r.stdout)
    assert len(lines) == 1 and lines[0]["n_gpus"] == 2 and lines[0]["impl"] == "reference"
    assert lines[0]["cpu_baseline"]["cores"] >= 1
