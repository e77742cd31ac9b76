"""gcc recipe for the float64 C restatement (oracle/_build/liboracle.so).  Test infrastructure."""
import hashlib
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "nbx_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liboracle.so")
STAMP = os.path.join(OUT_DIR, "liboracle.stamp")
FLAGS = ["-O2", "-fPIC", "-shared", "-fopenmp", "-std=c11", "-fno-fast-math"]


def _digest():
    with open(SRC, "rb") as f:
        return hashlib.sha256(f.read() + " ".join(FLAGS).encode()).hexdigest()


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == _digest():
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = LIB + f".tmp{os.getpid()}"
    r = subprocess.run(["gcc", *FLAGS, SRC, "-o", tmp, "-lm"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"gcc failed on the oracle:\n{r.stdout}")
    os.replace(tmp, LIB)
    with open(STAMP, "w") as f:
        f.write(_digest())
    return LIB


if __name__ == "__main__":
    print(build(force=True))
