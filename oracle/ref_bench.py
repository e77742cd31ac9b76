"""Timing harness for the reference's OWN numba FK / Jacobian kernels.  TEST / BENCH INFRASTRUCTURE ONLY.

bench.py's `cpu_baseline` and `--impl reference` legs use this to time numbotics/robots/helpers.py:91-187
(`nb_compute_transformation` + `nb_compute_jacobian`, float64, loaded by oracle/ref_loader.py from the live
checkout or from the copy staged under oracle/_ref/) through the restated Arm wrappers
(`oracle.fkjac.ArmOracle(helpers=...)`, arm.py:369-461) at BASELINE configs[0]'s batch size, B = 4096:

* `single_core_rate`  -- one process, one core (the helpers are serial: no `parallel=True`);
* `RefPool`           -- N worker processes (spawn), one numba instance each, independent slices: the
                         "N-process aggregate" of SURVEY.md section 8(d).

Nothing under numbotics_b200/ imports this module.
"""
import multiprocessing as mp
import os
import time

import numpy as np

B_REF = 4096          # BASELINE configs[0]

_W = {}               # worker state


def _make_oracle(model):
    from . import fkjac, ref_loader
    helpers = ref_loader.load()
    seq, dof, base, com = model
    return fkjac.ArmOracle(seq, dof, base, com, helpers=helpers)


def model_of(arm):
    """Picklable description of the arm: the reference-shaped per-frame cache (arm.py:61), dof, base pose, com offsets."""
    com = {n: np.asarray(l._offset, dtype=np.float64) for n, l in arm._links_from_nodes.items()}
    return (arm._link_joint_sequence, int(arm.dof), np.asarray(arm.base_pose, dtype=np.float64), com)


def _fkj(ora, q, frame):
    T = ora.forward_kinematics(q, frame)
    J = ora.jacobian(q, frame)
    return T, J


def single_core_rate(arm, q, frame="tool_frame", repeats=5):
    """(configs/s, T, J) of FK + Jacobian on one core, best of `repeats` after a warm-up call (which also JIT-compiles)."""
    ora = _make_oracle(model_of(arm))
    q = np.ascontiguousarray(q, dtype=np.float64)
    T, J = _fkj(ora, q, frame)
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        _fkj(ora, q, frame)
        best = min(best, time.perf_counter() - t0)
    return q.shape[0] / best, T, J


def _worker_init(model, lo, hi, frame, seed_base, barrier):
    _W["barrier"] = barrier
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["NUMBA_NUM_THREADS"] = "1"
    ora = _make_oracle(model)
    _W["ora"] = ora
    _W["frame"] = frame
    _W["lo"], _W["hi"] = lo, hi
    _W["seed_base"] = seed_base
    _W["q"] = None


def _worker_step(args):
    idx, batches = args
    _W["barrier"].wait(600)            # one task per worker and step: every worker takes part in every step
    if _W["q"] is None:
        rng = np.random.default_rng(_W["seed_base"] + idx)
        _W["q"] = rng.uniform(_W["lo"], _W["hi"], size=(B_REF, len(_W["lo"])))
        _fkj(_W["ora"], _W["q"], _W["frame"])           # JIT / cache load outside any timed step
    t0 = time.perf_counter()
    for _ in range(batches):
        _fkj(_W["ora"], _W["q"], _W["frame"])
    return time.perf_counter() - t0


class RefPool:
    """N spawn-ed worker processes, each holding its own compiled copy of the reference helpers."""

    def __init__(self, arm, n_procs, frame="tool_frame", seed_base=1000):
        from numbotics_b200 import workloads
        lo, hi = workloads.sample_limits(arm.joint_limits)
        self.n = int(n_procs)
        ctx = mp.get_context("spawn")
        self.pool = ctx.Pool(self.n, initializer=_worker_init,
                             initargs=(model_of(arm), lo, hi, frame, seed_base, ctx.Barrier(self.n)))
        self.step(1)                                   # every worker compiles / loads its cache here

    def step(self, batches_per_proc=2):
        """One step: every worker evaluates `batches_per_proc` batches of 4096 configurations.
        Returns (configs evaluated, wall seconds of the step)."""
        t0 = time.perf_counter()
        self.pool.map(_worker_step, [(i, batches_per_proc) for i in range(self.n)], chunksize=1)
        return self.n * batches_per_proc * B_REF, time.perf_counter() - t0

    def close(self):
        self.pool.terminate()
        self.pool.join()
