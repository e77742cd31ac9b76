"""Load the reference's own numba FK / Jacobian kernels by file path.  TEST INFRASTRUCTURE ONLY.

`import numbotics` is impossible here (numbotics/physics/__init__.py:31-35 needs PyBullet), but
numbotics/robots/helpers.py only needs `numbotics.physics.Constraint`, whose values are the
PyBullet JOINT_* integers (numbotics/physics/constraint.py:11-15).  A stub module supplies it.

Two places to load it from:

* /root/reference (the build container): used by tests/golden/make_golden.py and the live-reference test;
* oracle/_ref/ref_helpers.py: a byte-for-byte staging of that one file made by `stage()` -- called from
  `__graft_entry__.build()` while /root/reference is visible.  oracle/_ref/ is git-ignored (reference source
  never enters the history) but travels to the GPU box with the working tree, so that bench.py's
  `cpu_baseline` / `--impl reference` legs can time the reference's OWN numba kernels on the box's host cores
  (`kind = "reference-numba"`).  Only those two bench legs and tests/ may call this module.
"""
import enum
import importlib.util
import os
import sys
import types

import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("NUMBOTICS_REFERENCE", "/root/reference")
LIVE_PATH = os.path.join(REFERENCE_ROOT, "numbotics", "robots", "helpers.py")
STAGED_DIR = os.path.join(HERE, "_ref")
STAGED_PATH = os.path.join(STAGED_DIR, "ref_helpers.py")


def stage() -> str | None:
    """Copy the reference's helpers.py (unmodified) into oracle/_ref/ when the checkout is visible.
    Returns the staged path, or None when there is nothing to stage from."""
    if not os.path.exists(LIVE_PATH):
        return STAGED_PATH if os.path.exists(STAGED_PATH) else None
    os.makedirs(STAGED_DIR, exist_ok=True)
    with open(LIVE_PATH, "rb") as f:
        data = f.read()
    if not os.path.exists(STAGED_PATH) or open(STAGED_PATH, "rb").read() != data:
        tmp = STAGED_PATH + f".tmp{os.getpid()}"
        with open(tmp, "wb") as f:
            f.write(data)
        os.replace(tmp, STAGED_PATH)
    with open(os.path.join(STAGED_DIR, "README"), "w") as f:
        f.write("ref_helpers.py = numbotics/robots/helpers.py of the reference checkout, staged unmodified by\n"
                "oracle/ref_loader.py::stage() (git-ignored; timed by bench.py's reference-numba leg only).\n")
    return STAGED_PATH


def helpers_path() -> str | None:
    """The live checkout when it is visible, else the staged copy."""
    if os.path.exists(LIVE_PATH):
        return LIVE_PATH
    if os.path.exists(STAGED_PATH):
        return STAGED_PATH
    return None


def available() -> bool:
    if helpers_path() is None:
        return False
    try:
        import numba  # noqa: F401
    except Exception:
        return False
    return True


def load():
    """Return the reference helpers module (compiles with numba on first use, then cached)."""
    if "ref_helpers" in sys.modules:
        return sys.modules["ref_helpers"]
    if not available():
        raise RuntimeError("reference checkout or numba not available")
    # numba cache=True would otherwise write next to the (read-only) reference source
    os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join(os.path.expanduser("~"), ".cache", "nbx_ref_numba"))
    HELPERS_PATH = helpers_path()

    class Constraint(enum.Enum):
        FIXED = 4
        REVOLUTE = 0
        PRISMATIC = 1
        SPHERICAL = 2

    created = []
    if "numbotics" not in sys.modules:
        pkg = types.ModuleType("numbotics")
        pkg.__path__ = []
        sys.modules["numbotics"] = pkg
        created.append("numbotics")
    if "numbotics.physics" not in sys.modules:
        phys = types.ModuleType("numbotics.physics")
        phys.Constraint = Constraint
        sys.modules["numbotics.physics"] = phys
        created.append("numbotics.physics")
    try:
        spec = importlib.util.spec_from_file_location("ref_helpers", HELPERS_PATH)
        mod = importlib.util.module_from_spec(spec)
        sys.modules["ref_helpers"] = mod   # before exec: numba's cache needs a resolvable module
        spec.loader.exec_module(mod)
    except Exception:
        sys.modules.pop("ref_helpers", None)
        raise
    finally:
        for name in created:
            sys.modules.pop(name, None)
    return mod
