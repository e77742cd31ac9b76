"""Load the reference's own numba FK / Jacobian kernels by file path.  TEST INFRASTRUCTURE ONLY.

`import numbotics` is impossible here (numbotics/physics/__init__.py:31-35 needs PyBullet), but
numbotics/robots/helpers.py only needs `numbotics.physics.Constraint`, whose values are the
PyBullet JOINT_* integers (numbotics/physics/constraint.py:11-15).  A stub module supplies it.

Works only where /root/reference exists (the build container).  It never travels to the GPU
box: the vectors it produces are committed under tests/golden/ by tests/golden/make_golden.py.
"""
import enum
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("NUMBOTICS_REFERENCE", "/root/reference")
HELPERS_PATH = os.path.join(REFERENCE_ROOT, "numbotics", "robots", "helpers.py")


def available() -> bool:
    if not os.path.exists(HELPERS_PATH):
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
