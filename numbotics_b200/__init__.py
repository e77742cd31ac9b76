"""numbotics_b200 -- B200-native batched FK / Jacobian / collision-validity path for Numbotics.

Only the data-parallel hot path of landonclark97/numbotics is implemented
(SURVEY.md section 8); the package layout mirrors the reference's so call sites port
by changing the import root:

    numbotics.physics   -> numbotics_b200.physics    (model + obstacle types, URDF loader)
    numbotics.robots    -> numbotics_b200.robots     (Arm, nb_compute_* helpers)
    numbotics.planning  -> numbotics_b200.planning   (DiscreteConnector, ConnectorParams, unit_bspline)

All arithmetic runs in hand-written sm_100a CUDA behind the C-ABI of include/nbx.h
(numbotics_b200/_lib/libnbx.so).  There is no CPU fallback.
"""
__version__ = "0.1.0"

from . import math, physics  # noqa: F401


def __getattr__(name):
    # robots / planning import torch lazily through these attributes
    if name in ("robots", "planning", "dist"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
