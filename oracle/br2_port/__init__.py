"""br2_port -- TEST INFRASTRUCTURE: CPU restatement of the reference's own ``br2`` hot-path code
(operators, assembly builder, run loop) on top of ``oracle.mini_elastica``.

Pinned by tests/golden/*.npz|json, which were produced by executing the reference's real
sources (/root/reference/br2/*.py) over ``mini_elastica`` -- see tests/golden/make_golden.py.
"""
from .ops import (  # noqa: F401
    FreeBendActuation,
    FreeTwistActuation,
    FreeBaseEndSoftFixed,
    TipToTipStraightJoint,
    PointForces,
    LastEndFixedRod,
    rotation_matrix_single,
    log_map_single,
    log_map_batch,
)
from .assembly import BR2Simulator, FreeAssembly, FreeCallback  # noqa: F401
from .env import OracleEnvironment, count_steps  # noqa: F401
