"""PyElastica-style system-collection hooks re-implemented for the device path.

``from br2.hooks import *`` provides the names the reference pulls in with ``from elastica import *``
(/root/reference/br2/free_simulator.py:7, environment.py:14), including the re-exported ``defaultdict``.
"""
from collections import defaultdict  # noqa: F401

from .collection import (  # noqa: F401
    BaseSystemCollection,
    CallBacks,
    Connections,
    Constraints,
    Damping,
    Forcing,
)
from .ops import (  # noqa: F401
    AnalyticalLinearDamper,
    CallBackBaseClass,
    ConstraintBase,
    DamperBase,
    FreeJoint,
    GravityForces,
    NoForces,
    NotLowerable,
    SurfaceJointSideBySide,
    get_connection_vector_straight_straight_rod,
)
from .rod import CosseratRod, RodBase  # noqa: F401
from .stepper import PositionVerlet, extend_stepper_interface  # noqa: F401

__all__ = [
    "defaultdict", "BaseSystemCollection", "CallBacks", "Connections", "Constraints", "Damping", "Forcing",
    "AnalyticalLinearDamper", "CallBackBaseClass", "ConstraintBase", "DamperBase", "FreeJoint",
    "GravityForces", "NoForces", "NotLowerable", "SurfaceJointSideBySide",
    "get_connection_vector_straight_straight_rod", "CosseratRod", "RodBase", "PositionVerlet",
    "extend_stepper_interface",
]
