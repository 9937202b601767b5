"""mini_elastica -- TEST INFRASTRUCTURE, not product code.

A CPU restatement (numpy + numba, fp64, single thread, one dispatch per operator like the
original) of the parts of PyElastica 0.3.1.post1 that the reference's hot path runs through.
PyElastica is a third-party dependency of the reference (``pyelastica >=0.3.1`` at
/root/reference/pyproject.toml:11, locked to 0.3.1.post1 at /root/reference/poetry.lock:1784-1785);
it is not vendored under /root/reference, is not installed here and cannot be installed
(no network; needs python<3.11 / numpy<2).  The algorithm restated is the published one
(Gazzola et al. 2018, RSOS, eqs. 3.3-3.13; PyElastica 0.3.1 sources) as summarised in
SURVEY.md Appendix A, and every place the reference calls into it is cited in the module
docstrings.

**Parity unpinned** for the upstream arithmetic: no golden vectors exist for it.  The module
layout mirrors ``elastica`` so that ``install_as_elastica()`` can alias it into ``sys.modules``
and the reference's *real* ``br2`` sources run on top of it (tests/golden/make_golden.py).

The public surface mirrors what ``from elastica import *`` gives the reference, including the
re-exported ``defaultdict`` that /root/reference/br2/free_simulator.py:81,394 relies on.
"""
from collections import defaultdict  # noqa: F401  (re-exported on purpose, see docstring)

from . import (  # noqa: F401
    _calculus,
    _linalg,
    _rotations,
    boundary_conditions,
    callback_functions,
    dissipation,
    external_forces,
    interaction,
    joint,
    modules,
    restart,
    rod,
    timestepper,
    utils,
)
from .boundary_conditions import ConstraintBase, FreeBC  # noqa: F401
from .callback_functions import CallBackBaseClass  # noqa: F401
from .dissipation import AnalyticalLinearDamper, DamperBase  # noqa: F401
from .external_forces import GravityForces, NoForces, UniformTorques  # noqa: F401
from .joint import FreeJoint  # noqa: F401
from .modules import (  # noqa: F401
    BaseSystemCollection,
    CallBacks,
    Connections,
    Constraints,
    Damping,
    Forcing,
)
from .rod import CosseratRod, RodBase  # noqa: F401
from .timestepper import PositionVerlet, extend_stepper_interface, integrate  # noqa: F401
from .experimental.connection_contact_joint import parallel_connection  # noqa: F401


def install_as_elastica():
    """Alias this package as ``elastica`` in ``sys.modules`` (golden generation only)."""
    import sys

    me = sys.modules[__name__]
    sys.modules["elastica"] = me
    for sub in (
        "_calculus",
        "_linalg",
        "_rotations",
        "boundary_conditions",
        "callback_functions",
        "dissipation",
        "external_forces",
        "interaction",
        "joint",
        "modules",
        "restart",
        "rod",
        "timestepper",
        "utils",
    ):
        sys.modules["elastica." + sub] = getattr(me, sub)
    # stale private module paths still imported by unreferenced reference files
    # (/root/reference/br2/custom_activation.py:6, custom_constraint.py:3)
    sys.modules["elastica._external_forces"] = external_forces
    sys.modules["elastica._boundary_conditions"] = boundary_conditions
    from . import experimental
    from .experimental import connection_contact_joint

    sys.modules["elastica.experimental"] = experimental
    sys.modules["elastica.experimental.connection_contact_joint"] = connection_contact_joint
    sys.modules[
        "elastica.experimental.connection_contact_joint.parallel_connection"
    ] = parallel_connection
    return me
