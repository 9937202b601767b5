"""Integration path B: lower a simulator that was built with the REFERENCE's own classes.

A maintainer who keeps PyElastica for set-up -- ``br2.free_simulator.FreeAssembly.build`` registering the reference's
``FreeBendActuation`` / ``FreeTwistActuation`` / ``FreeBaseEndSoftFixed`` (/root/reference/br2/free_custom_systems.py:20-217),
``TipToTipStraightJoint`` (legacy_surface_connection_parallel_rod_numba.py:346-478), ``PointForces``, ``LastEndFixedRod`` and
upstream's ``GravityForces`` / ``AnalyticalLinearDamper`` / ``SurfaceJointSideBySide`` on a PyElastica system collection
(free_simulator.py:299-461) -- hands the *finalized* collection to ``lower_simulator`` and gets the same ``br2_program``
tables the product's own builder produces.  Those operator instances have no ``lower_*`` methods; they are recognised by
CLASS NAME and read through the public attributes their constructors set.  Nothing else of them is used: their
``apply_forces`` / ``constrain_values`` ... never run (there is no CPU path), which is why a subclass that overrides one of
those methods is refused instead of being silently replaced by its parent's arithmetic.

Upstream containers read (PyElastica 0.3.1 ``modules``): ``_systems``, ``_ext_forces_torques`` [(rod index, op)],
``_constraints`` [(rod index, op)], ``_dampers`` [(rod index, op)], ``_connections`` [(first rod, second rod, first index,
second index, op)], all as they are after ``finalize()``.
"""
import numpy as np

from . import custom_activation, custom_constraint, free_custom_systems
from . import legacy_surface_connection_parallel_rod_numba as legacy
from ._lowering import lower_collection
from .hooks import ops as _ops

# per kind: class name -> (methods whose arithmetic the kernel reproduces, descriptor factory)
_FORCING = {
    "NoForces": ((), lambda op: _ops.NoForces()),
    "GravityForces": (("apply_forces",), lambda op: _ops.GravityForces(np.asarray(op.acc_gravity, dtype=np.float64))),
    "FreeBendActuation": (("apply_torques",), lambda op: free_custom_systems.FreeBendActuation(
        op.actuation_ref, op.z_angle, op.magnitude_scale, ramp_up_time=op.ramp_up_time)),
    "FreeTwistActuation": (("apply_torques",), lambda op: free_custom_systems.FreeTwistActuation(
        op.actuation_ref, op.scale, ramp_up_time=op.ramp_up_time)),
    "PointForces": (("apply_forces",), lambda op: custom_activation.PointForces(op.force, op.location, ramp_up_time=op.ramp_up_time)),
}
_CONSTRAINTS = {
    "FreeBaseEndSoftFixed": (("constrain_values", "constrain_rates"), lambda op: free_custom_systems.FreeBaseEndSoftFixed(
        np.asarray(op.fixed_position), np.asarray(op.fixed_directors), op.k, op.nu, op.kt)),
    "LastEndFixedRod": (("constrain_values", "constrain_rates"), lambda op: custom_constraint.LastEndFixedRod(
        np.asarray(op.fixed_position), np.asarray(op.fixed_directors))),
}
_JOINTS = {
    "SurfaceJointSideBySide": (("apply_forces", "apply_torques"), lambda op: _surface_joint(op)),
    "TipToTipStraightJoint": (("apply_forces", "apply_torques"), lambda op: legacy.TipToTipStraightJoint(
        op.k, op.nu, op.kt, op.rod1_rd2_local, op.rod2_rd2_local)),
}


def _surface_joint(op):
    joint = _ops.SurfaceJointSideBySide.__new__(_ops.SurfaceJointSideBySide)
    # upstream keeps the two direction vectors transposed (``np.array(vec).T``); they are copied as stored
    joint.k, joint.nu, joint.k_repulsive = np.array(op.k), np.array(op.nu), np.array(op.k_repulsive)
    joint.offset_btw_rods = np.array(op.offset_btw_rods)
    joint.rod_one_direction_vec_in_material_frame = np.array(op.rod_one_direction_vec_in_material_frame)
    joint.rod_two_direction_vec_in_material_frame = np.array(op.rod_two_direction_vec_in_material_frame)
    return joint


class _Damper:
    """``AnalyticalLinearDamper`` read through the two coefficient arrays its constructor computes."""

    def __init__(self, op):
        self.c_t = float(op.translational_damping_coefficient)
        self.ln_c_r = np.log(np.asarray(op.rotational_damping_coefficient, dtype=np.float64))

    def lower_damper(self, ctx, rod_index):
        ctx.set_damper(rod_index, self.c_t, self.ln_c_r)


_DAMPERS = {"AnalyticalLinearDamper": (("dampen_rates",), _Damper)}


def adapt(op, table, kind):
    """The product's descriptor for a foreign operator instance, found by class name along its MRO."""
    if hasattr(op, "lower_" + {"forcing": "forcing", "constraint": "constraint", "joint": "joint", "damper": "damper"}[kind]):
        return op  # already one of this package's descriptors
    for depth, cls in enumerate(type(op).__mro__):
        entry = table.get(cls.__name__)
        if entry is None:
            continue
        methods, factory = entry
        for sub in type(op).__mro__[:depth]:  # classes derived from the recognised one
            overridden = [m for m in ("apply_forces", "apply_torques", "constrain_values", "constrain_rates", "dampen_rates")
                          if m in vars(sub)]
            if overridden:
                raise _ops.NotLowerable(
                    "%s %s.%s overrides %s of %s: its arithmetic would run on the host every step, and the B200 path has no "
                    "CPU fallback" % (kind, type(op).__module__, type(op).__name__, ", ".join(overridden), cls.__name__))
        return factory(op)
    raise _ops.NotLowerable(
        "%s %s.%s has no device lowering (recognised by class name: %s)"
        % (kind, type(op).__module__, type(op).__name__, ", ".join(sorted(table))))


class _ForeignCollection:
    """What ``_lowering.lower_collection`` reads of a collection, filled from the upstream containers."""

    def __init__(self, simulator):
        missing = [name for name in ("_systems",) if not hasattr(simulator, name)]
        if missing:
            raise TypeError("not a PyElastica-style system collection (no %s)" % ", ".join(missing))
        if not getattr(simulator, "_finalize_flag", True):
            raise RuntimeError("finalize() the simulator first: the operator instances are created there")
        self._systems = list(simulator._systems)
        self._forcing = [(int(idx), adapt(op, _FORCING, "forcing")) for idx, op in getattr(simulator, "_ext_forces_torques", ())]
        self._constraint_ops = [(int(idx), adapt(op, _CONSTRAINTS, "constraint")) for idx, op in getattr(simulator, "_constraints", ())]
        self._damper_ops = [(int(idx), adapt(op, _DAMPERS, "damper")) for idx, op in getattr(simulator, "_dampers", ())]
        self._joints = [(int(a), int(b), ia, ib, adapt(op, _JOINTS, "joint"))
                        for a, b, ia, ib, op in getattr(simulator, "_connections", ())]
        # stable per-rod order, as upstream's finalize leaves them (sorted by rod index)
        self._forcing.sort(key=lambda item: item[0])
        self._constraint_ops.sort(key=lambda item: item[0])
        self._damper_ops.sort(key=lambda item: item[0])


def lower_simulator(simulator):
    """``br2._lowering.Program`` (the ``br2_program`` tables of include/br2cuda.h) of a finalized PyElastica-style
    simulator whose operators are the reference's own class instances.  Host only; feed the result to
    ``br2._native.make_program`` / ``br2_create`` (path B)."""
    return lower_collection(_ForeignCollection(simulator))
