"""Operator descriptors for the PyElastica-style hooks.

In the reference every operator is a Python object whose ``apply_forces`` / ``constrain_values`` /
``dampen_rates`` ... methods are dispatched once per step on the host.  Here an operator is a
*descriptor*: it keeps its parameters and knows how to lower itself to a row of the device tables
(include/br2cuda.h).  ``finalize()`` refuses operators without a lowering -- the B200 path has no
CPU fallback to run arbitrary Python per step.

Upstream classes mirrored here (call sites in /root/reference/br2/free_simulator.py):
``NoForces`` / ``GravityForces`` (:334-338), ``ConstraintBase``, ``AnalyticalLinearDamper`` (:314-320),
``FreeJoint``, experimental ``SurfaceJointSideBySide`` + ``get_connection_vector_straight_straight_rod``
(:404-435), ``CallBackBaseClass`` (:42).  Semantics: SURVEY.md Appendix A.5-A.7.
"""
import numpy as np

# enum values of include/br2cuda.h
FORCE_GRAVITY, FORCE_TWIST, FORCE_BEND, FORCE_POINT = 0, 1, 2, 3
CONS_SOFT_FIXED_BASE, CONS_LAST_END_FIXED = 0, 1
JOINT_SURFACE, JOINT_TIP_TO_BASE = 0, 1


class NotLowerable(TypeError):
    pass


def _no_lowering(op, kind):
    raise NotLowerable(
        "%s %s.%s has no device lowering: the B200 path integrates on the GPU only and cannot call "
        "Python operators per step (no CPU fallback). Supported: see include/br2cuda.h."
        % (kind, type(op).__module__, type(op).__name__)
    )


# ------------------------------------------------------------------------------------ forcing
class NoForces:
    """Base of forcing operators.  A bare ``NoForces`` lowers to nothing."""

    def __init__(self):
        pass

    def lower_forcing(self, ctx, rod_index):
        if type(self) is NoForces:
            return
        _no_lowering(self, "forcing")


class GravityForces(NoForces):
    def __init__(self, acc_gravity=np.array([0.0, -9.80665, 0.0])):
        super().__init__()
        self.acc_gravity = np.asarray(acc_gravity, dtype=np.float64)

    def lower_forcing(self, ctx, rod_index):
        ctx.add_forcing(rod_index, FORCE_GRAVITY, d=list(self.acc_gravity))


# ------------------------------------------------------------------------------------ constraints
class ConstraintBase:
    def __init__(self, *args, **kwargs):
        self._system = kwargs.get("_system", None)
        self._constrained_position_idx = np.array(kwargs.get("constrained_position_idx", []), dtype=int)
        self._constrained_director_idx = np.array(kwargs.get("constrained_director_idx", []), dtype=int)

    @property
    def system(self):
        return self._system

    @property
    def constrained_position_idx(self):
        return self._constrained_position_idx

    @property
    def constrained_director_idx(self):
        return self._constrained_director_idx

    def lower_constraint(self, ctx, rod_index):
        _no_lowering(self, "constraint")


# ------------------------------------------------------------------------------------ dampers
class DamperBase:
    def __init__(self, *args, **kwargs):
        self._system = kwargs["_system"]

    @property
    def system(self):
        return self._system

    def lower_damper(self, ctx, rod_index):
        _no_lowering(self, "damper")


class AnalyticalLinearDamper(DamperBase):
    """v *= exp(-nu dt);  w_c *= exp(-nu dt m_elem / J_c) ** e   (SURVEY.md Appendix A.5)."""

    def __init__(self, damping_constant, time_step, **kwargs):
        super().__init__(**kwargs)
        rod = self._system
        nodal_mass = rod.mass
        self.translational_damping_coefficient = np.exp(-damping_constant * time_step)
        element_mass = 0.5 * (nodal_mass[1:] + nodal_mass[:-1])
        element_mass[0] += 0.5 * nodal_mass[0]
        element_mass[-1] += 0.5 * nodal_mass[-1]
        inv_j = np.diagonal(rod.inv_mass_second_moment_of_inertia).T
        self.rotational_damping_coefficient = np.exp(-damping_constant * time_step * element_mass * inv_j)

    def lower_damper(self, ctx, rod_index):
        ctx.set_damper(rod_index, float(self.translational_damping_coefficient),
                       np.log(self.rotational_damping_coefficient))


# ------------------------------------------------------------------------------------ joints
class FreeJoint:
    def __init__(self, k, nu):
        self.k = k
        self.nu = nu

    def lower_joint(self, ctx, rod_one, index_one, rod_two, index_two):
        _no_lowering(self, "joint")


def get_connection_vector_straight_straight_rod(rod_one, rod_two, rod_one_idx, rod_two_idx):
    """Per-element unit vectors rod1->rod2 and rod2->rod1 in each rod's material frame, and the
    surface gap at rest (upstream helper, SURVEY.md Appendix A.7 'Setup')."""
    a0, a1 = rod_one_idx
    b0, b1 = rod_two_idx
    x1, x2 = rod_one.position_collection, rod_two.position_collection
    c1 = (0.5 * (x1[..., 1:] + x1[..., :-1]))[:, a0:a1]
    c2 = (0.5 * (x2[..., 1:] + x2[..., :-1]))[:, b0:b1]
    sep = c2 - c1
    sep_norm = np.sqrt(sep[0] * sep[0] + sep[1] * sep[1] + sep[2] * sep[2])
    unit = sep / sep_norm
    Q1 = rod_one.director_collection[:, :, a0:a1]
    Q2 = rod_two.director_collection[:, :, b0:b1]
    # sum over the lab axis in index order, like upstream's accumulate-in-place matvec
    dv1 = np.zeros_like(unit)
    dv2 = np.zeros_like(unit)
    for j in range(3):
        dv1 += Q1[:, j, :] * unit[j]
        dv2 += Q2[:, j, :] * (-unit)[j]
    offset = sep_norm - (rod_one.radius[a0:a1] + rod_two.radius[b0:b1])
    return dv1, dv2, offset


class SurfaceJointSideBySide(FreeJoint):
    """Spring/damper/repulsion between the facing surface points of element ``index_one`` of rod one
    and element ``index_two`` of rod two."""

    def __init__(self, k, nu, k_repulsive, rod_one_direction_vec_in_material_frame,
                 rod_two_direction_vec_in_material_frame, offset_btw_rods, **kwargs):
        super().__init__(np.array(k), np.array(nu))
        self.k_repulsive = np.array(k_repulsive)
        self.offset_btw_rods = np.array(offset_btw_rods)
        self.rod_one_direction_vec_in_material_frame = np.array(rod_one_direction_vec_in_material_frame).T
        self.rod_two_direction_vec_in_material_frame = np.array(rod_two_direction_vec_in_material_frame).T

    def lower_joint(self, ctx, rod_one, index_one, rod_two, index_two):
        if np.ndim(index_one) != 0 or np.ndim(index_two) != 0:
            raise NotLowerable("SurfaceJointSideBySide lowers scalar element indices only "
                               "(one joint object per element, as the reference registers them)")
        d = [float(self.k), float(self.nu), float(self.k_repulsive)]
        d += [float(c) for c in np.ravel(self.rod_one_direction_vec_in_material_frame)]
        d += [float(c) for c in np.ravel(self.rod_two_direction_vec_in_material_frame)]
        d += [float(self.offset_btw_rods)]
        ctx.add_joint(JOINT_SURFACE, rod_one, int(index_one), rod_two, int(index_two), d, overwrite=False, torque=True)


# ------------------------------------------------------------------------------------ callbacks
class CallBackBaseClass:
    """Diagnostics run on the host, on snapshots the engine copies back at capture steps."""

    def __init__(self):
        pass

    def make_callback(self, system, time, current_step):
        pass

    def capture_steps(self):
        """``None`` = may fire on any step (forces a launch per step); otherwise the period."""
        return getattr(self, "every", None)
