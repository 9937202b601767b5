"""The reference's BR2 operators, restated (TEST INFRASTRUCTURE).

Each class follows the cited reference lines; arithmetic order is kept so that results are
bit-identical to the reference sources when both run over ``mini_elastica``.
"""
import numpy as np
from numba import njit

from ..mini_elastica import ConstraintBase, FreeJoint, NoForces


# --------------------------------------------------------------------------------------
# rotation helpers: /root/reference/br2/legacy_surface_connection_parallel_rod_numba.py:27-146
# --------------------------------------------------------------------------------------
@njit(cache=True)
def rotation_matrix_single(theta, unit_axis):
    """Row-director Rodrigues matrix (legacy...:27-49): transpose of the usual R(axis, theta)."""
    a0, a1, a2 = unit_axis[0], unit_axis[1], unit_axis[2]
    s = np.sin(theta)
    q = 1.0 - np.cos(theta)
    R = np.empty((3, 3))
    R[0, 0] = 1.0 - q * (a1 * a1 + a2 * a2)
    R[1, 1] = 1.0 - q * (a0 * a0 + a2 * a2)
    R[2, 2] = 1.0 - q * (a0 * a0 + a1 * a1)
    R[0, 1] = s * a2 + q * a0 * a1
    R[1, 0] = -s * a2 + q * a0 * a1
    R[0, 2] = -s * a1 + q * a0 * a2
    R[2, 0] = s * a1 + q * a0 * a2
    R[1, 2] = s * a0 + q * a1 * a2
    R[2, 1] = -s * a0 + q * a1 * a2
    return R


@njit(cache=True)
def log_map_single(R):
    """Axis-angle vector of a row-director rotation (legacy...:52-88), with the reference's
    near-identity (:63-66) and near-pi (:68-82) branches."""
    w = np.empty(3)
    w[0] = R[2, 1] - R[1, 2]
    w[1] = R[0, 2] - R[2, 0]
    w[2] = R[1, 0] - R[0, 1]
    tr = R[0, 0] + R[1, 1] + R[2, 2]
    if np.abs(tr - 3) <= (1e-8 + 1e-5 * 3):
        w *= 0.5 - (tr - 3.0) / 12.0
    elif np.abs(tr + 1) <= (1e-8 + 1e-5):
        a = np.argmax(np.diag(R))
        b = (a + 1) % 3
        c = (a + 2) % 3
        s = np.sqrt(R[a, a] - R[b, b] - R[c, c] + 1)
        v = np.array(
            [s / 2, (1 / (2 * s)) * (R[b, a] + R[a, b]), (1 / (2 * s)) * (R[c, a] + R[a, c])]
        )
        w = np.pi * v / np.sqrt(np.sum(v * v))
    else:
        theta = np.arccos(0.5 * tr - 0.5)
        w *= -0.5 * theta / np.sin(theta + 1e-14)
    return w


@njit(cache=True)
def log_map_batch(R_collection):
    """Batch form (legacy...:91-146): applies ``log_map_single`` to every [:, :, k]."""
    n = R_collection.shape[2]
    out = np.empty((3, n))
    for k in range(n):
        out[:, k] = log_map_single(np.ascontiguousarray(R_collection[:, :, k]))
    return out


# --------------------------------------------------------------------------------------
# fibre actuation: /root/reference/br2/free_custom_systems.py:20-101
# --------------------------------------------------------------------------------------
class FreeBendActuation(NoForces):
    """Straight-fibre (bending) actuation: material-frame couple on the LAST element only
    (free_custom_systems.py:49-57; the uniform variant after the ``return`` is dead code)."""

    def __init__(self, actuation_ref, z_angle, scale, ramp_up_time=0.2):
        super().__init__()
        self.actuation_ref = actuation_ref
        self.z_angle = z_angle
        self.magnitude_scale = scale
        self.ramp_up_time = ramp_up_time

    def apply_torques(self, system, time=0.0):
        ramp = min(1.0, time / self.ramp_up_time)
        magnitude = self.actuation_ref[0] * self.magnitude_scale * ramp
        axis = np.array([np.cos(self.z_angle), np.sin(self.z_angle), 0.0])
        system.external_torques[..., -1] += magnitude * axis


class FreeTwistActuation(NoForces):
    """Angled-fibre (twisting) actuation: P*scale*ramp about d3, shared evenly by all elements
    (free_custom_systems.py:96-101)."""

    def __init__(self, actuation_ref, scale, ramp_up_time=0.2):
        super().__init__()
        self.actuation_ref = actuation_ref
        self.direction = np.array([0.0, 0.0, 1.0])
        self.scale = scale
        self.ramp_up_time = ramp_up_time

    def apply_torques(self, system, time=0.0):
        ramp = min(1.0, time / self.ramp_up_time)
        total = self.actuation_ref[0] * self.scale * self.direction * ramp
        n = system.n_elems
        # (total_i * 1.0) / n for every element, like _batch_product_i_k_to_ik(...)/n_elems
        per_element = (total[:, None] * np.ones((n,))[None, :]) / n
        system.external_torques += per_element


# --------------------------------------------------------------------------------------
# soft-fixed base: /root/reference/br2/free_custom_systems.py:104-185
# --------------------------------------------------------------------------------------
@njit(cache=True)
def _soft_pin_force(base_position, fixed_position, base_velocity, external_forces, k, nu):
    """free_custom_systems.py:151-185.  The reference's two ``np.dot`` calls are written out as
    ordered sums (numba's ``np.dot`` needs SciPy's BLAS, which the GPU box lacks); the golden
    vectors confirm the results are unchanged."""
    d = fixed_position - base_position
    dist = np.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])
    unit = np.zeros(3)
    if not dist <= 1e-8:
        unit[:] = d / dist
    elastic = k * d
    rel_v = -base_velocity
    proj = rel_v[0] * unit[0] + rel_v[1] * unit[1] + rel_v[2] * unit[2]
    damping = -nu * (proj * unit)
    external_forces[..., 0] += elastic + damping


class FreeBaseEndSoftFixed(ConstraintBase):
    """Spring-pins node 0, hard-resets director 0 (orientation torque is unreachable,
    free_custom_systems.py:134-135), zeroes the rates of node/element 0 (:144-149)."""

    def __init__(self, fixed_position, fixed_directors, k, nu, kt, **kwargs):
        super().__init__(**kwargs)
        self.fixed_position = fixed_position
        self.fixed_directors = fixed_directors
        self.k = k
        self.nu = nu
        self.kt = kt

    def constrain_values(self, rod, time):
        _soft_pin_force(
            rod.position_collection[..., 0],
            self.fixed_position,
            rod.velocity_collection[..., 0],
            rod.external_forces,
            self.k,
            self.nu,
        )
        rod.director_collection[..., 0] = self.fixed_directors

    def constrain_rates(self, rod, time):
        rod.velocity_collection[..., 0] = 0.0
        rod.omega_collection[..., 0] = 0.0
        rod.acceleration_collection[..., 0] = 0.0
        rod.alpha_collection[..., 0] = 0.0


# --------------------------------------------------------------------------------------
# serial joint: /root/reference/br2/legacy_surface_connection_parallel_rod_numba.py:346-478
# --------------------------------------------------------------------------------------
@njit(cache=True)
def _tip_to_base_force(k, nu, l1, l2, x1, x2, r1, r2, v1, v2, Q1, Q2, f1, f2):
    """legacy...:386-464.  Spring between the surface points of rod-1's last and rod-2's first
    element, un-projected relative-velocity damping, halves to the two end nodes.  The reference's
    ``Q[..., idx].T @ (l * r)`` is written out as an ordered j = 0,1,2 sum (no BLAS dependency)."""
    n1, n2 = Q1.shape[2], Q2.shape[2]
    arm1 = np.zeros(3)
    arm2 = np.zeros(3)
    total = np.zeros(3)
    for i in range(3):
        for j in range(3):
            arm1[i] += Q1[j, i, n1 - 1] * (l1[j] * r1[0, n1 - 1])
            arm2[i] += Q2[j, i, 0] * (l2[j] * r2[0, 0])
        c1 = 0.5 * (x1[i, n1] + x1[i, n1 - 1])
        c2 = 0.5 * (x2[i, 0] + x2[i, 1])
        gap = (c2 + arm2[i]) - (c1 + arm1[i])
        u1 = 0.5 * (v1[i, n1] + v1[i, n1 - 1])
        u2 = 0.5 * (v2[i, 0] + v2[i, 1])
        total[i] = k * gap + (-nu * (u2 - u1))
        f1[i, n1] += 0.5 * total[i]
        f1[i, n1 - 1] += 0.5 * total[i]
        f2[i, 0] -= 0.5 * total[i]
        f2[i, 1] -= 0.5 * total[i]
    return arm1, arm2, total


class TipToTipStraightJoint(FreeJoint):
    """Glue rod-1's tip to rod-2's base.  ``apply_torques`` is the reference's DEBUG branch:
    it *copies* rod-1's last director / omega / alpha into rod-2's element 0 and returns
    (legacy...:474-478); the alignment-torque code below it never runs."""

    def __init__(self, k, nu, kt, rod1_rd2_local, rod2_rd2_local, stability_check=False):
        super().__init__(k, nu)
        self.kt = kt
        self.rod1_rd2_local = rod1_rd2_local
        self.rod2_rd2_local = rod2_rd2_local
        self.stability_check = stability_check

    def apply_forces(self, rod_one, index_one, rod_two, index_two):
        _tip_to_base_force(
            self.k,
            self.nu,
            self.rod1_rd2_local,
            self.rod2_rd2_local,
            rod_one.position_collection,
            rod_two.position_collection,
            rod_one.radius[None, :],
            rod_two.radius[None, :],
            rod_one.velocity_collection,
            rod_two.velocity_collection,
            rod_one.director_collection,
            rod_two.director_collection,
            rod_one.external_forces,
            rod_two.external_forces,
        )

    def apply_torques(self, rod_one, index_one, rod_two, index_two):
        rod_two.director_collection[..., 0] = rod_one.director_collection[..., -1]
        rod_two.omega_collection[..., 0] = rod_one.omega_collection[..., -1]
        rod_two.alpha_collection[..., 0] = rod_one.alpha_collection[..., -1]


# --------------------------------------------------------------------------------------
# optional ops named by the north star, unreferenced in the reference's own flows
# --------------------------------------------------------------------------------------
class PointForces(NoForces):
    """/root/reference/br2/custom_activation.py:9-81: ramped force on node int(n_nodes*loc)."""

    def __init__(self, force, location, ramp_up_time=1e-3):
        super().__init__()
        self.force = force
        assert ramp_up_time >= 0.0
        self.ramp_up_time = ramp_up_time
        assert 0.0 < location < 1.0
        self.location = location

    def apply_forces(self, system, time=0.0):
        node = int(system.external_forces.shape[-1] * self.location)
        ramp = min(1.0, time / self.ramp_up_time)
        system.external_forces[..., node] += self.force * ramp


class LastEndFixedRod(ConstraintBase):
    """/root/reference/br2/custom_constraint.py:9-88: hard-pins the LAST node / element."""

    def __init__(self, fixed_position, fixed_directors, **kwargs):
        super().__init__(**kwargs)
        self.fixed_position = fixed_position
        self.fixed_directors = fixed_directors

    def constrain_values(self, rod, time):
        rod.position_collection[..., -1] = self.fixed_position
        rod.director_collection[..., -1] = self.fixed_directors

    def constrain_rates(self, rod, time):
        rod.velocity_collection[..., -1] = 0.0
        rod.omega_collection[..., -1] = 0.0
