"""Cosserat rod state, straight-rod allocation and internal force/torque kernels.

TEST INFRASTRUCTURE -- oracle restatement of ``elastica.rod.cosserat_rod`` /
``elastica.rod.factory_function`` (PyElastica 0.3.1.post1, not vendored in /root/reference).
Reference call site: ``CosseratRod.straight_rod(**rod_spec)`` at
/root/reference/br2/free_simulator.py:293; attributes the reference reads are listed in
SURVEY.md section 8(b).  Algorithm: SURVEY.md Appendix A.2 (allocation) and A.3 (kernels).

Upstream concatenates rods into ghost-padded memory blocks; real entries equal the
stand-alone-rod result, so this oracle steps each rod on its own arrays.
"""
import numpy as np
from numba import njit

from ._linalg import _batch_cross, _batch_dot, _batch_matvec, _batch_norm
from ._calculus import (
    position_difference_kernel,
    position_average,
    difference_kernel_for_block_structure,
    trapezoidal_for_block_structure,
)
from ._rotations import _inv_rotate

# Kaneko (1975) shear correction used from PyElastica 0.3.0 on (was 4/3 in 0.2.x).
ALPHA_C = 27.0 / 28.0
# Guard added to element lengths upstream ("ghost lengths are 0").
LENGTH_GUARD = 1e-14


class RodBase:
    """Marker base class (``elastica.rod.RodBase``)."""

    REQUISITE_MODULES = []


# --------------------------------------------------------------------------------------
# kernels
# --------------------------------------------------------------------------------------
@njit(cache=True)
def _compute_geometry_from_state(position_collection, volume, lengths, tangents, radius):
    position_diff = position_difference_kernel(position_collection)
    lengths[:] = _batch_norm(position_diff) + LENGTH_GUARD
    for k in range(lengths.shape[0]):
        tangents[0, k] = position_diff[0, k] / lengths[k]
        tangents[1, k] = position_diff[1, k] / lengths[k]
        tangents[2, k] = position_diff[2, k] / lengths[k]
        radius[k] = np.sqrt(volume[k] / lengths[k] / np.pi)


@njit(cache=True)
def _compute_all_dilatations(
    position_collection,
    volume,
    lengths,
    tangents,
    radius,
    dilatation,
    rest_lengths,
    rest_voronoi_lengths,
    voronoi_dilatation,
):
    _compute_geometry_from_state(position_collection, volume, lengths, tangents, radius)
    for k in range(lengths.shape[0]):
        dilatation[k] = lengths[k] / rest_lengths[k]
    voronoi_lengths = position_average(lengths)
    for k in range(voronoi_lengths.shape[0]):
        voronoi_dilatation[k] = voronoi_lengths[k] / rest_voronoi_lengths[k]


@njit(cache=True)
def _compute_dilatation_rate(
    position_collection, velocity_collection, lengths, rest_lengths, dilatation_rate
):
    r_dot_v = _batch_dot(position_collection, velocity_collection)
    r_plus_one_dot_v = _batch_dot(
        position_collection[..., 1:], velocity_collection[..., :-1]
    )
    r_dot_v_plus_one = _batch_dot(
        position_collection[..., :-1], velocity_collection[..., 1:]
    )
    for k in range(lengths.shape[0]):
        dilatation_rate[k] = (
            (r_dot_v[k] + r_dot_v[k + 1] - r_dot_v_plus_one[k] - r_plus_one_dot_v[k])
            / lengths[k]
            / rest_lengths[k]
        )


@njit(cache=True)
def _compute_shear_stretch_strains(
    position_collection,
    volume,
    lengths,
    tangents,
    radius,
    rest_lengths,
    rest_voronoi_lengths,
    dilatation,
    voronoi_dilatation,
    director_collection,
    sigma,
):
    _compute_all_dilatations(
        position_collection,
        volume,
        lengths,
        tangents,
        radius,
        dilatation,
        rest_lengths,
        rest_voronoi_lengths,
        voronoi_dilatation,
    )
    z_vector = np.array([0.0, 0.0, 1.0]).reshape(3, -1)
    sigma[:] = dilatation * _batch_matvec(director_collection, tangents) - z_vector


@njit(cache=True)
def _compute_internal_shear_stretch_stresses_from_model(
    position_collection,
    volume,
    lengths,
    tangents,
    radius,
    rest_lengths,
    rest_voronoi_lengths,
    dilatation,
    voronoi_dilatation,
    director_collection,
    sigma,
    rest_sigma,
    shear_matrix,
    internal_stress,
):
    _compute_shear_stretch_strains(
        position_collection,
        volume,
        lengths,
        tangents,
        radius,
        rest_lengths,
        rest_voronoi_lengths,
        dilatation,
        voronoi_dilatation,
        director_collection,
        sigma,
    )
    internal_stress[:] = _batch_matvec(shear_matrix, sigma - rest_sigma)


@njit(cache=True)
def _compute_internal_forces(
    position_collection,
    volume,
    lengths,
    tangents,
    radius,
    rest_lengths,
    rest_voronoi_lengths,
    dilatation,
    voronoi_dilatation,
    director_collection,
    sigma,
    rest_sigma,
    shear_matrix,
    internal_stress,
    internal_forces,
    ghost_elems_idx,
):
    _compute_internal_shear_stretch_stresses_from_model(
        position_collection,
        volume,
        lengths,
        tangents,
        radius,
        rest_lengths,
        rest_voronoi_lengths,
        dilatation,
        voronoi_dilatation,
        director_collection,
        sigma,
        rest_sigma,
        shear_matrix,
        internal_stress,
    )
    # Q^T n_L / e
    blocksize = internal_stress.shape[1]
    cosserat_internal_stress = np.zeros((3, blocksize))
    for i in range(3):
        for j in range(3):
            for k in range(blocksize):
                cosserat_internal_stress[i, k] += (
                    director_collection[j, i, k] * internal_stress[j, k]
                )
    cosserat_internal_stress /= dilatation
    internal_forces[:] = difference_kernel_for_block_structure(
        cosserat_internal_stress, ghost_elems_idx
    )


@njit(cache=True)
def _compute_bending_twist_strains(director_collection, rest_voronoi_lengths, kappa):
    temp = _inv_rotate(director_collection)
    for k in range(rest_voronoi_lengths.shape[0]):
        kappa[0, k] = temp[0, k] / rest_voronoi_lengths[k]
        kappa[1, k] = temp[1, k] / rest_voronoi_lengths[k]
        kappa[2, k] = temp[2, k] / rest_voronoi_lengths[k]


@njit(cache=True)
def _compute_internal_bending_twist_stresses_from_model(
    director_collection,
    rest_voronoi_lengths,
    internal_couple,
    bend_matrix,
    kappa,
    rest_kappa,
):
    _compute_bending_twist_strains(director_collection, rest_voronoi_lengths, kappa)
    blocksize = kappa.shape[1]
    temp = np.empty((3, blocksize))
    for i in range(3):
        for k in range(blocksize):
            temp[i, k] = kappa[i, k] - rest_kappa[i, k]
    internal_couple[:] = _batch_matvec(bend_matrix, temp)


@njit(cache=True)
def _compute_internal_torques(
    position_collection,
    velocity_collection,
    tangents,
    lengths,
    rest_lengths,
    director_collection,
    rest_voronoi_lengths,
    bend_matrix,
    rest_kappa,
    kappa,
    voronoi_dilatation,
    mass_second_moment_of_inertia,
    omega_collection,
    internal_stress,
    internal_couple,
    dilatation,
    dilatation_rate,
    internal_torques,
    ghost_voronoi_idx,
):
    _compute_internal_bending_twist_stresses_from_model(
        director_collection,
        rest_voronoi_lengths,
        internal_couple,
        bend_matrix,
        kappa,
        rest_kappa,
    )
    _compute_dilatation_rate(
        position_collection, velocity_collection, lengths, rest_lengths, dilatation_rate
    )

    voronoi_dilatation_inv_cube_cached = 1.0 / voronoi_dilatation ** 3
    # Delta^h (tau_L / E^3)
    bend_twist_couple_2D = difference_kernel_for_block_structure(
        internal_couple * voronoi_dilatation_inv_cube_cached, ghost_voronoi_idx
    )
    # A^h ((kappa x tau_L) D / E^3)
    bend_twist_couple_3D = trapezoidal_for_block_structure(
        _batch_cross(kappa, internal_couple)
        * rest_voronoi_lengths
        * voronoi_dilatation_inv_cube_cached,
        ghost_voronoi_idx,
    )
    # (Q t x n_L) l_hat
    shear_stretch_couple = (
        _batch_cross(_batch_matvec(director_collection, tangents), internal_stress)
        * rest_lengths
    )
    # J w / e, reused by the transport and the unsteady-dilatation terms
    J_omega_upon_e = (
        _batch_matvec(mass_second_moment_of_inertia, omega_collection) / dilatation
    )
    lagrangian_transport = _batch_cross(J_omega_upon_e, omega_collection)
    unsteady_dilatation = J_omega_upon_e * dilatation_rate / dilatation

    blocksize = internal_torques.shape[1]
    for i in range(3):
        for k in range(blocksize):
            internal_torques[i, k] = (
                bend_twist_couple_2D[i, k]
                + bend_twist_couple_3D[i, k]
                + shear_stretch_couple[i, k]
                + lagrangian_transport[i, k]
                + unsteady_dilatation[i, k]
            )


@njit(cache=True)
def _update_accelerations(
    acceleration_collection,
    internal_forces,
    external_forces,
    mass,
    alpha_collection,
    inv_mass_second_moment_of_inertia,
    internal_torques,
    external_torques,
    dilatation,
):
    blocksize_acc = internal_forces.shape[1]
    blocksize_alpha = internal_torques.shape[1]
    for i in range(3):
        for k in range(blocksize_acc):
            acceleration_collection[i, k] = (
                internal_forces[i, k] + external_forces[i, k]
            ) / mass[k]
    alpha_collection *= 0.0
    for i in range(3):
        for j in range(3):
            for k in range(blocksize_alpha):
                alpha_collection[i, k] += (
                    inv_mass_second_moment_of_inertia[i, j, k]
                    * (internal_torques[j, k] + external_torques[j, k])
                ) * dilatation[k]


@njit(cache=True)
def _zeroed_out_external_forces_and_torques(external_forces, external_torques):
    external_forces *= 0.0
    external_torques *= 0.0


# --------------------------------------------------------------------------------------
# allocation
# --------------------------------------------------------------------------------------
def allocate_straight_rod(
    n_elements,
    start,
    direction,
    normal,
    base_length,
    base_radius,
    density,
    youngs_modulus,
    shear_modulus=None,
    alpha_c=ALPHA_C,
):
    """Arrays of a straight, uniform, unstrained rod at rest."""
    n = int(n_elements)
    start = np.asarray(start, dtype=np.float64)
    direction = np.asarray(direction, dtype=np.float64)
    normal = np.asarray(normal, dtype=np.float64)
    end = start + direction * base_length
    position = np.zeros((3, n + 1))
    for i in range(3):
        position[i, :] = np.linspace(start[i], end[i], n + 1)

    position_diff = position[:, 1:] - position[:, :-1]
    rest_lengths = _batch_norm(position_diff)
    tangents = position_diff / rest_lengths
    normal = normal / np.linalg.norm(normal)

    directors = np.zeros((3, 3, n))
    normal_collection = np.repeat(normal[:, np.newaxis], n, axis=1)
    assert np.allclose(_batch_dot(normal_collection, tangents), 0.0, atol=1e-12), (
        "rod normal and tangent must be perpendicular"
    )
    directors[0, :, :] = normal_collection
    directors[1, :, :] = _batch_cross(tangents, normal_collection)
    directors[2, :, :] = tangents

    radius = np.zeros(n)
    radius[:] = base_radius
    density_array = np.zeros(n)
    density_array[:] = density

    A0 = np.pi * radius * radius
    I0_1 = A0 * A0 / (4.0 * np.pi)
    I0_2 = I0_1
    I0_3 = 2.0 * I0_2
    I0 = np.array([I0_1, I0_2, I0_3]).transpose()  # (n, 3)

    mass_second_moment_of_inertia = np.zeros((3, 3, n))
    msm_temp = np.einsum("ij,i->ij", I0, density_array * rest_lengths)
    for k in range(n):
        np.fill_diagonal(mass_second_moment_of_inertia[:, :, k], msm_temp[k, :])
    inv_mass_second_moment_of_inertia = np.zeros((3, 3, n))
    for k in range(n):
        inv_mass_second_moment_of_inertia[:, :, k] = np.linalg.inv(
            mass_second_moment_of_inertia[:, :, k]
        )

    if shear_modulus is None:
        # >=0.3.1 fallback, Poisson ratio 0.5 (not exercised: the rod library supplies G,
        # /root/reference/sample_database/sample_rod_library.json:13)
        shear_modulus = youngs_modulus / (2.0 * (1.0 + 0.5))

    shear_matrix = np.zeros((3, 3, n))
    for k in range(n):
        np.fill_diagonal(
            shear_matrix[:, :, k],
            [
                alpha_c * shear_modulus * A0[k],
                alpha_c * shear_modulus * A0[k],
                youngs_modulus * A0[k],
            ],
        )

    bend_matrix_elem = np.zeros((3, 3, n))
    for k in range(n):
        np.fill_diagonal(
            bend_matrix_elem[:, :, k],
            [
                youngs_modulus * I0_1[k],
                youngs_modulus * I0_2[k],
                shear_modulus * I0_3[k],
            ],
        )
    bend_matrix = (
        bend_matrix_elem[..., 1:] * rest_lengths[1:]
        + bend_matrix_elem[..., :-1] * rest_lengths[0:-1]
    ) / (rest_lengths[1:] + rest_lengths[:-1])

    volume = np.pi * radius ** 2 * rest_lengths
    mass = np.zeros(n + 1)
    mass[:-1] += 0.5 * density_array * volume
    mass[1:] += 0.5 * density_array * volume

    rest_voronoi_lengths = 0.5 * (rest_lengths[1:] + rest_lengths[:-1])

    return dict(
        n_elements=n,
        position=position,
        velocity=np.zeros((3, n + 1)),
        omega=np.zeros((3, n)),
        acceleration=np.zeros((3, n + 1)),
        angular_acceleration=np.zeros((3, n)),
        directors=directors,
        radius=radius,
        mass_second_moment_of_inertia=mass_second_moment_of_inertia,
        inv_mass_second_moment_of_inertia=inv_mass_second_moment_of_inertia,
        shear_matrix=shear_matrix,
        bend_matrix=bend_matrix,
        density=density_array,
        volume=volume,
        mass=mass,
        internal_forces=np.zeros((3, n + 1)),
        internal_torques=np.zeros((3, n)),
        external_forces=np.zeros((3, n + 1)),
        external_torques=np.zeros((3, n)),
        lengths=np.zeros(n),
        rest_lengths=rest_lengths,
        tangents=np.zeros((3, n)),
        dilatation=np.zeros(n),
        dilatation_rate=np.zeros(n),
        voronoi_dilatation=np.zeros(n - 1),
        rest_voronoi_lengths=rest_voronoi_lengths,
        sigma=np.zeros((3, n)),
        kappa=np.zeros((3, n - 1)),
        rest_sigma=np.zeros((3, n)),
        rest_kappa=np.zeros((3, n - 1)),
        internal_stress=np.zeros((3, n)),
        internal_couple=np.zeros((3, n - 1)),
    )


class CosseratRod(RodBase):
    """Rod state container + the ``compute_internal_forces_and_torques`` entry point."""

    def __init__(self, **arrays):
        self.n_elems = arrays["n_elements"]
        self.position_collection = arrays["position"]
        self.velocity_collection = arrays["velocity"]
        self.omega_collection = arrays["omega"]
        self.acceleration_collection = arrays["acceleration"]
        self.alpha_collection = arrays["angular_acceleration"]
        self.director_collection = arrays["directors"]
        self.radius = arrays["radius"]
        self.mass_second_moment_of_inertia = arrays["mass_second_moment_of_inertia"]
        self.inv_mass_second_moment_of_inertia = arrays[
            "inv_mass_second_moment_of_inertia"
        ]
        self.shear_matrix = arrays["shear_matrix"]
        self.bend_matrix = arrays["bend_matrix"]
        self.density = arrays["density"]
        self.volume = arrays["volume"]
        self.mass = arrays["mass"]
        self.internal_forces = arrays["internal_forces"]
        self.internal_torques = arrays["internal_torques"]
        self.external_forces = arrays["external_forces"]
        self.external_torques = arrays["external_torques"]
        self.lengths = arrays["lengths"]
        self.rest_lengths = arrays["rest_lengths"]
        self.tangents = arrays["tangents"]
        self.dilatation = arrays["dilatation"]
        self.dilatation_rate = arrays["dilatation_rate"]
        self.voronoi_dilatation = arrays["voronoi_dilatation"]
        self.rest_voronoi_lengths = arrays["rest_voronoi_lengths"]
        self.sigma = arrays["sigma"]
        self.kappa = arrays["kappa"]
        self.rest_sigma = arrays["rest_sigma"]
        self.rest_kappa = arrays["rest_kappa"]
        self.internal_stress = arrays["internal_stress"]
        self.internal_couple = arrays["internal_couple"]
        self.ghost_elems_idx = np.empty((0,), dtype=np.int64)
        self.ghost_voronoi_idx = np.empty((0,), dtype=np.int64)

        # Upstream populates lengths / tangents / radius / dilatation / sigma / kappa at
        # construction.  Note radius is *recomputed* from volume and the guarded length here,
        # which is what the reference later reads as rod.radius[-1]
        # (/root/reference/br2/free_simulator.py:356).
        _compute_shear_stretch_strains(
            self.position_collection,
            self.volume,
            self.lengths,
            self.tangents,
            self.radius,
            self.rest_lengths,
            self.rest_voronoi_lengths,
            self.dilatation,
            self.voronoi_dilatation,
            self.director_collection,
            self.sigma,
        )
        _compute_bending_twist_strains(
            self.director_collection, self.rest_voronoi_lengths, self.kappa
        )

    @classmethod
    def straight_rod(
        cls,
        n_elements,
        start,
        direction,
        normal,
        base_length,
        base_radius,
        density,
        nu=None,
        youngs_modulus=None,
        *args,
        **kwargs,
    ):
        # Extra keys the reference passes (outer_radius, inner_radius, damping_constant,
        # hollow, Info, is_first_segment is popped earlier) ride along in kwargs and are
        # ignored, as upstream does (/root/reference/br2/free_simulator.py:291-293).
        arrays = allocate_straight_rod(
            n_elements,
            start,
            direction,
            normal,
            base_length,
            base_radius,
            density,
            youngs_modulus,
            shear_modulus=kwargs.get("shear_modulus", None),
        )
        return cls(**arrays)

    # -- stepper-facing API ------------------------------------------------------------
    def compute_internal_forces_and_torques(self, time):
        _compute_internal_forces(
            self.position_collection,
            self.volume,
            self.lengths,
            self.tangents,
            self.radius,
            self.rest_lengths,
            self.rest_voronoi_lengths,
            self.dilatation,
            self.voronoi_dilatation,
            self.director_collection,
            self.sigma,
            self.rest_sigma,
            self.shear_matrix,
            self.internal_stress,
            self.internal_forces,
            self.ghost_elems_idx,
        )
        _compute_internal_torques(
            self.position_collection,
            self.velocity_collection,
            self.tangents,
            self.lengths,
            self.rest_lengths,
            self.director_collection,
            self.rest_voronoi_lengths,
            self.bend_matrix,
            self.rest_kappa,
            self.kappa,
            self.voronoi_dilatation,
            self.mass_second_moment_of_inertia,
            self.omega_collection,
            self.internal_stress,
            self.internal_couple,
            self.dilatation,
            self.dilatation_rate,
            self.internal_torques,
            self.ghost_voronoi_idx,
        )

    update_internal_forces_and_torques = compute_internal_forces_and_torques

    def update_accelerations(self, time):
        _update_accelerations(
            self.acceleration_collection,
            self.internal_forces,
            self.external_forces,
            self.mass,
            self.alpha_collection,
            self.inv_mass_second_moment_of_inertia,
            self.internal_torques,
            self.external_torques,
            self.dilatation,
        )

    def zeroed_out_external_forces_and_torques(self, time):
        _zeroed_out_external_forces_and_torques(
            self.external_forces, self.external_torques
        )

    def compute_position_center_of_mass(self):
        mass_times_position = np.einsum("j,ij->ij", self.mass, self.position_collection)
        return np.einsum("ij->i", mass_times_position) / self.mass.sum()

    def compute_velocity_center_of_mass(self):
        mass_times_velocity = np.einsum("j,ij->ij", self.mass, self.velocity_collection)
        return np.einsum("ij->i", mass_times_velocity) / self.mass.sum()
