"""Surface joint between two side-by-side rods (oracle restatement of PyElastica 0.3.1's
``elastica.experimental.connection_contact_joint.parallel_connection``).

TEST INFRASTRUCTURE.  Reference call sites: ``get_connection_vector_straight_straight_rod`` at
/root/reference/br2/free_simulator.py:404-413 and one ``SurfaceJointSideBySide`` per element per
rod pair at :417-435 (scalar ``first_connect_idx == second_connect_idx == i``).

NOTE: /root/reference/br2/legacy_surface_connection_parallel_rod_numba.py:149-343 holds an older
class of the same name; the reference shadows it with this upstream one
(/root/reference/br2/free_simulator.py:21-26), so that legacy formulation is NOT followed.
"""
import numpy as np
from numba import njit

from ..._linalg import _batch_matvec, _batch_norm
from ...joint import FreeJoint

# Named constants (SURVEY.md Appendix E confidence ledger).
DISTANCE_ROUND_DECIMALS = 12
DISTANCE_ROUND_SCALE = 10.0 ** DISTANCE_ROUND_DECIMALS
ZERO_DISTANCE_GUARD = 1e-12
# upstream also rounds the penetration strain to 12 decimals (np.round_(penetration_strain, 12)) before the `< 0` test
PENETRATION_ROUND_DECIMALS = 12
PENETRATION_ROUND_SCALE = 10.0 ** PENETRATION_ROUND_DECIMALS


def get_connection_vector_straight_straight_rod(rod_one, rod_two, rod_one_idx, rod_two_idx):
    """Unit vectors rod1->rod2 / rod2->rod1 in each rod's material frame, per element, and the
    surface-to-surface gap at rest."""
    rod_one_element_position = 0.5 * (
        rod_one.position_collection[..., 1:] + rod_one.position_collection[..., :-1]
    )
    rod_one_element_position = rod_one_element_position[:, rod_one_idx[0] : rod_one_idx[1]]
    rod_two_element_position = 0.5 * (
        rod_two.position_collection[..., 1:] + rod_two.position_collection[..., :-1]
    )
    rod_two_element_position = rod_two_element_position[:, rod_two_idx[0] : rod_two_idx[1]]

    distance_vector_rod_one_to_rod_two = rod_two_element_position - rod_one_element_position
    distance_vector_rod_one_to_rod_two_norm = _batch_norm(distance_vector_rod_one_to_rod_two)
    distance_vector_rod_one_to_rod_two /= distance_vector_rod_one_to_rod_two_norm
    distance_vector_rod_two_to_rod_one = -distance_vector_rod_one_to_rod_two

    rod_one_direction_vec_in_material_frame = _batch_matvec(
        rod_one.director_collection[:, :, rod_one_idx[0] : rod_one_idx[1]],
        distance_vector_rod_one_to_rod_two,
    )
    rod_two_direction_vec_in_material_frame = _batch_matvec(
        rod_two.director_collection[:, :, rod_two_idx[0] : rod_two_idx[1]],
        distance_vector_rod_two_to_rod_one,
    )
    offset_btw_rods = distance_vector_rod_one_to_rod_two_norm - (
        rod_one.radius[rod_one_idx[0] : rod_one_idx[1]]
        + rod_two.radius[rod_two_idx[0] : rod_two_idx[1]]
    )
    return (
        rod_one_direction_vec_in_material_frame,
        rod_two_direction_vec_in_material_frame,
        offset_btw_rods,
    )


class SurfaceJointSideBySide(FreeJoint):
    def __init__(
        self,
        k,
        nu,
        k_repulsive,
        rod_one_direction_vec_in_material_frame,
        rod_two_direction_vec_in_material_frame,
        offset_btw_rods,
        **kwargs,
    ):
        super().__init__(np.array(k), np.array(nu))
        self.k_repulsive = np.array(k_repulsive)
        self.offset_btw_rods = np.array(offset_btw_rods)
        self.rod_one_direction_vec_in_material_frame = np.array(
            rod_one_direction_vec_in_material_frame
        ).T
        self.rod_two_direction_vec_in_material_frame = np.array(
            rod_two_direction_vec_in_material_frame
        ).T

    def apply_forces(self, rod_one, index_one, rod_two, index_two):
        (self.rod_one_rd2, self.rod_two_rd2, self.spring_force) = self._apply_forces(
            float(self.k),
            float(self.nu),
            float(self.k_repulsive),
            index_one,
            index_two,
            self.rod_one_direction_vec_in_material_frame,
            self.rod_two_direction_vec_in_material_frame,
            float(self.offset_btw_rods),
            rod_one.director_collection,
            rod_two.director_collection,
            rod_one.position_collection,
            rod_two.position_collection,
            rod_one.radius,
            rod_two.radius,
            rod_one.dilatation,
            rod_two.dilatation,
            rod_one.velocity_collection,
            rod_two.velocity_collection,
            rod_one.external_forces,
            rod_two.external_forces,
        )

    @staticmethod
    @njit(cache=True)
    def _apply_forces(
        k,
        nu,
        k_repulsive,
        index_one,
        index_two,
        rod_one_direction_vec_in_material_frame,
        rod_two_direction_vec_in_material_frame,
        rest_offset_btw_rods,
        rod_one_director_collection,
        rod_two_director_collection,
        rod_one_position_collection,
        rod_two_position_collection,
        rod_one_radius,
        rod_two_radius,
        rod_one_dilatation,
        rod_two_dilatation,
        rod_one_velocity_collection,
        rod_two_velocity_collection,
        rod_one_external_forces,
        rod_two_external_forces,
    ):
        # connection directions back in the lab frame: Q^T d  (accumulated j = 0,1,2)
        rod_one_to_rod_two_connection_vec = np.zeros(3)
        rod_two_to_rod_one_connection_vec = np.zeros(3)
        for i in range(3):
            for j in range(3):
                rod_one_to_rod_two_connection_vec[i] += (
                    rod_one_director_collection[j, i, index_one]
                    * rod_one_direction_vec_in_material_frame[j]
                )
                rod_two_to_rod_one_connection_vec[i] += (
                    rod_two_director_collection[j, i, index_two]
                    * rod_two_direction_vec_in_material_frame[j]
                )

        rod_one_element_position = 0.5 * (
            rod_one_position_collection[:, index_one]
            + rod_one_position_collection[:, index_one + 1]
        )
        rod_two_element_position = 0.5 * (
            rod_two_position_collection[:, index_two]
            + rod_two_position_collection[:, index_two + 1]
        )

        # the rest gap is shared by the two surfaces and thins with stretch
        offset_rod_one = 0.5 * rest_offset_btw_rods / np.sqrt(rod_one_dilatation[index_one])
        offset_rod_two = 0.5 * rest_offset_btw_rods / np.sqrt(rod_two_dilatation[index_two])

        rod_one_rd2 = rod_one_to_rod_two_connection_vec * (
            rod_one_radius[index_one] + offset_rod_one
        )
        rod_two_rd2 = rod_two_to_rod_one_connection_vec * (
            rod_two_radius[index_two] + offset_rod_two
        )

        distance_vector = (rod_two_element_position + rod_two_rd2) - (
            rod_one_element_position + rod_one_rd2
        )
        # np.round_(x, 12) == rint(x * 1e12) / 1e12
        for i in range(3):
            distance_vector[i] = np.rint(distance_vector[i] * DISTANCE_ROUND_SCALE) / DISTANCE_ROUND_SCALE
        spring_force = k * distance_vector

        rod_one_element_velocity = 0.5 * (
            rod_one_velocity_collection[:, index_one]
            + rod_one_velocity_collection[:, index_one + 1]
        )
        rod_two_element_velocity = 0.5 * (
            rod_two_velocity_collection[:, index_two]
            + rod_two_velocity_collection[:, index_two + 1]
        )
        relative_velocity = rod_two_element_velocity - rod_one_element_velocity

        distance = np.sqrt(
            distance_vector[0] * distance_vector[0]
            + distance_vector[1] * distance_vector[1]
            + distance_vector[2] * distance_vector[2]
        )
        normalized_distance_vector = np.zeros(3)
        if distance >= ZERO_DISTANCE_GUARD:
            normalized_distance_vector[:] = distance_vector / distance
        normal_relative_velocity = (
            relative_velocity[0] * normalized_distance_vector[0]
            + relative_velocity[1] * normalized_distance_vector[1]
            + relative_velocity[2] * normalized_distance_vector[2]
        )
        damping_force = -nu * (normal_relative_velocity * normalized_distance_vector)

        total_force = spring_force + damping_force

        # Hertz-like repulsion when the cylinders interpenetrate
        center_distance = rod_two_element_position - rod_one_element_position
        center_distance_norm = np.sqrt(
            center_distance[0] * center_distance[0]
            + center_distance[1] * center_distance[1]
            + center_distance[2] * center_distance[2]
        )
        center_distance_unit_vec = center_distance / center_distance_norm
        penetration_strain = center_distance_norm - (
            rod_one_radius[index_one]
            + offset_rod_one
            + rod_two_radius[index_two]
            + offset_rod_two
        )
        penetration_strain = np.rint(penetration_strain * PENETRATION_ROUND_SCALE) / PENETRATION_ROUND_SCALE
        if penetration_strain < 0:
            total_force += (
                -k_repulsive * np.abs(penetration_strain) ** 1.5 * center_distance_unit_vec
            )

        for i in range(3):
            rod_one_external_forces[i, index_one] += 0.5 * total_force[i]
            rod_one_external_forces[i, index_one + 1] += 0.5 * total_force[i]
            rod_two_external_forces[i, index_two] -= 0.5 * total_force[i]
            rod_two_external_forces[i, index_two + 1] -= 0.5 * total_force[i]

        return rod_one_rd2, rod_two_rd2, spring_force

    def apply_torques(self, rod_one, index_one, rod_two, index_two):
        self._apply_torques(
            self.spring_force,
            self.rod_one_rd2,
            self.rod_two_rd2,
            index_one,
            index_two,
            rod_one.director_collection,
            rod_two.director_collection,
            rod_one.external_torques,
            rod_two.external_torques,
        )

    @staticmethod
    @njit(cache=True)
    def _apply_torques(
        spring_force,
        rod_one_rd2,
        rod_two_rd2,
        index_one,
        index_two,
        rod_one_director_collection,
        rod_two_director_collection,
        rod_one_external_torques,
        rod_two_external_torques,
    ):
        # r x F with the spring part only, rotated into each rod's material frame
        torque_on_rod_one = np.empty(3)
        torque_on_rod_two = np.empty(3)
        torque_on_rod_one[0] = rod_one_rd2[1] * spring_force[2] - rod_one_rd2[2] * spring_force[1]
        torque_on_rod_one[1] = rod_one_rd2[2] * spring_force[0] - rod_one_rd2[0] * spring_force[2]
        torque_on_rod_one[2] = rod_one_rd2[0] * spring_force[1] - rod_one_rd2[1] * spring_force[0]
        torque_on_rod_two[0] = rod_two_rd2[1] * (-spring_force[2]) - rod_two_rd2[2] * (-spring_force[1])
        torque_on_rod_two[1] = rod_two_rd2[2] * (-spring_force[0]) - rod_two_rd2[0] * (-spring_force[2])
        torque_on_rod_two[2] = rod_two_rd2[0] * (-spring_force[1]) - rod_two_rd2[1] * (-spring_force[0])
        for i in range(3):
            t1 = 0.0
            t2 = 0.0
            for j in range(3):
                t1 += rod_one_director_collection[i, j, index_one] * torque_on_rod_one[j]
                t2 += rod_two_director_collection[i, j, index_two] * torque_on_rod_two[j]
            rod_one_external_torques[i, index_one] += t1
            rod_two_external_torques[i, index_two] += t2
