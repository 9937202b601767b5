"""``elastica.joint.FreeJoint`` (TEST INFRASTRUCTURE).  Base class of
``TipToTipStraightJoint`` (/root/reference/br2/legacy_surface_connection_parallel_rod_numba.py:346)
and of upstream's ``SurfaceJointSideBySide``."""
import numpy as np


class FreeJoint:
    def __init__(self, k, nu):
        self.k = k
        self.nu = nu

    def apply_forces(self, rod_one, index_one, rod_two, index_two):
        end_distance_vector = (
            rod_two.position_collection[..., index_two]
            - rod_one.position_collection[..., index_one]
        )
        elastic_force = self.k * end_distance_vector
        relative_velocity = (
            rod_two.velocity_collection[..., index_two]
            - rod_one.velocity_collection[..., index_one]
        )
        damping_force = self.nu * relative_velocity
        contact_force = elastic_force + damping_force
        rod_one.external_forces[..., index_one] += contact_force
        rod_two.external_forces[..., index_two] -= contact_force

    def apply_torques(self, rod_one, index_one, rod_two, index_two):
        pass
