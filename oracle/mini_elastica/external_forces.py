"""Forcing operators (oracle restatement of ``elastica.external_forces``, PyElastica 0.3.1).

TEST INFRASTRUCTURE.  ``NoForces`` is the base of the reference's fibre actuation classes
(/root/reference/br2/free_custom_systems.py:20,66) and ``PointForces``
(/root/reference/br2/custom_activation.py:9); ``GravityForces`` is registered at
/root/reference/br2/free_simulator.py:334-338; ``UniformTorques`` is imported (:10-13) but unused.
"""
import numpy as np
from numba import njit

from ._linalg import _batch_matvec, _batch_product_i_k_to_ik


class NoForces:
    def __init__(self):
        pass

    def apply_forces(self, system, time=0.0):
        pass

    def apply_torques(self, system, time=0.0):
        pass


class GravityForces(NoForces):
    def __init__(self, acc_gravity=np.array([0.0, -9.80665, 0.0])):
        super().__init__()
        self.acc_gravity = acc_gravity

    def apply_forces(self, system, time=0.0):
        self.compute_gravity_forces(self.acc_gravity, system.mass, system.external_forces)

    @staticmethod
    @njit(cache=True)
    def compute_gravity_forces(acc_gravity, mass, external_forces):
        # f_ext[i, k] += g[i] * m[k]
        contribution = _batch_product_i_k_to_ik(acc_gravity, mass)
        for i in range(3):
            for k in range(mass.shape[0]):
                external_forces[i, k] += contribution[i, k]


class UniformTorques(NoForces):
    def __init__(self, torque, direction=np.array([0.0, 0.0, 0.0])):
        super().__init__()
        self.torque = torque * direction

    def apply_torques(self, system, time=0.0):
        n_elems = system.n_elems
        torque_on_one_element = (
            _batch_product_i_k_to_ik(self.torque, np.ones((n_elems))) / n_elems
        )
        system.external_torques += _batch_matvec(
            system.director_collection, torque_on_one_element
        )
