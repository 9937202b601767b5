"""Dampers (oracle restatement of ``elastica.dissipation``, PyElastica 0.3.1;
SURVEY.md Appendix A.5).

TEST INFRASTRUCTURE.  Registered by the reference at /root/reference/br2/free_simulator.py:314-320
with ``damping_constant`` (0.5 in the sample library) and ``time_step``."""
import numpy as np


class DamperBase:
    def __init__(self, *args, **kwargs):
        self._system = kwargs["_system"]

    @property
    def system(self):
        return self._system

    def dampen_rates(self, rod, time):
        pass


class AnalyticalLinearDamper(DamperBase):
    def __init__(self, damping_constant, time_step, **kwargs):
        super().__init__(**kwargs)
        nodal_mass = self._system.mass
        self.translational_damping_coefficient = np.exp(-damping_constant * time_step)
        element_mass = 0.5 * (nodal_mass[1:] + nodal_mass[:-1])
        element_mass[0] += 0.5 * nodal_mass[0]
        element_mass[-1] += 0.5 * nodal_mass[-1]
        self.rotational_damping_coefficient = np.exp(
            -damping_constant
            * time_step
            * element_mass
            * np.diagonal(self._system.inv_mass_second_moment_of_inertia).T
        )

    def dampen_rates(self, rod, time):
        rod.velocity_collection[:] = (
            rod.velocity_collection * self.translational_damping_coefficient
        )
        rod.omega_collection[:] = rod.omega_collection * np.power(
            self.rotational_damping_coefficient, rod.dilatation
        )
