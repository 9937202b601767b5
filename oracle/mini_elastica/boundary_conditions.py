"""``elastica.boundary_conditions.ConstraintBase`` (TEST INFRASTRUCTURE).

Base of ``FreeBaseEndSoftFixed`` (/root/reference/br2/free_custom_systems.py:104-149) and
``LastEndFixedRod`` (/root/reference/br2/custom_constraint.py:9).  The wrapper passes copies of
``x[:, idx]`` / ``Q[:, :, idx]`` as leading positional args and the rod as ``_system``."""
import numpy as np


class ConstraintBase:
    def __init__(self, *args, **kwargs):
        self._system = kwargs.get("_system", None)
        self._constrained_position_idx = np.array(
            kwargs.get("constrained_position_idx", []), dtype=int
        )
        self._constrained_director_idx = np.array(
            kwargs.get("constrained_director_idx", []), dtype=int
        )

    @property
    def system(self):
        return self._system

    @property
    def constrained_position_idx(self):
        return self._constrained_position_idx

    @property
    def constrained_director_idx(self):
        return self._constrained_director_idx

    def constrain_values(self, rod, time):
        pass

    def constrain_rates(self, rod, time):
        pass


FreeBC = ConstraintBase
