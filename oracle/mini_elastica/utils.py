"""``elastica.utils`` names the reference imports (TEST INFRASTRUCTURE).

``Tolerance`` is imported at /root/reference/br2/free_custom_systems.py:3 and
/root/reference/br2/legacy_surface_connection_parallel_rod_numba.py:7 (unused there)."""
import numpy as np


class MaxDimension:
    @staticmethod
    def value():
        return 3


class Tolerance:
    @staticmethod
    def atol():
        return float(np.finfo(np.float64).eps * 1e4)

    @staticmethod
    def rtol():
        return float(np.finfo(np.float64).eps * 1e11)
