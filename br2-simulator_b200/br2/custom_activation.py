"""``PointForces`` (reference: /root/reference/br2/custom_activation.py:9-81) as a device-lowerable
forcing: ``f_ext[:, int(n_nodes * location)] += force * min(1, t / ramp_up_time)``."""
import numpy as np

from .hooks.ops import FORCE_POINT, NoForces


class PointForces(NoForces):
    def __init__(self, force, location, ramp_up_time=1e-3):
        super().__init__()
        self.force = np.asarray(force, dtype=np.float64)
        if ramp_up_time < 0.0:
            raise AssertionError("ramp_up_time must be non-negative")
        if not 0.0 < location < 1.0:
            raise AssertionError("location must lie strictly between 0 and 1")
        self.ramp_up_time = ramp_up_time
        self.location = location

    def lower_forcing(self, ctx, rod_index):
        n_nodes = ctx.rods[rod_index].n_elems + 1
        ctx.add_forcing(rod_index, FORCE_POINT, node=int(n_nodes * self.location),
                        d=[*self.force, self.ramp_up_time])
