"""BR2 fibre actuation and the soft-fixed base, as device-lowerable descriptors.

Same class names and constructor signatures as /root/reference/br2/free_custom_systems.py; the
arithmetic itself runs inside the CUDA step kernel (csrc/br2_cuda.cu, phase P4):

* ``FreeBendActuation``   (reference :20-57)  -- material-frame couple ``P * scale * ramp * [cos z, sin z, 0]``
  on the LAST element only (the uniformly distributed variant after the ``return`` is dead code);
* ``FreeTwistActuation``  (reference :66-101) -- ``P * scale * ramp`` about d3, shared evenly by all elements;
* ``FreeBaseEndSoftFixed``(reference :104-185)-- spring ``k (x_fix - x_0) - nu (...)`` on node 0, director 0
  reset to its initial value, all rates of node/element 0 zeroed; the orientation spring ``kt`` is
  unreachable in the reference (:135, :196) and is therefore not implemented.

``ramp = min(1, t / ramp_up_time)`` with ``t`` the mid-step time.
"""
import numpy as np

from .hooks.ops import CONS_SOFT_FIXED_BASE, FORCE_BEND, FORCE_TWIST, ConstraintBase, NoForces


class FreeBendActuation(NoForces):
    def __init__(self, actuation_ref, z_angle, scale, ramp_up_time=0.2):
        super().__init__()
        self.actuation_ref = actuation_ref
        self.z_angle = z_angle
        self.magnitude_scale = scale
        self.ramp_up_time = ramp_up_time

    def lower_forcing(self, ctx, rod_index):
        ctx.add_forcing(rod_index, FORCE_BEND, action=ctx.action_index(self.actuation_ref),
                        d=[self.magnitude_scale, self.ramp_up_time, np.cos(self.z_angle), np.sin(self.z_angle)])


class FreeTwistActuation(NoForces):
    def __init__(self, actuation_ref, scale, ramp_up_time=0.2):
        super().__init__()
        self.actuation_ref = actuation_ref
        self.direction = np.array([0.0, 0.0, 1.0])
        self.scale = scale
        self.ramp_up_time = ramp_up_time

    def lower_forcing(self, ctx, rod_index):
        ctx.add_forcing(rod_index, FORCE_TWIST, action=ctx.action_index(self.actuation_ref),
                        d=[self.scale, self.ramp_up_time])


class FreeBaseEndSoftFixed(ConstraintBase):
    def __init__(self, fixed_position, fixed_directors, k, nu, kt, **kwargs):
        super().__init__(**kwargs)
        self.fixed_position = fixed_position
        self.fixed_directors = fixed_directors
        self.k = k
        self.nu = nu
        self.kt = kt
        self.rev = 0

    def lower_constraint(self, ctx, rod_index):
        ctx.add_constraint(rod_index, CONS_SOFT_FIXED_BASE, self.fixed_position, self.fixed_directors,
                           k=float(self.k), nu=float(self.nu))
