"""Serial (tip-to-base) joint descriptor and host versions of the rotation helpers.

Module name kept for import compatibility with the reference
(``from br2.legacy_surface_connection_parallel_rod_numba import TipToTipStraightJoint``,
/root/reference/br2/free_simulator.py:26).  Of the reference module only two things are live:

* ``TipToTipStraightJoint`` (reference :346-478): spring between the surface points of rod one's last
  and rod two's first element, un-projected damping, halves to the end nodes; its ``apply_torques``
  is a DEBUG branch that copies rod one's tip director / omega / alpha onto rod two's base element
  and returns (:474-478).  Both behaviours are reproduced by the kernel (csrc/br2_cuda.cu, phase P3).
* the single-matrix exp / log maps (:27-88), provided here as numpy helpers for host-side set-up.

The module's own ``SurfaceJointSideBySide`` (:149-343) is shadowed in the reference by PyElastica's
class of the same name and never runs; it is intentionally absent (use ``br2.hooks``'s).
"""
import numpy as np

from .hooks.ops import JOINT_TIP_TO_BASE, FreeJoint


def _single_get_rotation_matrix(theta, unit_axis):
    """Row-director Rodrigues matrix: the transpose of the usual R(axis, theta)."""
    a = np.asarray(unit_axis, dtype=np.float64)
    s, q = np.sin(theta), 1.0 - np.cos(theta)
    skew = np.array([[0.0, a[2], -a[1]], [-a[2], 0.0, a[0]], [a[1], -a[0], 0.0]])
    return np.eye(3) + s * skew + q * (np.outer(a, a) - (a @ a) * np.eye(3))


def _single_inv_rotate(director):
    """Axis-angle vector of a row-director rotation matrix (general branch plus the reference's
    near-identity / near-pi special cases)."""
    R = np.asarray(director, dtype=np.float64)
    w = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    tr = np.trace(R)
    if abs(tr - 3.0) <= 1e-8 + 3e-5:
        return w * (0.5 - (tr - 3.0) / 12.0)
    if abs(tr + 1.0) <= 1e-8 + 1e-5:
        a = int(np.argmax(np.diag(R)))
        b, c = (a + 1) % 3, (a + 2) % 3
        s = np.sqrt(R[a, a] - R[b, b] - R[c, c] + 1.0)
        v = np.array([s / 2.0, (R[b, a] + R[a, b]) / (2.0 * s), (R[c, a] + R[a, c]) / (2.0 * s)])
        return np.pi * v / np.linalg.norm(v)
    theta = np.arccos(0.5 * tr - 0.5)
    return w * (-0.5 * theta / np.sin(theta + 1e-14))


def _inv_rotate(director_collection):
    return np.stack([_single_inv_rotate(director_collection[:, :, k])
                     for k in range(director_collection.shape[2])], axis=1)


class TipToTipStraightJoint(FreeJoint):
    def __init__(self, k, nu, kt, rod1_rd2_local, rod2_rd2_local, stability_check=False):
        super().__init__(k, nu)
        self.kt = kt
        self.rod1_rd2_local = np.asarray(rod1_rd2_local, dtype=np.float64)
        self.rod2_rd2_local = np.asarray(rod2_rd2_local, dtype=np.float64)
        self.stability_check = stability_check

    def lower_joint(self, ctx, rod_one, index_one, rod_two, index_two):
        # the reference class ignores the connect indices: always rod one's LAST and rod two's FIRST element
        d = [float(self.k), float(self.nu), 0.0, *self.rod1_rd2_local, *self.rod2_rd2_local, 0.0]
        ctx.add_joint(JOINT_TIP_TO_BASE, rod_one, -1, rod_two, 0, d, overwrite=True, torque=False)
