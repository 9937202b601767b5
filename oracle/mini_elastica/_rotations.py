"""Exponential / logarithm maps on row-director frames (oracle restatement of
``elastica._rotations``, PyElastica 0.3.1).

TEST INFRASTRUCTURE.  Convention check against the in-repo single-matrix versions:
/root/reference/br2/legacy_surface_connection_parallel_rod_numba.py:27-49 (same matrix layout)
and :52-88 (same log-map general branch; upstream's rod kernel has no small-angle branches
and uses the -1e-10 / +1e-14 guards instead).
"""
import numpy as np
from numba import njit

# Named guards: flip here if a check against real PyElastica disagrees.
ROTATION_NORM_GUARD = 1e-14
ARCCOS_GUARD = 1e-10
SIN_GUARD = 1e-14


@njit(cache=True)
def _get_rotation_matrix(scale, axis_collection):
    blocksize = axis_collection.shape[1]
    rot_mat = np.empty((3, 3, blocksize))
    for k in range(blocksize):
        v0 = axis_collection[0, k]
        v1 = axis_collection[1, k]
        v2 = axis_collection[2, k]
        theta = np.sqrt(v0 * v0 + v1 * v1 + v2 * v2)
        v0 /= theta + ROTATION_NORM_GUARD
        v1 /= theta + ROTATION_NORM_GUARD
        v2 /= theta + ROTATION_NORM_GUARD
        theta *= scale
        u_prefix = np.sin(theta)
        u_sq_prefix = 1.0 - np.cos(theta)

        rot_mat[0, 0, k] = 1.0 - u_sq_prefix * (v1 * v1 + v2 * v2)
        rot_mat[1, 1, k] = 1.0 - u_sq_prefix * (v0 * v0 + v2 * v2)
        rot_mat[2, 2, k] = 1.0 - u_sq_prefix * (v0 * v0 + v1 * v1)

        rot_mat[0, 1, k] = u_prefix * v2 + u_sq_prefix * v0 * v1
        rot_mat[1, 0, k] = -u_prefix * v2 + u_sq_prefix * v0 * v1
        rot_mat[0, 2, k] = -u_prefix * v1 + u_sq_prefix * v0 * v2
        rot_mat[2, 0, k] = u_prefix * v1 + u_sq_prefix * v0 * v2
        rot_mat[1, 2, k] = u_prefix * v0 + u_sq_prefix * v1 * v2
        rot_mat[2, 1, k] = -u_prefix * v0 + u_sq_prefix * v1 * v2
    return rot_mat


@njit(cache=True)
def _inv_rotate(director_collection):
    """Log map of Q_{k+1} Q_k^T for consecutive elements: (3,3,n) -> (3,n-1)."""
    blocksize = director_collection.shape[2] - 1
    vector_collection = np.empty((3, blocksize))
    for k in range(blocksize):
        # antisymmetric part of R = Q_{k+1} Q_k^T  (R_ab = row a of Q_{k+1} . row b of Q_k)
        vector_collection[0, k] = (
            director_collection[2, 0, k + 1] * director_collection[1, 0, k]
            + director_collection[2, 1, k + 1] * director_collection[1, 1, k]
            + director_collection[2, 2, k + 1] * director_collection[1, 2, k]
        ) - (
            director_collection[1, 0, k + 1] * director_collection[2, 0, k]
            + director_collection[1, 1, k + 1] * director_collection[2, 1, k]
            + director_collection[1, 2, k + 1] * director_collection[2, 2, k]
        )
        vector_collection[1, k] = (
            director_collection[0, 0, k + 1] * director_collection[2, 0, k]
            + director_collection[0, 1, k + 1] * director_collection[2, 1, k]
            + director_collection[0, 2, k + 1] * director_collection[2, 2, k]
        ) - (
            director_collection[2, 0, k + 1] * director_collection[0, 0, k]
            + director_collection[2, 1, k + 1] * director_collection[0, 1, k]
            + director_collection[2, 2, k + 1] * director_collection[0, 2, k]
        )
        vector_collection[2, k] = (
            director_collection[1, 0, k + 1] * director_collection[0, 0, k]
            + director_collection[1, 1, k + 1] * director_collection[0, 1, k]
            + director_collection[1, 2, k + 1] * director_collection[0, 2, k]
        ) - (
            director_collection[0, 0, k + 1] * director_collection[1, 0, k]
            + director_collection[0, 1, k + 1] * director_collection[1, 1, k]
            + director_collection[0, 2, k + 1] * director_collection[1, 2, k]
        )
        trace = (
            (
                director_collection[0, 0, k + 1] * director_collection[0, 0, k]
                + director_collection[0, 1, k + 1] * director_collection[0, 1, k]
                + director_collection[0, 2, k + 1] * director_collection[0, 2, k]
            )
            + (
                director_collection[1, 0, k + 1] * director_collection[1, 0, k]
                + director_collection[1, 1, k + 1] * director_collection[1, 1, k]
                + director_collection[1, 2, k + 1] * director_collection[1, 2, k]
            )
            + (
                director_collection[2, 0, k + 1] * director_collection[2, 0, k]
                + director_collection[2, 1, k + 1] * director_collection[2, 1, k]
                + director_collection[2, 2, k + 1] * director_collection[2, 2, k]
            )
        )
        theta = np.arccos(0.5 * trace - 0.5 - ARCCOS_GUARD)
        scale = -0.5 * theta / np.sin(theta + SIN_GUARD)
        vector_collection[0, k] *= scale
        vector_collection[1, k] *= scale
        vector_collection[2, k] *= scale
    return vector_collection


@njit(cache=True)
def _inv_skew_symmetrize(matrix_collection):
    blocksize = matrix_collection.shape[2]
    out = np.empty((3, blocksize))
    for k in range(blocksize):
        out[0, k] = matrix_collection[2, 1, k]
        out[1, k] = matrix_collection[0, 2, k]
        out[2, k] = matrix_collection[1, 0, k]
    return out
