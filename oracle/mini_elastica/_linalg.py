"""Batched 3-vector / 3x3 helpers (oracle restatement of ``elastica._linalg``, PyElastica 0.3.1).

TEST INFRASTRUCTURE.  Reference call sites that import these names:
/root/reference/br2/free_simulator.py:18-19, /root/reference/br2/free_custom_systems.py:6-7,
/root/reference/br2/legacy_surface_connection_parallel_rod_numba.py:10-17.

All loops accumulate in index order (no FMA contraction: numba does not enable fast-math),
which the plain-C fused oracle reproduces term for term.
"""
import numpy as np
from numba import njit


@njit(cache=True)
def _batch_matvec(matrix_collection, vector_collection):
    blocksize = vector_collection.shape[1]
    out = np.zeros((3, blocksize))
    for i in range(3):
        for j in range(3):
            for k in range(blocksize):
                out[i, k] += matrix_collection[i, j, k] * vector_collection[j, k]
    return out


@njit(cache=True)
def _batch_matmul(first_matrix_collection, second_matrix_collection):
    blocksize = first_matrix_collection.shape[2]
    out = np.zeros((3, 3, blocksize))
    for i in range(3):
        for j in range(3):
            for m in range(3):
                for k in range(blocksize):
                    out[i, m, k] += (
                        first_matrix_collection[i, j, k]
                        * second_matrix_collection[j, m, k]
                    )
    return out


@njit(cache=True)
def _batch_cross(first_vector_collection, second_vector_collection):
    blocksize = first_vector_collection.shape[1]
    out = np.empty((3, blocksize))
    for k in range(blocksize):
        out[0, k] = (
            first_vector_collection[1, k] * second_vector_collection[2, k]
            - first_vector_collection[2, k] * second_vector_collection[1, k]
        )
        out[1, k] = (
            first_vector_collection[2, k] * second_vector_collection[0, k]
            - first_vector_collection[0, k] * second_vector_collection[2, k]
        )
        out[2, k] = (
            first_vector_collection[0, k] * second_vector_collection[1, k]
            - first_vector_collection[1, k] * second_vector_collection[0, k]
        )
    return out


@njit(cache=True)
def _batch_dot(first_vector, second_vector):
    blocksize = first_vector.shape[1]
    out = np.zeros((blocksize,))
    for i in range(3):
        for k in range(blocksize):
            out[k] += first_vector[i, k] * second_vector[i, k]
    return out


@njit(cache=True)
def _batch_norm(vector):
    blocksize = vector.shape[1]
    out = np.empty((blocksize,))
    for k in range(blocksize):
        out[k] = np.sqrt(
            vector[0, k] * vector[0, k]
            + vector[1, k] * vector[1, k]
            + vector[2, k] * vector[2, k]
        )
    return out


@njit(cache=True)
def _batch_product_i_k_to_ik(vector1, vector2):
    blocksize = vector2.shape[0]
    out = np.empty((3, blocksize))
    for i in range(3):
        for k in range(blocksize):
            out[i, k] = vector1[i] * vector2[k]
    return out


@njit(cache=True)
def _batch_product_i_ik_to_k(vector1, vector2):
    blocksize = vector2.shape[1]
    out = np.zeros((blocksize,))
    for i in range(3):
        for k in range(blocksize):
            out[k] += vector1[i] * vector2[i, k]
    return out


@njit(cache=True)
def _batch_product_k_ik_to_ik(vector1, vector2):
    blocksize = vector2.shape[1]
    out = np.empty((3, blocksize))
    for i in range(3):
        for k in range(blocksize):
            out[i, k] = vector1[k] * vector2[i, k]
    return out


@njit(cache=True)
def _batch_vector_sum(vector1, vector2):
    blocksize = vector1.shape[1]
    out = np.empty((3, blocksize))
    for i in range(3):
        for k in range(blocksize):
            out[i, k] = vector1[i, k] + vector2[i, k]
    return out


@njit(cache=True)
def _batch_matrix_transpose(input_matrix):
    blocksize = input_matrix.shape[2]
    out = np.empty((3, 3, blocksize))
    for i in range(3):
        for j in range(3):
            for k in range(blocksize):
                out[j, i, k] = input_matrix[i, j, k]
    return out
