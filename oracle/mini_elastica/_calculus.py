"""Difference / averaging stencils (oracle restatement of ``elastica._calculus``, PyElastica 0.3.1).

TEST INFRASTRUCTURE.  Names imported by the reference: ``_isnan_check``
(/root/reference/br2/environment.py:15), ``_clip_array`` (/root/reference/br2/free_simulator.py:17).

For a stand-alone rod (no ghost nodes) the block-structure kernels of upstream reduce to the
zero-padded two-point stencils below:
    difference:  out_0 = a_0,  out_i = a_i - a_{i-1},  out_n = -a_{n-1}
    trapezoid :  out_0 = a_0/2, out_i = (a_i + a_{i-1})/2, out_n = a_{n-1}/2
"""
import numpy as np
from numba import njit


@njit(cache=True)
def _isnan_check(array):
    return bool(np.isnan(array).any())


@njit(cache=True)
def _clip_array(input_array, vmin, vmax):
    flat = input_array.reshape(-1)
    for i in range(flat.shape[0]):
        if flat[i] < vmin:
            flat[i] = vmin
        elif flat[i] > vmax:
            flat[i] = vmax
    return input_array


@njit(cache=True)
def position_difference_kernel(array_collection):
    """(3, n+1) nodes -> (3, n) element differences."""
    blocksize = array_collection.shape[1] - 1
    out = np.empty((3, blocksize))
    for i in range(3):
        for k in range(blocksize):
            out[i, k] = array_collection[i, k + 1] - array_collection[i, k]
    return out


@njit(cache=True)
def position_average(array_collection):
    """(m+1,) -> (m,) two-point mean, written as 0.5*(a[k+1] + a[k]) like upstream."""
    blocksize = array_collection.shape[0] - 1
    out = np.empty((blocksize,))
    for k in range(blocksize):
        out[k] = 0.5 * (array_collection[k + 1] + array_collection[k])
    return out


@njit(cache=True)
def _difference(vector):
    """(3, m) -> (3, m-1): out_k = a_{k+1} - a_k."""
    blocksize = vector.shape[1] - 1
    out = np.empty((3, blocksize))
    for i in range(3):
        for k in range(blocksize):
            out[i, k] = vector[i, k + 1] - vector[i, k]
    return out


@njit(cache=True)
def _average(vector):
    blocksize = vector.shape[0] - 1
    out = np.empty((blocksize,))
    for k in range(blocksize):
        out[k] = 0.5 * (vector[k + 1] + vector[k])
    return out


@njit(cache=True)
def difference_kernel(array_collection):
    """Zero-padded two-point difference, (3, m) -> (3, m+1)."""
    blocksize = array_collection.shape[1]
    out = np.empty((3, blocksize + 1))
    for i in range(3):
        out[i, 0] = array_collection[i, 0]
        out[i, blocksize] = -array_collection[i, blocksize - 1]
        for k in range(1, blocksize):
            out[i, k] = array_collection[i, k] - array_collection[i, k - 1]
    return out


@njit(cache=True)
def quadrature_kernel(array_collection):
    """Zero-padded trapezoid, (3, m) -> (3, m+1)."""
    blocksize = array_collection.shape[1]
    out = np.empty((3, blocksize + 1))
    for i in range(3):
        out[i, 0] = 0.5 * array_collection[i, 0]
        out[i, blocksize] = 0.5 * array_collection[i, blocksize - 1]
        for k in range(1, blocksize):
            out[i, k] = 0.5 * (array_collection[i, k] + array_collection[i, k - 1])
    return out


# Upstream's block-structure variants take a ghost-index argument; a stand-alone rod has none.
@njit(cache=True)
def difference_kernel_for_block_structure(array_collection, ghost_idx):
    return difference_kernel(array_collection)


@njit(cache=True)
def trapezoidal_for_block_structure(array_collection, ghost_idx):
    return quadrature_kernel(array_collection)
