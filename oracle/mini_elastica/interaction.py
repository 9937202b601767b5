"""Node<->element helpers (``elastica.interaction``; TEST INFRASTRUCTURE).  Imported at
/root/reference/br2/legacy_surface_connection_parallel_rod_numba.py:18-22 and used only by the
dead legacy parallel joint (:212-262)."""
import numpy as np
from numba import njit


@njit(cache=True)
def node_to_element_position(node_position_collection):
    n_elem = node_position_collection.shape[1] - 1
    out = np.empty((3, n_elem))
    for k in range(n_elem):
        out[0, k] = 0.5 * (node_position_collection[0, k + 1] + node_position_collection[0, k])
        out[1, k] = 0.5 * (node_position_collection[1, k + 1] + node_position_collection[1, k])
        out[2, k] = 0.5 * (node_position_collection[2, k + 1] + node_position_collection[2, k])
    return out


@njit(cache=True)
def node_to_element_velocity(mass, node_velocity_collection):
    n_elem = node_velocity_collection.shape[1] - 1
    out = np.empty((3, n_elem))
    for k in range(n_elem):
        for i in range(3):
            out[i, k] = (
                mass[k + 1] * node_velocity_collection[i, k + 1]
                + mass[k] * node_velocity_collection[i, k]
            )
            out[i, k] /= mass[k + 1] + mass[k]
    return out


@njit(cache=True)
def elements_to_nodes_inplace(vector_in_element_frame, vector_in_node_frame):
    for i in range(3):
        for k in range(vector_in_element_frame.shape[1]):
            vector_in_node_frame[i, k] += 0.5 * vector_in_element_frame[i, k]
            vector_in_node_frame[i, k + 1] += 0.5 * vector_in_element_frame[i, k]
