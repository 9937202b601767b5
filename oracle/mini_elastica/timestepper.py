"""PositionVerlet over a system collection (oracle restatement of
``elastica.timestepper``, PyElastica 0.3.1).

TEST INFRASTRUCTURE.  Reference call sites: ``PositionVerlet()`` at
/root/reference/br2/environment.py:155, ``extend_stepper_interface`` at :221-223 and the
``do_step(stepper, stages, collection, time, dt) -> time`` call at :279-285.
"""
from numba import njit

from ._linalg import _batch_matmul
from ._rotations import _get_rotation_matrix


@njit(cache=True)
def overload_operator_kinematic_numba(
    n_nodes, prefac, position_collection, director_collection, velocity_collection, omega_collection
):
    # x += prefac * v
    for i in range(3):
        for k in range(n_nodes):
            position_collection[i, k] += prefac * velocity_collection[i, k]
    # Q <- R(prefac * w) Q
    rotation_matrix = _get_rotation_matrix(1.0, prefac * omega_collection)
    director_collection[:] = _batch_matmul(rotation_matrix, director_collection)


@njit(cache=True)
def overload_operator_dynamic_numba(
    velocity_collection, omega_collection, scaled_acceleration, scaled_alpha
):
    # v += (dt * a), w += (dt * alpha): the product is formed first, then added
    for i in range(3):
        for k in range(velocity_collection.shape[1]):
            velocity_collection[i, k] += scaled_acceleration[i, k]
    for i in range(3):
        for k in range(omega_collection.shape[1]):
            omega_collection[i, k] += scaled_alpha[i, k]


class PositionVerlet:
    """Second-order symplectic stepper: half kinematic, full dynamic, half kinematic."""

    def _first_prefactor(self, dt):
        return 0.5 * dt

    def _first_kinematic_step(self, system, time, dt):
        prefac = self._first_prefactor(dt)
        overload_operator_kinematic_numba(
            system.position_collection.shape[1],
            prefac,
            system.position_collection,
            system.director_collection,
            system.velocity_collection,
            system.omega_collection,
        )

    def _first_dynamic_step(self, system, time, dt):
        system.update_accelerations(time)
        overload_operator_dynamic_numba(
            system.velocity_collection,
            system.omega_collection,
            dt * system.acceleration_collection,
            dt * system.alpha_collection,
        )


class _SystemCollectionStepper:
    """The ``stages_and_updates`` object handed back by ``extend_stepper_interface``."""

    def __init__(self, stepper):
        self.kin_prefactor = stepper._first_prefactor
        self.kin_step = stepper._first_kinematic_step
        self.dyn_step = stepper._first_dynamic_step

    def do_step(self, collection, time, dt):
        systems = collection._systems
        # -- stage 1
        for system in systems:
            self.kin_step(system, time, dt)
        time += self.kin_prefactor(dt)
        collection.constrain_values(time)
        for system in systems:
            system.update_internal_forces_and_torques(time)
        collection.synchronize(time)
        for system in systems:
            self.dyn_step(system, time, dt)
        collection.constrain_rates(time)
        # -- trailing kinematic half step
        for system in systems:
            self.kin_step(system, time, dt)
        time += self.kin_prefactor(dt)
        collection.constrain_values(time)
        collection.apply_callbacks(time, round(time / dt))
        for system in systems:
            system.zeroed_out_external_forces_and_torques(time)
        return time


def _do_step(stepper, stages_and_updates, collection, time, dt):
    return stages_and_updates.do_step(collection, time, dt)


def extend_stepper_interface(stepper, system_collection):
    """Returns ``(do_step, stages_and_updates)`` with
    ``do_step(stepper, stages_and_updates, system_collection, time, dt) -> time``."""
    return _do_step, _SystemCollectionStepper(stepper)


def integrate(stepper, system_collection, final_time, n_steps, progress_bar=False):
    do_step, stages = extend_stepper_interface(stepper, system_collection)
    dt = float(final_time) / n_steps
    time = 0.0
    for _ in range(n_steps):
        time = do_step(stepper, stages, system_collection, time, dt)
    return time
