"""System collection and its feature mixins (oracle restatement of ``elastica.modules`` /
``elastica.wrappers``, PyElastica 0.3.1).

TEST INFRASTRUCTURE.  The reference composes them as
``class BR2Simulator(BaseSystemCollection, Constraints, Connections, Forcing, Damping, CallBacks)``
(/root/reference/br2/free_simulator.py:36-38) and uses ``append`` (:299), ``dampen().using`` (:316),
``constrain().using`` (:324), ``add_forcing_to().using`` (:335,351,358), ``connect().using``
(:418-435, :451-461), ``collect_diagnostics().using`` (:397) and ``finalize`` (environment.py:218).

Order semantics that matter numerically: every mixin calls
``super().__init__()`` *before* registering its feature, so with the MRO above
synchronize = [forcing, connections] and constrain_rates = [dampers, constraints].
Forcing / constraints / dampers / callbacks are stably sorted by system index at finalize;
connections keep registration order.
"""
from collections.abc import MutableSequence

import numpy as np

from .rod import RodBase


class BaseSystemCollection(MutableSequence):
    def __init__(self):
        # feature groups must exist before the mixins register into them
        self._feature_group_synchronize = []
        self._feature_group_constrain_values = []
        self._feature_group_constrain_rates = []
        self._feature_group_callback = []
        self._feature_group_finalize = []
        super(BaseSystemCollection, self).__init__()
        self.allowed_sys_types = (RodBase,)
        self._systems = []
        self._finalize_flag = False

    # -- MutableSequence ---------------------------------------------------------------
    def _check_type(self, sys_to_be_added):
        if not issubclass(sys_to_be_added.__class__, self.allowed_sys_types):
            raise TypeError(
                "{0} is not a system this collection can integrate".format(
                    sys_to_be_added.__class__
                )
            )
        return True

    def __len__(self):
        return len(self._systems)

    def __getitem__(self, idx):
        return self._systems[idx]

    def __delitem__(self, idx):
        del self._systems[idx]

    def __setitem__(self, idx, system):
        self._check_type(system)
        self._systems[idx] = system

    def insert(self, idx, system):
        self._check_type(system)
        self._systems.insert(idx, system)

    def __str__(self):
        return str(self._systems)

    def extend_allowed_types(self, additional_types):
        self.allowed_sys_types += additional_types

    def override_allowed_types(self, allowed_types):
        self.allowed_sys_types = allowed_types

    def _get_sys_idx_if_valid(self, sys_to_be_tested):
        n_systems = len(self._systems)
        if isinstance(sys_to_be_tested, (int, np.integer)):
            assert -n_systems <= sys_to_be_tested < n_systems, "system index out of range"
            return int(sys_to_be_tested)
        self._check_type(sys_to_be_tested)
        for idx, system in enumerate(self._systems):
            if system is sys_to_be_tested:
                return idx
        raise ValueError("system was not appended to this collection")

    # -- life cycle --------------------------------------------------------------------
    def finalize(self):
        assert not self._finalize_flag, "The finalize cannot be called twice."
        for finalize in self._feature_group_finalize:
            finalize()
        self._feature_group_finalize = None
        self._finalize_flag = True

    def synchronize(self, time):
        for feature in self._feature_group_synchronize:
            feature(time)

    def constrain_values(self, time):
        for feature in self._feature_group_constrain_values:
            feature(time)

    def constrain_rates(self, time):
        for feature in self._feature_group_constrain_rates:
            feature(time)

    def apply_callbacks(self, time, current_step):
        for feature in self._feature_group_callback:
            feature(time, current_step)


class _Deferred:
    """``xxx(system).using(cls, *args, **kwargs)`` builder shared by all features."""

    def __init__(self, sys_idx):
        self._sys_idx = sys_idx
        self._cls = None
        self._args = ()
        self._kwargs = {}

    def using(self, cls, *args, **kwargs):
        self._cls = cls
        self._args = args
        self._kwargs = kwargs
        return self

    def id(self):
        return self._sys_idx


# ------------------------------------------------------------------------------ constraints
class _Constraint(_Deferred):
    def using(self, cls, *args, **kwargs):
        from .boundary_conditions import ConstraintBase

        assert issubclass(cls, ConstraintBase), "not a constraint class"
        return super().using(cls, *args, **kwargs)

    def __call__(self, rod):
        pos_idx = self._kwargs.get("constrained_position_idx", ())
        dir_idx = self._kwargs.get("constrained_director_idx", ())
        positions = [rod.position_collection[..., idx].copy() for idx in pos_idx]
        directors = [rod.director_collection[..., idx].copy() for idx in dir_idx]
        return self._cls(*positions, *directors, *self._args, _system=rod, **self._kwargs)


class Constraints:
    def __init__(self):
        self._constraints = []
        super(Constraints, self).__init__()
        self._feature_group_constrain_values.append(self._constrain_values)
        self._feature_group_constrain_rates.append(self._constrain_rates)
        self._feature_group_finalize.append(self._finalize_constraints)

    def constrain(self, system):
        sys_idx = self._get_sys_idx_if_valid(system)
        _constraint = _Constraint(sys_idx)
        self._constraints.append(_constraint)
        return _constraint

    def _finalize_constraints(self):
        self._constraints[:] = [
            (c.id(), c(self._systems[c.id()])) for c in self._constraints
        ]
        self._constraints.sort(key=lambda x: x[0])
        # initial conditions must satisfy the constraints
        self._constrain_values(0.0)
        self._constrain_rates(0.0)

    def _constrain_values(self, time):
        for sys_id, constraint in self._constraints:
            constraint.constrain_values(self._systems[sys_id], time)

    def _constrain_rates(self, time):
        for sys_id, constraint in self._constraints:
            constraint.constrain_rates(self._systems[sys_id], time)


# ------------------------------------------------------------------------------ connections
class _Connect:
    def __init__(self, first_sys_idx, second_sys_idx):
        self._first_sys_idx = first_sys_idx
        self._second_sys_idx = second_sys_idx
        self.first_sys_connection_idx = None
        self.second_sys_connection_idx = None
        self._cls = None
        self._args = ()
        self._kwargs = {}

    def set_index(self, first_idx, second_idx):
        assert type(first_idx) is type(second_idx), "connection index types differ"
        self.first_sys_connection_idx = first_idx
        self.second_sys_connection_idx = second_idx

    def using(self, cls, *args, **kwargs):
        from .joint import FreeJoint

        assert issubclass(cls, FreeJoint), "not a joint class"
        self._cls = cls
        self._args = args
        self._kwargs = kwargs
        return self

    def id(self):
        return (
            self._first_sys_idx,
            self._second_sys_idx,
            self.first_sys_connection_idx,
            self.second_sys_connection_idx,
        )

    def __call__(self):
        return self._cls(*self._args, **self._kwargs)


class Connections:
    def __init__(self):
        self._connections = []
        super(Connections, self).__init__()
        self._feature_group_synchronize.append(self._call_connections)
        self._feature_group_finalize.append(self._finalize_connections)

    def connect(self, first_rod, second_rod, first_connect_idx=0, second_connect_idx=-1):
        sys_idx = [
            self._get_sys_idx_if_valid(first_rod),
            self._get_sys_idx_if_valid(second_rod),
        ]
        _connector = _Connect(*sys_idx)
        _connector.set_index(first_connect_idx, second_connect_idx)
        self._connections.append(_connector)
        return _connector

    def _finalize_connections(self):
        # registration order is execution order (never sorted)
        self._connections[:] = [(*c.id(), c()) for c in self._connections]

    def _call_connections(self, *args, **kwargs):
        for first_sys, second_sys, first_idx, second_idx, connection in self._connections:
            connection.apply_forces(
                self._systems[first_sys], first_idx, self._systems[second_sys], second_idx
            )
            connection.apply_torques(
                self._systems[first_sys], first_idx, self._systems[second_sys], second_idx
            )


# ------------------------------------------------------------------------------ forcing
class _ExtForceTorque(_Deferred):
    def using(self, cls, *args, **kwargs):
        from .external_forces import NoForces

        assert issubclass(cls, NoForces), "not a forcing class"
        return super().using(cls, *args, **kwargs)

    def __call__(self):
        return self._cls(*self._args, **self._kwargs)


class Forcing:
    def __init__(self):
        self._ext_forces_torques = []
        super(Forcing, self).__init__()
        self._feature_group_synchronize.append(self._call_ext_forces_torques)
        self._feature_group_finalize.append(self._finalize_forcing)

    def add_forcing_to(self, system):
        sys_idx = self._get_sys_idx_if_valid(system)
        _ext_force_torque = _ExtForceTorque(sys_idx)
        self._ext_forces_torques.append(_ext_force_torque)
        return _ext_force_torque

    def _finalize_forcing(self):
        self._ext_forces_torques[:] = [(f.id(), f()) for f in self._ext_forces_torques]
        self._ext_forces_torques.sort(key=lambda x: x[0])

    def _call_ext_forces_torques(self, time, *args, **kwargs):
        for sys_id, ext_force_torque in self._ext_forces_torques:
            ext_force_torque.apply_forces(self._systems[sys_id], time, *args, **kwargs)
            ext_force_torque.apply_torques(self._systems[sys_id], time, *args, **kwargs)


# ------------------------------------------------------------------------------ damping
class _Damper(_Deferred):
    def using(self, cls, *args, **kwargs):
        from .dissipation import DamperBase

        assert issubclass(cls, DamperBase), "not a damper class"
        return super().using(cls, *args, **kwargs)

    def __call__(self, rod):
        return self._cls(*self._args, _system=rod, **self._kwargs)


class Damping:
    def __init__(self):
        self._dampers = []
        super(Damping, self).__init__()
        self._feature_group_constrain_rates.append(self._dampen_rates)
        self._feature_group_finalize.append(self._finalize_dampers)

    def dampen(self, system):
        sys_idx = self._get_sys_idx_if_valid(system)
        _damper = _Damper(sys_idx)
        self._dampers.append(_damper)
        return _damper

    def _finalize_dampers(self):
        self._dampers[:] = [(d.id(), d(self._systems[d.id()])) for d in self._dampers]
        self._dampers.sort(key=lambda x: x[0])

    def _dampen_rates(self, time, *args, **kwargs):
        for sys_id, damper in self._dampers:
            damper.dampen_rates(self._systems[sys_id], time)


# ------------------------------------------------------------------------------ callbacks
class _CallBack(_Deferred):
    def using(self, cls, *args, **kwargs):
        from .callback_functions import CallBackBaseClass

        assert issubclass(cls, CallBackBaseClass), "not a callback class"
        return super().using(cls, *args, **kwargs)

    def __call__(self):
        return self._cls(*self._args, **self._kwargs)


class CallBacks:
    def __init__(self):
        self._callback_list = []
        super(CallBacks, self).__init__()
        self._feature_group_callback.append(self._callback_execution)
        self._feature_group_finalize.append(self._finalize_callback)

    def collect_diagnostics(self, system):
        sys_idx = self._get_sys_idx_if_valid(system)
        _callbacks = _CallBack(sys_idx)
        self._callback_list.append(_callbacks)
        return _callbacks

    def _finalize_callback(self):
        self._callback_list[:] = [(c.id(), c()) for c in self._callback_list]
        self._callback_list.sort(key=lambda x: x[0])

    def _callback_execution(self, time, current_step, *args, **kwargs):
        for sys_id, callback in self._callback_list:
            callback.make_callback(self._systems[sys_id], time, current_step, *args, **kwargs)
