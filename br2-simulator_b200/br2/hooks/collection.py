"""System collection + feature mixins with the PyElastica hook surface the reference uses:

    class BR2Simulator(BaseSystemCollection, Constraints, Connections, Forcing, Damping, CallBacks)
        (/root/reference/br2/free_simulator.py:36-38)
    simulator.append(rod)                                           (:299)
    simulator.dampen(rod).using(cls, **kw)                          (:316)
    simulator.constrain(rod).using(cls, constrained_position_idx=..., constrained_director_idx=..., **kw)   (:324)
    simulator.add_forcing_to(rod).using(cls, *a, **kw)              (:335, :351, :358)
    simulator.connect(first_rod, second_rod, first_connect_idx, second_connect_idx).using(cls, **kw)  (:418-435, :451-461)
    simulator.collect_diagnostics(rod).using(cls, step_skip=..., callback_params=...)  (:397)
    simulator.finalize()                                            (/root/reference/br2/environment.py:218)

Unlike upstream, ``finalize()`` does not set up host-side dispatch lists to be walked every step:
it instantiates the operators, *lowers* them to flat device tables (``br2._lowering``) and binds a
``br2._engine.Engine`` that owns the batch state in HBM.  Execution order inside a step is fixed by
the kernel to the order this exact mixin composition produces upstream:
forcing (stably sorted by rod) then connections (registration order) inside ``synchronize``;
dampers then constraint rates inside ``constrain_rates``.
"""
from collections.abc import MutableSequence

import numpy as np

from .rod import RodBase


class _Pending:
    """What ``xxx(system).using(cls, *args, **kwargs)`` records until ``finalize()``."""

    def __init__(self, *key):
        self.key = key
        self.cls = None
        self.args = ()
        self.kwargs = {}

    def using(self, cls, *args, **kwargs):
        self.cls, self.args, self.kwargs = cls, args, kwargs
        return self

    def require_class(self, what):
        if self.cls is None:
            raise RuntimeError("%s registered without .using(<class>, ...)" % what)


class BaseSystemCollection(MutableSequence):
    def __init__(self):
        self._finalizers = []
        self._systems = []
        self._finalize_flag = False
        self.engine = None
        self.batch = 1
        self.device = None
        super().__init__()
        self.allowed_sys_types = (RodBase,)

    # MutableSequence protocol ---------------------------------------------------------
    def _check_type(self, system):
        if not isinstance(system, self.allowed_sys_types):
            raise TypeError("%r is not a rod this collection can integrate" % type(system))

    def __len__(self):
        return len(self._systems)

    def __getitem__(self, idx):
        return self._systems[idx]

    def __delitem__(self, idx):
        self._require_open()
        del self._systems[idx]

    def __setitem__(self, idx, system):
        self._require_open()
        self._check_type(system)
        self._systems[idx] = system

    def insert(self, idx, system):
        self._require_open()
        self._check_type(system)
        self._systems.insert(idx, system)

    def extend_allowed_types(self, additional_types):
        self.allowed_sys_types += tuple(additional_types)

    def override_allowed_types(self, allowed_types):
        self.allowed_sys_types = tuple(allowed_types)

    def _require_open(self):
        if self._finalize_flag:
            raise RuntimeError("the collection is finalized; nothing can be added or removed")

    def _get_sys_idx_if_valid(self, system):
        if isinstance(system, (int, np.integer)):
            if not -len(self._systems) <= system < len(self._systems):
                raise IndexError("system index out of range")
            return int(system) % len(self._systems)
        self._check_type(system)
        for idx, candidate in enumerate(self._systems):
            if candidate is system:
                return idx
        raise ValueError("rod was not appended to this collection")

    # life cycle -----------------------------------------------------------------------
    def finalize(self, batch=None, device=None):
        """Instantiate the registered operators, lower them to device tables, allocate the batch
        state in HBM and bind the rods to it.  ``batch`` / ``device`` default to the attributes of
        the same name (set by ``Environment``)."""
        if self._finalize_flag:
            raise RuntimeError("The finalize cannot be called twice.")
        from .._engine import Engine

        if batch is not None:
            self.batch = int(batch)
        if device is not None:
            self.device = device
        self.lower()
        self.engine = Engine(self.program, batch=self.batch, device=self.device)
        self.engine.load_template(self._systems)
        for idx, rod in enumerate(self._systems):
            rod._bind(self.engine, idx)
        self._finalize_flag = True

    def lower(self):
        """Host-only half of ``finalize()``: instantiate the registered operators and lower them to
        the device tables (``self.program``).  Needs no GPU."""
        if getattr(self, "program", None) is None:
            from .._lowering import lower_collection

            for fin in self._finalizers:
                fin()
            self.program = lower_collection(self)
        return self.program

    def kernel_plan(self):
        """Which device kernel would integrate this collection and how (one line; needs no GPU)."""
        from .. import _native

        return _native.describe_plan(self.lower())

    def snapshot_views(self):
        """Host snapshot of every rod (one D2H copy of the batch state) for diagnostics callbacks."""
        from .rod import HostRodView

        return [HostRodView(rod, arrays) for rod, arrays in zip(self._systems, self.engine.snapshot())]

    def callbacks_due(self, current_step):
        for _, op in self._callback_ops:
            every = op.capture_steps()
            if every is None or current_step % every == 0:
                return True
        return False

    # introspection used by the lowering (empty unless the mixin is composed in)
    _forcing = ()
    _constraint_ops = ()
    _damper_ops = ()
    _joints = ()
    _callback_ops = ()


class Constraints:
    def __init__(self):
        self._pending_constraints = []
        self._constraint_ops = []
        super().__init__()
        self._finalizers.append(self._finalize_constraints)

    def constrain(self, system):
        self._require_open()
        pending = _Pending(self._get_sys_idx_if_valid(system))
        self._pending_constraints.append(pending)
        return pending

    def _finalize_constraints(self):
        ops = []
        for pending in self._pending_constraints:
            pending.require_class("constraint")
            (rod_idx,) = pending.key
            rod = self._systems[rod_idx]
            kw = pending.kwargs
            # upstream hands copies of x[:, idx] / Q[:, :, idx] to the constructor as leading
            # positional arguments (free_custom_systems.py:105 relies on it)
            fixed_x = [rod.position_collection[..., i].copy() for i in kw.get("constrained_position_idx", ())]
            fixed_Q = [rod.director_collection[..., i].copy() for i in kw.get("constrained_director_idx", ())]
            ops.append((rod_idx, pending.cls(*fixed_x, *fixed_Q, *pending.args, _system=rod, **kw)))
        ops.sort(key=lambda item: item[0])  # stable
        self._constraint_ops = ops


class Connections:
    def __init__(self):
        self._pending_joints = []
        self._joints = []
        super().__init__()
        self._finalizers.append(self._finalize_connections)

    def connect(self, first_rod, second_rod, first_connect_idx=0, second_connect_idx=-1):
        self._require_open()
        if type(first_connect_idx) is not type(second_connect_idx):
            raise TypeError("first_connect_idx and second_connect_idx must have the same type")
        pending = _Pending(self._get_sys_idx_if_valid(first_rod), self._get_sys_idx_if_valid(second_rod),
                           first_connect_idx, second_connect_idx)
        self._pending_joints.append(pending)
        return pending

    def _finalize_connections(self):
        joints = []
        for pending in self._pending_joints:  # registration order IS execution order
            pending.require_class("connection")
            joints.append((*pending.key, pending.cls(*pending.args, **pending.kwargs)))
        self._joints = joints


class Forcing:
    def __init__(self):
        self._pending_forcing = []
        self._forcing = []
        super().__init__()
        self._finalizers.append(self._finalize_forcing)

    def add_forcing_to(self, system):
        self._require_open()
        pending = _Pending(self._get_sys_idx_if_valid(system))
        self._pending_forcing.append(pending)
        return pending

    def _finalize_forcing(self):
        ops = []
        for pending in self._pending_forcing:
            pending.require_class("forcing")
            ops.append((pending.key[0], pending.cls(*pending.args, **pending.kwargs)))
        ops.sort(key=lambda item: item[0])  # stable: per rod, registration order
        self._forcing = ops


class Damping:
    def __init__(self):
        self._pending_dampers = []
        self._damper_ops = []
        super().__init__()
        self._finalizers.append(self._finalize_dampers)

    def dampen(self, system):
        self._require_open()
        pending = _Pending(self._get_sys_idx_if_valid(system))
        self._pending_dampers.append(pending)
        return pending

    def _finalize_dampers(self):
        ops = []
        for pending in self._pending_dampers:
            pending.require_class("damper")
            rod = self._systems[pending.key[0]]
            ops.append((pending.key[0], pending.cls(*pending.args, _system=rod, **pending.kwargs)))
        ops.sort(key=lambda item: item[0])
        self._damper_ops = ops


class CallBacks:
    def __init__(self):
        self._pending_callbacks = []
        self._callback_ops = []
        super().__init__()
        self._finalizers.append(self._finalize_callbacks)

    def collect_diagnostics(self, system):
        self._require_open()
        pending = _Pending(self._get_sys_idx_if_valid(system))
        self._pending_callbacks.append(pending)
        return pending

    def _finalize_callbacks(self):
        ops = []
        for pending in self._pending_callbacks:
            pending.require_class("callback")
            ops.append((pending.key[0], pending.cls(*pending.args, **pending.kwargs)))
        ops.sort(key=lambda item: item[0])
        self._callback_ops = ops
