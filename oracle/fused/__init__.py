"""ctypes front end of the plain-C fused oracle (TEST INFRASTRUCTURE).

``FusedOracle.from_simulator(sim, actions)`` walks a *finalized* ``oracle.br2_port.BR2Simulator``
(the faithful oracle's own objects -- it never looks at the product's tables) and flattens its
rods and registered operators, in execution order, into the tables ``br2_oracle.c`` interprets.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libbr2oracle.so")
LIB_F32 = os.path.join(HERE, "libbr2oracle_f32.so")  # same code in single precision: precision study only, never a checker

ROD_NI, F_NI, F_ND, C_NI, C_ND, J_NI, J_ND = 4, 4, 6, 2, 14, 5, 10
STATE = ("x", "v", "Q", "w", "a", "alpha", "f_int", "t_int", "f_ext", "t_ext", "len", "dil", "rad")
_ATTR = {
    "x": "position_collection", "v": "velocity_collection", "Q": "director_collection",
    "w": "omega_collection", "a": "acceleration_collection", "alpha": "alpha_collection",
    "f_int": "internal_forces", "t_int": "internal_torques", "f_ext": "external_forces",
    "t_ext": "external_torques", "len": "lengths", "dil": "dilatation", "rad": "radius",
}


def build(force=False):
    src = os.path.join(HERE, "br2_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "-B", "libbr2oracle.so"])
    if force or not os.path.exists(LIB_F32) or os.path.getmtime(LIB_F32) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "-B", "libbr2oracle_f32.so"])
    return LIB


class _Program(ctypes.Structure):
    _fields_ = [
        ("n_rods", ctypes.c_int32), ("n_forcing", ctypes.c_int32), ("n_constraints", ctypes.c_int32),
        ("n_joints", ctypes.c_int32), ("n_actions", ctypes.c_int32),
        ("words_per_instance", ctypes.c_int64),
        ("rod_i", ctypes.c_void_p), ("consts", ctypes.c_void_p),
        ("forcing_i", ctypes.c_void_p), ("forcing_d", ctypes.c_void_p),
        ("cons_i", ctypes.c_void_p), ("cons_d", ctypes.c_void_p),
        ("joint_i", ctypes.c_void_p), ("joint_d", ctypes.c_void_p),
    ]


def _sizes(n):
    return {"x": 3 * (n + 1), "v": 3 * (n + 1), "Q": 9 * n, "w": 3 * n, "a": 3 * (n + 1), "alpha": 3 * n,
            "f_int": 3 * (n + 1), "t_int": 3 * n, "f_ext": 3 * (n + 1), "t_ext": 3 * n,
            "len": n, "dil": n, "rad": n}


def _diag(m):
    return np.ascontiguousarray(np.stack([m[0, 0], m[1, 1], m[2, 2]]))


class FusedOracle:
    def __init__(self, real=np.float64):
        """``real=np.float32`` loads the single-precision build of the same C code (how far would an fp32 path drift?)."""
        self.real = np.dtype(real)
        build()
        self._lib = ctypes.CDLL(LIB if self.real == np.float64 else LIB_F32)
        self._lib.br2o_run.restype = ctypes.c_double
        self._lib.br2o_run.argtypes = [ctypes.POINTER(_Program), ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_int64, ctypes.c_int64, ctypes.c_double, ctypes.c_double,
                                       ctypes.c_int32]

    @classmethod
    def from_simulator(cls, simulator, action_names, real=np.float64):
        """``action_names``: ordered list of action keys; pressures are passed in that order."""
        from ..br2_port import ops
        from ..mini_elastica import AnalyticalLinearDamper, GravityForces
        from ..mini_elastica.experimental.connection_contact_joint.parallel_connection import (
            SurfaceJointSideBySide,
        )

        self = cls(real)
        self.action_names = list(action_names)
        rods = list(simulator._systems)
        self.rods = rods
        rod_i = np.zeros((len(rods), ROD_NI), dtype=np.int64)
        consts, w_off = [], 0
        self.layout = []
        dampers = {sys_id: d for sys_id, d in simulator._dampers}
        for r, rod in enumerate(rods):
            n = rod.n_elems
            off, fields = w_off, {}
            for name in STATE:
                fields[name] = (off, _sizes(n)[name])
                off += _sizes(n)[name]
            self.layout.append(fields)
            flags = 0
            c_r, c_t = np.ones((3, n)), 1.0
            if r in dampers:
                assert isinstance(dampers[r], AnalyticalLinearDamper)
                flags |= 1
                c_r = np.asarray(dampers[r].rotational_damping_coefficient)
                c_t = float(dampers[r].translational_damping_coefficient)
            for m in (rod.shear_matrix, rod.bend_matrix, rod.mass_second_moment_of_inertia,
                      rod.inv_mass_second_moment_of_inertia):
                offdiag = m.copy()
                for i in range(3):
                    offdiag[i, i] = 0.0
                assert not offdiag.any(), "fused oracle handles diagonal constitutive matrices only"
            block = np.concatenate([
                rod.rest_lengths, rod.rest_voronoi_lengths, rod.volume, rod.mass,
                _diag(rod.shear_matrix).ravel(), _diag(rod.bend_matrix).ravel(),
                _diag(rod.mass_second_moment_of_inertia).ravel(),
                _diag(rod.inv_mass_second_moment_of_inertia).ravel(),
                rod.rest_sigma.ravel(), rod.rest_kappa.ravel(), np.ascontiguousarray(c_r).ravel(), [c_t],
            ])
            rod_i[r] = (n, w_off, sum(len(c) for c in consts), flags)
            consts.append(block)
            w_off = off
        self.words = w_off

        def rod_index(rod):
            return next(i for i, r in enumerate(rods) if r is rod)

        f_i, f_d = [], []
        action_of = {}
        for name in self.action_names:
            action_of[id(simulator_actuation(simulator, name))] = self.action_names.index(name)
        for sys_id, op in simulator._ext_forces_torques:
            d = np.zeros(F_ND)
            if isinstance(op, GravityForces):
                d[:3] = op.acc_gravity
                f_i.append((0, sys_id, -1, 0))
            elif isinstance(op, ops.FreeTwistActuation):
                d[0], d[1] = op.scale, op.ramp_up_time
                f_i.append((1, sys_id, action_of[id(op.actuation_ref)], 0))
            elif isinstance(op, ops.FreeBendActuation):
                d[0], d[1], d[2] = op.magnitude_scale, op.ramp_up_time, op.z_angle
                f_i.append((2, sys_id, action_of[id(op.actuation_ref)], 0))
            elif isinstance(op, ops.PointForces):
                d[:3], d[3] = op.force, op.ramp_up_time
                node = int(rods[sys_id].external_forces.shape[-1] * op.location)
                f_i.append((3, sys_id, -1, node))
            else:
                raise TypeError("fused oracle: unsupported forcing %r" % type(op))
            f_d.append(d)
        c_i, c_d = [], []
        for sys_id, op in simulator._constraints:
            d = np.zeros(C_ND)
            d[:3] = np.asarray(op.fixed_position).ravel()
            d[3:12] = np.asarray(op.fixed_directors).reshape(3, 3).ravel()
            if isinstance(op, ops.FreeBaseEndSoftFixed):
                d[12], d[13] = op.k, op.nu
                c_i.append((0, sys_id))
            elif isinstance(op, ops.LastEndFixedRod):
                c_i.append((1, sys_id))
            else:
                raise TypeError("fused oracle: unsupported constraint %r" % type(op))
            c_d.append(d)
        j_i, j_d = [], []
        for first, second, idx1, idx2, op in simulator._connections:
            d = np.zeros(J_ND)
            d[0], d[1] = float(op.k), float(op.nu)
            if isinstance(op, SurfaceJointSideBySide):
                d[2] = float(op.k_repulsive)
                d[3:6] = op.rod_one_direction_vec_in_material_frame
                d[6:9] = op.rod_two_direction_vec_in_material_frame
                d[9] = float(op.offset_btw_rods)
                j_i.append((0, first, second, idx1, idx2))
            elif isinstance(op, ops.TipToTipStraightJoint):
                d[3:6], d[6:9] = op.rod1_rd2_local, op.rod2_rd2_local
                j_i.append((1, first, second, idx1, idx2))
            else:
                raise TypeError("fused oracle: unsupported joint %r" % type(op))
            j_d.append(d)

        def arr(rows, width, dtype):
            return np.ascontiguousarray(np.array(rows, dtype=dtype).reshape(-1, width))

        self._keep = dict(
            rod_i=np.ascontiguousarray(rod_i), consts=np.ascontiguousarray(np.concatenate(consts).astype(self.real)),
            forcing_i=arr(f_i, F_NI, np.int64), forcing_d=arr(f_d, F_ND, self.real),
            cons_i=arr(c_i, C_NI, np.int64), cons_d=arr(c_d, C_ND, self.real),
            joint_i=arr(j_i, J_NI, np.int64), joint_d=arr(j_d, J_ND, self.real),
        )
        k = self._keep
        self._prog = _Program(
            len(rods), len(f_i), len(c_i), len(j_i), len(self.action_names), self.words,
            *[k[name].ctypes.data for name in ("rod_i", "consts", "forcing_i", "forcing_d",
                                              "cons_i", "cons_d", "joint_i", "joint_d")])
        return self

    # ------------------------------------------------------------------ state marshalling
    def pack(self, batch=1):
        """Workspace [batch, words] initialised from the oracle rods' current arrays."""
        one = np.zeros(self.words)
        for rod, fields in zip(self.rods, self.layout):
            for name, (off, size) in fields.items():
                one[off:off + size] = np.asarray(getattr(rod, _ATTR[name])).ravel()
        return np.ascontiguousarray(np.tile(one, (batch, 1)).astype(self.real))

    def view(self, W, rod, name):
        """[batch, ...] view of one field of one rod inside a workspace."""
        off, size = self.layout[rod][name]
        n = self.rods[rod].n_elems
        shape = {"Q": (3, 3, n), "len": (n,), "dil": (n,), "rad": (n,)}.get(name, (3, size // 3))
        return W[:, off:off + size].reshape((W.shape[0],) + shape)

    def run(self, W, pressures, n_steps, t0, dt, n_threads=0):
        """Advance every instance ``n_steps`` steps in place.  ``pressures``: [n_actions, batch]."""
        assert W.flags.c_contiguous and W.dtype == self.real and W.shape[1] == self.words
        pressures = np.ascontiguousarray(np.asarray(pressures, dtype=self.real).reshape(len(self.action_names), -1))
        assert pressures.shape[1] == W.shape[0]
        return self._lib.br2o_run(ctypes.byref(self._prog), W.ctypes.data, pressures.ctypes.data,
                                  W.shape[0], int(n_steps), float(t0), float(dt), int(n_threads))


def simulator_actuation(simulator, name):
    """The shared python list an action name maps to (free_simulator.py:283-289)."""
    return simulator._actuation_lists[name]
