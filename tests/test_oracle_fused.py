"""The plain-C fused oracle against the faithful (per-operator numba) oracle.

Two library calls in the faithful path round in a library-dependent way: numpy's vectorised
``np.power`` (damper, SVML on AVX-512 hosts) and the BLAS-backed ``@`` of the tip-to-base joint.
With those two swapped for libm ``pow`` / an explicit j=0,1,2 accumulation the C oracle must be
BIT-IDENTICAL to the faithful path; with them left alone the two must agree to the rounding-noise
floor of the dynamics (~5e-13 positions, ~2e-12 directors)."""
import math
import os

import numpy as np
import pytest
from numba import njit

from conftest import DATA
from oracle import br2_port
from oracle.br2_port import ops
from oracle.fused import FusedOracle
from oracle.mini_elastica import dissipation

ROD_DB = os.path.join(DATA, "rod_library.json")
PSI = 6895.0
NAMES = ["action1", "action2", "action3"]
CASES = [
    ("single", {"gravity": True}, {"action1": 30 * PSI, "action2": 30 * PSI, "action3": 0.0}, 600),
    ("double", {}, {"action1": 25 * PSI, "action2": 10 * PSI, "action3": 20 * PSI}, 400),
    ("triple", {"gravity": True, "nu_multiplier": 2.0}, {"action1": 15 * PSI, "action2": 30 * PSI, "action3": 5 * PSI}, 300),
]
PAIRS = (("x", "position_collection"), ("v", "velocity_collection"), ("Q", "director_collection"),
         ("w", "omega_collection"), ("a", "acceleration_collection"), ("f_ext", "external_forces"),
         ("t_ext", "external_torques"), ("f_int", "internal_forces"), ("t_int", "internal_torques"),
         ("len", "lengths"), ("dil", "dilatation"), ("rad", "radius"))


def _libm_dampen_rates(self, rod, time):
    rod.velocity_collection[:] = rod.velocity_collection * self.translational_damping_coefficient
    c, e = self.rotational_damping_coefficient, rod.dilatation
    p = np.array([[math.pow(c[i, k], e[k]) for k in range(e.shape[0])] for i in range(3)])
    rod.omega_collection[:] = rod.omega_collection * p


@njit(cache=True)
def _loop_tip_to_base_force(k, nu, l1, l2, x1, x2, r1, r2, v1, v2, Q1, Q2, f1, f2):
    n1, n2 = Q1.shape[2], Q2.shape[2]
    for i in range(3):
        a1 = 0.0
        a2 = 0.0
        for j in range(3):
            a1 += Q1[j, i, n1 - 1] * (l1[j] * r1[0, n1 - 1])
            a2 += Q2[j, i, 0] * (l2[j] * r2[0, 0])
        c1 = 0.5 * (x1[i, n1] + x1[i, n1 - 1])
        c2 = 0.5 * (x2[i, 0] + x2[i, 1])
        gap = (c2 + a2) - (c1 + a1)
        u1 = 0.5 * (v1[i, n1] + v1[i, n1 - 1])
        u2 = 0.5 * (v2[i, 0] + v2[i, 1])
        total = k * gap + (-nu * (u2 - u1))
        f1[i, n1] += 0.5 * total
        f1[i, n1 - 1] += 0.5 * total
        f2[i, 0] -= 0.5 * total
        f2[i, 1] -= 0.5 * total
    return 0.0, 0.0, 0.0


def _run_both(assembly, kwargs, action, n_steps):
    env = br2_port.OracleEnvironment()
    env.reset(ROD_DB, os.path.join(DATA, "assembly_%s.json" % assembly), **kwargs)
    fused = FusedOracle.from_simulator(env.simulator, NAMES)
    W = fused.pack(batch=2)
    pressures = np.array([[action[k], 0.5 * action[k]] for k in NAMES])
    t_end = fused.run(W, pressures, n_steps, 0.0, env.time_step)
    seen = {}

    class Grab:  # f_ext / t_ext as the callbacks see them (before the end-of-step zeroing)
        def make_callback(self, system, time, current_step):
            if current_step == n_steps:
                seen[id(system)] = (system.external_forces.copy(), system.external_torques.copy())

    env.simulator._callback_list[:] = [(i, Grab()) for i in range(len(env.simulator))]
    env.run(action, n_steps=n_steps)
    assert t_end == env.time
    out = {}
    for r, rod in enumerate(env.simulator):
        for short, attr in PAIRS:
            want = getattr(rod, attr)
            if short == "f_ext":
                want = seen[id(rod)][0]
            elif short == "t_ext":
                want = seen[id(rod)][1]
            out[(r, short)] = (fused.view(W, r, short)[0].copy(), want.copy())
    # second instance ran with halved pressures: it must differ (batching is real)
    assert np.abs(fused.view(W, 0, "x")[1] - fused.view(W, 0, "x")[0]).max() > 1e-9
    return out


@pytest.mark.parametrize("assembly,kwargs,action,n_steps", CASES, ids=[c[0] for c in CASES])
def test_fused_c_is_bit_identical_with_libm_rounding(assembly, kwargs, action, n_steps, monkeypatch):
    monkeypatch.setattr(dissipation.AnalyticalLinearDamper, "dampen_rates", _libm_dampen_rates)
    monkeypatch.setattr(ops, "_tip_to_base_force", _loop_tip_to_base_force)
    for (r, short), (got, want) in _run_both(assembly, kwargs, action, n_steps).items():
        np.testing.assert_array_equal(got, want, err_msg="rod %d field %s" % (r, short))


@pytest.mark.parametrize("assembly,kwargs,action,n_steps", CASES, ids=[c[0] for c in CASES])
def test_fused_c_matches_faithful_to_noise_floor(assembly, kwargs, action, n_steps):
    res = _run_both(assembly, kwargs, action, n_steps)
    x_scale = max(np.abs(want).max() for (r, s), (got, want) in res.items() if s == "x")
    for (r, short), (got, want) in res.items():
        err = np.abs(got - want).max()
        if short == "x":
            assert err / x_scale <= 1e-11, (r, short, err)
        elif short == "Q":
            assert err <= 1e-10, (r, short, err)
        elif short in ("len", "rad", "dil"):
            assert err / np.abs(want).max() <= 1e-11, (r, short, err)


# tip of RodBend after 5000 / 10000 steps of single_br2 (gravity on, action1 = action2 = 30 psi,
# dt = 2e-5), measured by the survey's own throw-away mimic of the reference -- an implementation independent of this
# repository's oracle.  Not a golden vector of the reference (it ships none), but a second opinion: five digits agree.
SURVEY_ANCHORS = ((5000, (-0.00470, 0.17958, 0.00057)), (10000, (-0.01113, 0.17870, 0.00216)))


def test_oracle_reproduces_the_surveys_independent_probe_anchors():
    ref = br2_port.OracleEnvironment()
    ref.reset(ROD_DB, os.path.join(DATA, "assembly_single.json"), gravity=True)
    fused = FusedOracle.from_simulator(ref.simulator, NAMES)
    W = fused.pack(1)
    P = np.array([[30 * PSI], [30 * PSI], [0.0]])
    done = 0
    for n_steps, tip in SURVEY_ANCHORS:
        fused.run(W, P, n_steps - done, done * 2.0e-5, 2.0e-5)
        done = n_steps
        got = fused.view(W, 0, "x")[0][:, -1]
        assert np.abs(got - np.array(tip)).max() < 6e-6, (n_steps, got)  # the anchors are quoted to 1e-5 m
    assert float(np.abs(fused.view(W, 0, "v")).max()) < 0.12  # "max |v| ~ 6e-2 m/s"
