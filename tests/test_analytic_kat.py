"""Analytic known-answer tests: textbook beam answers for a cantilevered rod.

The reference ships no tests, and the arithmetic of its hot path lives in PyElastica 0.3.1, which cannot be installed
here -- so nothing pins the oracle's restatement of it (`oracle/mini_elastica`) to PyElastica's own numbers.  These
tests pin it, and the CUDA path independently of the oracle, to the physics PyElastica itself is validated against:

* axial stretch under a tip force:      dL = F L / (E A)                                  (shear/stretch matrix, dilatation)
* twist under distributed couples:      phi = T L (n-1) / (2 n G I3)                      (bend/twist matrix on the Voronoi domain)
* transverse tip force, free vibration: mean deflection over one period = static answer
                                        F ds^3 / (E I) * sum_{m<n} m^2 + F L / (alpha_c G A),  swing ~ twice that
  (the sum is the discrete form of L^3 / 3: curvature lives on the n-1 interior nodes), alpha_c = 27/28.

AnalyticalLinearDamper damps omega by exp(-nu dt m/J) with m/J = 4/r^2 = 1.6e5 1/m^2 here, so any nu that settles a
translation freezes every rotation; hence the static cases use a nu suited to their own degree of freedom and the
bending case is dynamic and undamped.
"""
import numpy as np
import pytest

E, G, R, L, RHO = 2.0e6, 1.2e6, 0.005, 0.2, 1200.0
AREA, I1 = np.pi * R * R, np.pi * R ** 4 / 4.0
ALPHA_C = 27.0 / 28.0


def oracle_api():
    from oracle import br2_port, mini_elastica

    return dict(sim=br2_port.BR2Simulator, rod=mini_elastica.CosseratRod, damper=mini_elastica.AnalyticalLinearDamper,
                point=br2_port.PointForces, twist=br2_port.FreeTwistActuation, fixed=br2_port.FreeBaseEndSoftFixed,
                stepper=mini_elastica.PositionVerlet, extend=mini_elastica.extend_stepper_interface)


def product_api():
    from br2.custom_activation import PointForces
    from br2.free_custom_systems import FreeBaseEndSoftFixed, FreeTwistActuation
    from br2.free_simulator import BR2Simulator
    from br2.hooks import AnalyticalLinearDamper, CosseratRod, PositionVerlet, extend_stepper_interface

    return dict(sim=BR2Simulator, rod=CosseratRod, damper=AnalyticalLinearDamper, point=PointForces,
                twist=FreeTwistActuation, fixed=FreeBaseEndSoftFixed, stepper=PositionVerlet, extend=extend_stepper_interface)


def cantilever(api, n, dt, nu, force=None, couple=None):
    spec = dict(n_elements=n, start=np.zeros(3), direction=np.array([0.0, 1.0, 0.0]), normal=np.array([0.0, 0.0, 1.0]),
                base_length=L, base_radius=R, density=RHO, youngs_modulus=E, shear_modulus=G)
    sim = api["sim"]()
    rod = api["rod"].straight_rod(**spec)
    sim.append(rod)
    sim.dampen(rod).using(api["damper"], damping_constant=nu, time_step=dt)
    if force is not None:
        sim.add_forcing_to(rod).using(api["point"], np.asarray(force, dtype=np.float64), 0.999, ramp_up_time=1e-9)
    if couple is not None:
        sim.add_forcing_to(rod).using(api["twist"], actuation_ref=[couple], scale=1.0, ramp_up_time=1e-3)
    sim.constrain(rod).using(api["fixed"], constrained_position_idx=(0,), constrained_director_idx=(0,), k=1e9, nu=0.0, kt=0.0)
    sim.finalize()
    return sim, rod


def advance(api, sim, dt, n_steps, time=0.0, launch=None):
    stepper = api["stepper"]()
    do_step, stages = api["extend"](stepper, sim)
    if launch is not None:  # device path: n_steps in one launch
        return do_step(stepper, stages, sim, time, dt, n_steps=n_steps)
    for _ in range(n_steps):
        time = do_step(stepper, stages, sim, time, dt)
    return time


def check_stretch(api, device):
    n, dt, force = 10, 1.0e-5, 0.01
    sim, rod = cantilever(api, n, dt, nu=400.0, force=[0.0, force, 0.0])
    advance(api, sim, dt, 6000, launch=device)
    want = force * L / (E * AREA)
    got = rod.position_collection[1, -1] - L
    assert abs(got - want) <= 1e-3 * want, (got, want)          # the residual is O(strain) = 6e-5
    assert np.abs(rod.velocity_collection).max() < 1e-6


def check_twist(api, device):
    n, dt, couple = 10, 1.0e-5, 1.0e-5
    sim, rod = cantilever(api, n, dt, nu=4.0e-3, couple=couple)
    advance(api, sim, dt, 8000, launch=device)
    want = couple * L * (n - 1) / (2 * n * G * 2 * I1)         # every element carries couple / n; I3 = 2 I1
    Q = rod.director_collection[:, :, -1]
    got = np.arctan2(Q[0, 0], Q[0, 2])                          # d1 starts along z and turns towards x
    assert abs(got - want) <= 1e-3 * want, (got, want)


def check_bending(api, device):
    n, dt, force = 20, 5.0e-5, 1.0e-4
    sim, rod = cantilever(api, n, dt, nu=1e-9, force=[0.0, 0.0, force])
    ds = L / n
    static = force * ds ** 3 / (E * I1) * sum(m * m for m in range(1, n)) + force * L / (ALPHA_C * G * AREA)
    tip, time = [], 0.0
    stride = 50 if device else 1
    for _ in range(16000 // stride):
        time = advance(api, sim, dt, stride, time=time, launch=device)
        tip.append(rod.position_collection[2, -1])
    tip = np.array(tip)
    half = int(np.argmax(tip))
    back = half + int(np.argmin(tip[half:2 * half + 3000 // stride]))   # first return to (almost) rest: one period
    period = (back + 1) * stride * dt
    euler_bernoulli = 1.8751 ** 2 / (2 * np.pi) * np.sqrt(E * I1 / (RHO * AREA * L ** 4))
    assert abs(1.0 / period - euler_bernoulli) <= 0.08 * euler_bernoulli     # 20 elements are 7 % stiffer than the continuum
    # a step load rings mostly the first mode (97 % of the static deflection; the higher modes make the swing a little
    # irregular): it reaches almost twice the static answer, and the average over one period is the static answer
    assert abs(tip.max() - 2 * static) <= 0.03 * 2 * static, (tip.max(), 2 * static)
    assert abs(tip[:back + 1].mean() - static) <= 0.02 * static, (tip[:back + 1].mean(), static)


# ---- the oracle (CPU) ----------------------------------------------------------------------------
def test_oracle_axial_stretch_matches_hooke():
    check_stretch(oracle_api(), device=None)


def test_oracle_twist_matches_torsion_theory():
    check_twist(oracle_api(), device=None)


def test_oracle_cantilever_vibration_matches_timoshenko_statics():
    check_bending(oracle_api(), device=None)


# ---- the CUDA path, independently of the oracle ----------------------------------------------------
@pytest.mark.gpu
def test_device_axial_stretch_matches_hooke():
    check_stretch(product_api(), device=True)


@pytest.mark.gpu
def test_device_twist_matches_torsion_theory():
    check_twist(product_api(), device=True)


@pytest.mark.gpu
def test_device_cantilever_vibration_matches_timoshenko_statics():
    check_bending(product_api(), device=True)
