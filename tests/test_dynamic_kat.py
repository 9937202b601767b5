"""Known-answer tests for the DYNAMIC coefficients of the Cosserat layer (masses, inertias, wave speeds, damper laws) --
the statics tests (tests/test_convergent_kat.py, tests/test_analytic_kat.py) cannot see them.  Like those, they hold the
restatement of PyElastica 0.3.1 (`oracle/mini_elastica`, parity unpinned: the package cannot be installed here) and,
independently, the CUDA path to closed-form answers of the DISCRETE equations, so the bounds are tight:

* axial standing wave of a free rod: nodes of mass rho A ds (half at the ends) on springs E A / ds; the first free-free
  mode cos(pi i / n) is an exact eigenvector with omega = (2 / ds) sqrt(E / rho) sin(pi / 2n); position Verlet turns it
  into omega_h with cos(omega_h dt) = 1 - (omega dt)^2 / 2.  Pins E A (the stretch entry of the shear matrix -- no shear
  correction factor on it) against the nodal masses;
* torsional standing wave: elements of inertia rho I3 ds on torsional springs G I3 / ds (Voronoi rest length ds), mode
  cos(pi (i + 1/2) / n), omega = (2 / ds) sqrt(G / rho) sin(pi / 2n).  Pins G I3 (twist entry of the bend matrix, polar
  moment I3 = 2 I1) against the mass second moment of inertia, and the sign / order of the director update;
* AnalyticalLinearDamper on a rod that drifts and spins about its axis without deforming: v(t) = v0 exp(-nu t) and
  omega3(t) = omega3(0) exp(-nu t m_e / J3) with m_e / J3 = A / I3 = 2 / r^2 (upstream: translational coefficient
  exp(-nu dt), rotational exp(-nu dt m_e J^-1) to the power of the dilatation), to 1e-9 (rounding-level ringing).
"""
import numpy as np
import pytest

from test_convergent_kat import AREA, E, G, L, R, RHO, advance, oracle_api, product_api, straight


def verlet_frequency(omega, dt):
    return np.arccos(1.0 - 0.5 * (omega * dt) ** 2) / dt


# ---------------------------------------------------------------------------------------------------- axial wave
def check_axial_wave(api, device):
    n, dt, n_steps, amp = 20, 1.0e-5, 1300, 1.0e-4
    ds = L / n
    rod = straight(api, n)
    x0 = np.array(rod.position_collection)
    shape = np.cos(np.pi * np.arange(n + 1) / n)
    v0 = np.zeros((3, n + 1))
    v0[1] = amp * shape
    rod.velocity_collection = v0
    sim = api["sim"]()
    sim.append(rod)
    sim.finalize()
    t = advance(api, sim, dt, n_steps, device)
    omega = 2.0 / ds * np.sqrt(E / RHO) * np.sin(np.pi / (2 * n))
    omega_h = verlet_frequency(omega, dt)
    assert 1.2 < omega_h * t / (2 * np.pi) < 1.5          # a good part of the second period: phase errors show
    weights = np.ones(n + 1)
    weights[[0, -1]] = 0.5                                 # nodal masses rho A ds, half at the ends
    norm = (weights * shape * shape).sum()
    v = np.asarray(rod.velocity_collection)
    u = np.asarray(rod.position_collection) - x0
    modal_v = (weights * shape * v[1]).sum() / norm
    modal_u = (weights * shape * u[1]).sum() / norm
    # the drift-kick-drift scheme samples positions half a step off the velocities: u(t) = (amp / (omega_h cos(omega_h dt / 2))) sin(omega_h t)
    # to second order; both bounds are 1e-5 of the amplitude (strain amplitude 2.5e-6: the elastic law's own nonlinearity)
    assert abs(modal_v / amp - np.cos(omega_h * t)) <= 1e-5, (modal_v / amp, np.cos(omega_h * t))
    assert abs(modal_u * omega_h / amp - np.sin(omega_h * t)) <= 2e-5, (modal_u * omega_h / amp, np.sin(omega_h * t))
    # nothing leaks out of the mode or out of the axis
    assert np.abs(v[1] - modal_v * shape).max() <= 1e-5 * amp
    assert np.abs(v[[0, 2]]).max() <= 1e-12 * amp + 1e-18
    # a wrong modulus or mass would show at full size: E -> alpha_c E (27/28) alone shifts cos(omega t) by 0.13
    wrong = verlet_frequency(omega * np.sqrt(27.0 / 28.0), dt)
    assert abs(np.cos(wrong * t) - np.cos(omega_h * t)) > 0.05


def test_oracle_axial_wave_frequency():
    check_axial_wave(oracle_api(), device=False)


@pytest.mark.gpu
def test_device_axial_wave_frequency():
    check_axial_wave(product_api(), device=True)


# ---------------------------------------------------------------------------------------------------- torsional wave
def check_torsional_wave(api, device):
    n, dt, n_steps, amp = 20, 1.0e-5, 1700, 1.0
    ds = L / n
    rod = straight(api, n)
    shape = np.cos(np.pi * (np.arange(n) + 0.5) / n)
    w0 = np.zeros((3, n))
    w0[2] = amp * shape                                    # spin about d3, the rod's axis
    rod.omega_collection = w0
    sim = api["sim"]()
    sim.append(rod)
    sim.finalize()
    t = advance(api, sim, dt, n_steps, device)
    omega = 2.0 / ds * np.sqrt(G / RHO) * np.sin(np.pi / (2 * n))
    omega_h = verlet_frequency(omega, dt)
    assert 1.2 < omega_h * t / (2 * np.pi) < 1.5
    w = np.asarray(rod.omega_collection)
    modal_w = (shape * w[2]).sum() / (shape * shape).sum()
    # (the reference's 1e-14 guard on the rotation angle shortens every rotation by 1e-14 / |omega dt / 2| <= 1e-8 here)
    assert abs(modal_w / amp - np.cos(omega_h * t)) <= 1e-5, (modal_w / amp, np.cos(omega_h * t))
    assert np.abs(w[2] - modal_w * shape).max() <= 1e-5 * amp
    assert np.abs(w[:2]).max() <= 1e-11 * amp
    assert np.abs(np.asarray(rod.velocity_collection)).max() <= 1e-11   # pure twist moves no node
    # the twist angle itself: element 0 has turned by (amp / omega_h) shape_0 sin(omega_h t) about y, up to the half-step offset
    Q = np.asarray(rod.director_collection)[:, :, 0]
    angle = np.arctan2(Q[0, 0], Q[0, 2])                  # d1 started along z and turns in the z-x plane
    assert abs(abs(angle) * omega_h / (amp * shape[0]) - abs(np.sin(omega_h * t))) <= 2e-5
    wrong = verlet_frequency(omega * np.sqrt(0.5), dt)    # I1 in place of the polar moment I3 = 2 I1
    assert abs(np.cos(wrong * t) - np.cos(omega_h * t)) > 0.05


def test_oracle_torsional_wave_frequency():
    check_torsional_wave(oracle_api(), device=False)


@pytest.mark.gpu
def test_device_torsional_wave_frequency():
    check_torsional_wave(product_api(), device=True)


# ---------------------------------------------------------------------------------------------------- damper
def check_damper_decay(api, device):
    n, dt, n_steps, nu = 12, 2.0e-5, 4000, 5.0e-4   # rotational rate nu A / I3 = 40 / s, translational 5e-4 / s
    rod = straight(api, n)
    v0 = np.array([0.02, -0.01, 0.03])
    spin = 5.0
    rod.velocity_collection = np.repeat(v0[:, None], n + 1, axis=1)
    w0 = np.zeros((3, n))
    w0[2] = spin
    rod.omega_collection = w0
    sim = api["sim"]()
    sim.append(rod)
    sim.dampen(rod).using(api["damper"], damping_constant=nu, time_step=dt)
    sim.finalize()
    t = advance(api, sim, dt, n_steps, device)
    v = np.asarray(rod.velocity_collection)
    w = np.asarray(rod.omega_collection)
    want_v = v0 * np.exp(-nu * t)
    # (rounding-level strains against E A / m make the velocities ring at a few 1e-12 m/s, as in the free-fall test)
    assert np.abs(v - want_v[:, None]).max() <= 1e-9 * np.abs(want_v).max()
    assert np.abs(v0 - want_v).max() >= 1e-6 * np.abs(v0).max()   # (the decay is 4e-5 of v0: 4e4 times the bound)
    rate = nu * 2.0 / (R * R)                              # nu m_e / J3 = nu A / I3
    want_w = spin * np.exp(-rate * t)
    assert 1e-3 < want_w / spin < 0.9                      # neither untouched nor gone
    assert np.abs(w[2] - want_w).max() <= 1e-9 * want_w
    assert np.abs(w[:2]).max() <= 1e-13
    # the rod has drifted by v0 (1 - exp(-nu t)) / nu up to the scheme's half-step bookkeeping, rigidly
    x = np.asarray(rod.position_collection)
    drift = x[:, 0] - np.array([0.0, 0.0, 0.0])
    assert np.abs(x - x[:, :1] - np.array([[0.0], [1.0], [0.0]]) * np.linspace(0.0, L, n + 1)).max() <= 1e-13
    exact = v0 * (1.0 - np.exp(-nu * t)) / nu
    assert np.abs(drift - exact).max() <= 2.0 * nu * dt * np.abs(exact).max()   # first-order in dt, as for any one-step scheme
    assert AREA > 0.0


def test_oracle_damper_decay_laws():
    check_damper_decay(oracle_api(), device=False)


@pytest.mark.gpu
def test_device_damper_decay_laws():
    check_damper_decay(product_api(), device=True)
