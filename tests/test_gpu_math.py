"""The fast kernel's fp64 helpers (csrc/br2_math.cuh) against libm / mpmath on the device, over the
validity ranges inside which the fast kernel uses them (outside, the exact-path kernel takes over).
Stated bound: 4 ulp (8.9e-16 relative) -- seven orders of magnitude inside the 1e-9 parity budget."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ULP4 = 4 * 2.0 ** -52


def run(which, x):
    import torch

    from br2 import _native

    lib = _native.load()
    xin = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).cuda()
    out = torch.empty_like(xin)
    _native.check(lib.br2_selftest_math(which, xin.data_ptr(), out.data_ptr(), xin.numel()))
    return out.cpu().numpy()


def relerr(got, want):
    return np.abs(got - want) / np.abs(want)


def test_rcp_and_rsqrt():
    rng = np.random.default_rng(0)
    x = np.concatenate([10.0 ** rng.uniform(-30, 30, 200000), rng.uniform(0.5, 2.0, 200000), [1e-300, 1e300, 1.0]])
    assert relerr(run(0, x), 1.0 / x).max() <= ULP4
    assert relerr(run(1, x), 1.0 / np.sqrt(x)).max() <= ULP4
    assert relerr(run(0, -x), -1.0 / x).max() <= ULP4


def test_sinc_cosc():
    import mpmath as mp

    rng = np.random.default_rng(1)
    t = np.concatenate([10.0 ** rng.uniform(-40, -2, 4000), rng.uniform(0.0, 0.01, 4000), [0.0, 0.01]])
    mp.mp.dps = 120
    sinc = np.array([float(mp.sin(mp.sqrt(v)) / mp.sqrt(v)) if v > 0 else 1.0 for v in t])
    cosc = np.array([float((1 - mp.cos(mp.sqrt(v))) / mp.mpf(v)) if v > 0 else 0.5 for v in t])
    assert relerr(run(2, t), sinc).max() <= ULP4
    assert relerr(run(3, t), cosc).max() <= ULP4


def test_logmap_scale():
    import mpmath as mp

    mp.mp.dps = 60
    rng = np.random.default_rng(2)
    theta = np.concatenate([10.0 ** rng.uniform(-5, -0.3, 4000), rng.uniform(0.0, 0.5, 2000)])
    c = np.cos(theta)
    c = c[(c < 1.0 - 1e-11) & (1.0 - c <= 0.125)]  # the rod kernel always subtracts 1e-10 first
    want = np.array([float(-mp.mpf(0.5) * mp.acos(mp.mpf(v)) / mp.sin(mp.acos(mp.mpf(v)) + mp.mpf("1e-14"))) for v in c])
    # compared as a function of c (the conditioning of acos near 1 belongs to the input, not the method)
    assert relerr(run(4, c), want).max() <= 8 * 2.0 ** -52


def test_exp_near_zero():
    rng = np.random.default_rng(3)
    d = np.concatenate([rng.uniform(-0.02, 0.02, 100000), 10.0 ** rng.uniform(-20, -2, 1000), [0.0, 0.02, -0.02]])
    assert relerr(run(5, d), np.exp(d)).max() <= ULP4


# ---- second generation (tile kernel): shorter series, tighter ranges ---------------------------------
def test_tile_sinc_cosc():
    import mpmath as mp

    rng = np.random.default_rng(4)
    t = np.concatenate([10.0 ** rng.uniform(-40, -3, 4000), rng.uniform(0.0, 2e-3, 4000), [0.0, 2e-3]])
    mp.mp.dps = 120
    sinc = np.array([float(mp.sin(mp.sqrt(v)) / mp.sqrt(v)) if v > 0 else 1.0 for v in t])
    cosc = np.array([float((1 - mp.cos(mp.sqrt(v))) / mp.mpf(v)) if v > 0 else 0.5 for v in t])
    assert relerr(run(6, t), sinc).max() <= ULP4
    assert relerr(run(7, t), cosc).max() <= ULP4


def test_tile_logmap_scale():
    import mpmath as mp

    mp.mp.dps = 60
    rng = np.random.default_rng(5)
    theta = np.concatenate([10.0 ** rng.uniform(-5, -0.7, 4000), rng.uniform(0.0, 0.245, 2000)])
    c = np.cos(theta)
    c = c[(c < 1.0 - 1e-11) & (1.0 - c <= 0.03)]
    want = np.array([float(-mp.mpf(0.5) * mp.acos(mp.mpf(v)) / mp.sin(mp.acos(mp.mpf(v)) + mp.mpf("1e-14"))) for v in c])
    assert relerr(run(8, c), want).max() <= 8 * 2.0 ** -52


def test_tile_exp_and_sqrt():
    rng = np.random.default_rng(6)
    d = np.concatenate([rng.uniform(-0.09, 0.09, 100000), 10.0 ** rng.uniform(-20, -2, 1000), [0.0, 0.09, -0.09]])
    assert relerr(run(9, d), np.exp(d)).max() <= ULP4
    x = np.concatenate([10.0 ** rng.uniform(-30, 2, 100000), [1.0, 4.0]])
    assert relerr(run(10, x), np.sqrt(x)).max() <= 2.0 ** -38
    assert run(10, np.zeros(4)).tolist() == [0.0] * 4
    # the radius' square root (cubic step on x * rsqrt-seed): full precision over anything a cross-section area can be
    y = np.concatenate([10.0 ** rng.uniform(-12, 4, 200000), [1.0, 4.0, 2.5e-5]])
    assert relerr(run(12, y), np.sqrt(y)).max() <= ULP4


def test_tile_rotation_angle_keeps_the_reference_guard():
    """One kinematic half step rotates by |u| * |u| / (|u| + 1e-14) in the reference (axis = u / (|u| + 1e-14));
    the tile kernel evaluates that guard from hardware-seed reciprocals."""
    import mpmath as mp

    mp.mp.dps = 50
    omega = np.concatenate([10.0 ** np.random.default_rng(7).uniform(-9, 3.3, 4000), [0.0]])
    u = [mp.mpf(float(w)) * mp.mpf("1e-5") for w in omega]
    want = np.array([float(v * v / (v + mp.mpf("1e-14"))) if v > 0 else 0.0 for v in u])
    got = run(11, omega)
    # the seeds (2^-20 relative) only enter through the guard term 1e-14 |u| / (|u| + 1e-14) <= 1e-14 rad: an absolute
    # error below 1e-19 rad per half step whatever the angle; the rest is rounding (atan2 of the test included)
    assert (np.abs(got - want) <= 8 * 2.0 ** -52 * want + 1e-19).all()
