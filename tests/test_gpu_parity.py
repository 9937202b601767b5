"""GPU parity: the CUDA path (through the br2 package and the C ABI) against the oracle.

Tolerances (BASELINE.json north_star): positions max|dx|/max|x| <= 1e-9 and directors max|dQ| <= 1e-9
after 10^4 steps, fp64.  Velocities / omegas sit on the rounding-noise floor of the dynamics
(tests/test_oracle_fused.py measures ~1e-9 / ~3e-8 between two CPU builds that differ in one ulp of
``pow``), so they get looser, stated bounds."""
import os

import numpy as np
import pytest

from conftest import DATA, GOLDEN
from golden.cases import CALLBACK_KEYS, TRAJECTORIES

pytestmark = pytest.mark.gpu

ROD_DB = os.path.join(DATA, "rod_library.json")
ASSEMBLY = {k: os.path.join(DATA, "assembly_%s.json" % k) for k in ("single", "double", "triple")}
PSI = 6895.0
NAMES = ["action1", "action2", "action3"]
POS_TOL, DIR_TOL, VEL_TOL, OMEGA_TOL = 1e-9, 1e-9, 1e-7, 1e-5

SHORT = {"position_collection": "x", "velocity_collection": "v", "director_collection": "Q",
         "omega_collection": "w", "acceleration_collection": "a"}


VARIANTS = {"generic": 0, "fast": 1, "fast-uniform": 2, "tile": 3}
TILE_ASSEMBLIES = ("single", "double", "triple")
AUTO_VARIANT = {"single": "tile", "double": "tile", "triple": "tile"}


def make_env(assembly, batch, fps=25, variant=None, **kwargs):
    from br2.environment import Environment

    env = Environment("gpu_test", rendering_fps=fps, batch=batch, create_dirs=False)
    env.reset(ROD_DB, ASSEMBLY[assembly], **kwargs)
    if variant is not None:
        env.engine.select_kernel(VARIANTS[variant])
        assert env.engine.kernel_info()["variant_name"] == variant
    return env


@pytest.mark.parametrize("case", TRAJECTORIES, ids=[c[0] for c in TRAJECTORIES])
def test_golden_trajectories(case):
    """Same assembly JSON, action, dt and duration as the golden runs of the reference's real br2 code."""
    name, assembly, kwargs, action, duration, fps = case
    g = np.load(os.path.join(GOLDEN, "traj_%s.npz" % name))
    env = make_env(assembly, 1, fps=fps, **kwargs)
    assert env.engine.kernel_info()["variant_name"] == AUTO_VARIANT[assembly]  # the sample assemblies take the fast paths
    status = env.run(action, duration=duration, disable_progress_bar=True, check_nan=True, check_steady_state=1)
    assert env.time == float(g["final_time"]) and env.step_skip == int(g["step_skip"])
    assert list(env.shearable_rods.keys()) == list(g["rod_names"])
    assert not (status.position_nan_status or status.director_nan_status)
    assert abs(status.max_velocity - float(g["max_velocity"])) <= POS_TOL * float(g["max_velocity"])
    x_scale = max(np.abs(g["rod%d_position_collection" % i]).max() for i in range(len(env.simulator)))
    for idx, rod in enumerate(env.shearable_rods.values()):
        assert np.abs(rod.position_collection - g["rod%d_position_collection" % idx]).max() <= POS_TOL * x_scale
        assert np.abs(rod.director_collection - g["rod%d_director_collection" % idx]).max() <= DIR_TOL
        assert np.abs(rod.velocity_collection - g["rod%d_velocity_collection" % idx]).max() <= VEL_TOL
        assert np.abs(rod.omega_collection - g["rod%d_omega_collection" % idx]).max() <= OMEGA_TOL
        # callback history: same cadence, keys and shapes as FreeCallback in the reference
        sink = env.data_rods[idx]
        np.testing.assert_array_equal(np.array(sink["step"]), g["rod%d_cb_step" % idx])
        np.testing.assert_array_equal(np.array(sink["time"]), g["rod%d_cb_time" % idx])
        for key in CALLBACK_KEYS:
            got, want = np.array(sink[key]), g["rod%d_cb_%s" % (idx, key)]
            assert got.shape == want.shape, key
        for key, tol in (("position", POS_TOL * x_scale), ("director", DIR_TOL), ("com", POS_TOL * x_scale),
                         ("lengths", 1e-12), ("radius", 1e-12), ("dilatation", 1e-9)):
            assert np.abs(np.array(sink[key]) - g["rod%d_cb_%s" % (idx, key)]).max() <= tol, key
        # loads and accelerations carry the joint's round-to-12-decimals noise (5.8e-9 N per flipped digit,
        # over node masses of ~1e-3 kg): stated bounds, relative to the field's magnitude
        for key, rtol in (("internal_forces", 1e-6), ("external_forces", 1e-6), ("internal_torques", 1e-6),
                          ("external_torques", 1e-6), ("acceleration", 1e-4)):
            want = g["rod%d_cb_%s" % (idx, key)]
            assert np.abs(np.array(sink[key]) - want).max() <= rtol * max(np.abs(want).max(), 1.0), key
    env.close()


def fused_reference(assembly, kwargs, pressures, n_steps, dt=2.0e-5):
    from oracle import br2_port
    from oracle.fused import FusedOracle

    ref = br2_port.OracleEnvironment(time_step=dt)
    ref.reset(ROD_DB, ASSEMBLY[assembly], **kwargs)
    fused = FusedOracle.from_simulator(ref.simulator, NAMES)
    W = fused.pack(pressures.shape[1])
    fused.run(W, pressures, n_steps, 0.0, dt)
    return fused, W


BATCH_CASES = [
    # BASELINE.json configs 2 / 3 (and the sample triple assembly at n=41), sampled: 64 random instances
    ("single", {"gravity": True}, 0, 10000),
    ("double", {}, 1, 10000),
    ("triple", {"gravity": True}, 2, 10000),
]


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("assembly,kwargs,seed,n_steps", BATCH_CASES, ids=[c[0] for c in BATCH_CASES])
def test_batched_random_pressures_match_oracle(assembly, kwargs, seed, n_steps, variant):
    if variant == "tile" and assembly not in TILE_ASSEMBLIES:
        pytest.skip("assembly has tip-to-base joints: served by the one-slot-per-thread fast kernel")
    batch, dt = 64, 2.0e-5
    rng = np.random.default_rng(seed)
    pressures = rng.uniform(0.0, 40.0, size=(3, batch)) * PSI
    env = make_env(assembly, batch, variant=variant, **kwargs)
    env.simulator._callback_ops = []
    action = {name: pressures[i] for i, name in enumerate(NAMES)}
    status = env.run(action, duration=n_steps * dt - 0.5 * dt, disable_progress_bar=True, check_nan=True)
    assert round(env.time / dt) == n_steps
    assert not any(status.nan_per_instance[k].any() for k in status.nan_per_instance)
    # every variant's OWN result: no instance went through the exact-path replay (the one-slot-per-thread kernels used to
    # flag every instance of the multi-segment assemblies -- their strain waves left the first damper series' range)
    assert env.engine.replay_count() == 0 and env.engine.unresolved_count() == 0
    fused, W = fused_reference(assembly, kwargs, pressures, n_steps, dt)
    x_scale = max(np.abs(fused.view(W, r, "x")).max() for r in range(len(env.simulator)))
    worst = {"x": 0.0, "Q": 0.0, "v": 0.0, "w": 0.0}
    for r, rod in enumerate(env.simulator):
        for attr, short in SHORT.items():
            if short == "a":
                continue
            err = np.abs(getattr(rod, attr) - fused.view(W, r, short)).max()
            worst[short] = max(worst[short], err / x_scale if short == "x" else err)
    print(assembly, variant, "worst errors", worst)
    assert worst["x"] <= POS_TOL and worst["Q"] <= DIR_TOL
    assert worst["v"] <= VEL_TOL and worst["w"] <= OMEGA_TOL
    # instances really are independent: instance 5 alone reproduces its batch result bit for bit
    solo = make_env(assembly, 1, variant=variant, **kwargs)
    solo.simulator._callback_ops = []
    solo.run({name: pressures[i, 5] for i, name in enumerate(NAMES)}, duration=n_steps * dt - 0.5 * dt,
             disable_progress_bar=True)
    for r, rod in enumerate(env.simulator):
        np.testing.assert_array_equal(solo.simulator[r].position_collection, rod.position_collection[5])
        np.testing.assert_array_equal(solo.simulator[r].director_collection, rod.director_collection[5])
    # directors stay orthonormal
    for rod in env.simulator:
        Q = rod.director_collection
        gram = np.einsum("bijk,bljk->bilk", Q, Q)
        assert np.abs(gram - np.eye(3)[None, :, :, None]).max() < 1e-10
    env.close()
    solo.close()


def _cluster_case(assembly, rod_db, kwargs, seed, n_steps, dt, batch, monkeypatch=None, one_cta_threads=None):
    """Random-pressure batch on the tile kernel vs the C oracle; returns the worst errors."""
    from br2.environment import Environment
    from oracle import br2_port
    from oracle.fused import FusedOracle

    if one_cta_threads is not None:
        monkeypatch.setenv("BR2_TILE_ONE_CTA_THREADS", str(one_cta_threads))
    rng = np.random.default_rng(seed)
    pressures = rng.uniform(0.0, 40.0, size=(3, batch)) * PSI
    env = Environment("gpu_test", time_step=dt, batch=batch, create_dirs=False)
    env.reset(rod_db, ASSEMBLY[assembly], **kwargs)
    env.simulator._callback_ops = []
    info = env.engine.kernel_info()
    assert info["variant_name"] == "tile"
    status = env.run({name: pressures[i] for i, name in enumerate(NAMES)}, duration=n_steps * dt - 0.5 * dt,
                     disable_progress_bar=True, check_nan=True)
    assert round(env.time / dt) == n_steps
    assert not any(status.nan_per_instance[k].any() for k in status.nan_per_instance)
    assert env.engine.replay_count() == 0 and env.engine.unresolved_count() == 0
    ref = br2_port.OracleEnvironment(time_step=dt)
    ref.reset(rod_db, ASSEMBLY[assembly], **kwargs)
    fused = FusedOracle.from_simulator(ref.simulator, NAMES)
    W = fused.pack(batch)
    fused.run(W, pressures, n_steps, 0.0, dt)
    x_scale = max(np.abs(fused.view(W, r, "x")).max() for r in range(len(env.simulator)))
    worst = {"x": 0.0, "Q": 0.0, "v": 0.0, "w": 0.0}
    for r, rod in enumerate(env.simulator):
        for attr, short in SHORT.items():
            if short == "a":
                continue
            err = np.abs(getattr(rod, attr) - fused.view(W, r, short)).max()
            worst[short] = max(worst[short], err / x_scale if short == "x" else err)
    env.close()
    return worst, info


def test_cluster_path_on_the_sample_triple(monkeypatch):
    """The sample triple assembly forced onto the cluster path (three CTAs of 63 threads, one per segment, tip-to-base
    joints through distributed shared memory) gives the same answer as on one CTA."""
    worst, info = _cluster_case("triple", ROD_DB, {"gravity": True}, 2, 2000, 2.0e-5, 16, monkeypatch, one_cta_threads=100)
    print("cluster triple n=41", info, worst)
    assert info["threads"] == 192
    assert worst["x"] <= POS_TOL and worst["Q"] <= DIR_TOL and worst["v"] <= VEL_TOL and worst["w"] <= OMEGA_TOL


def test_cluster_path_reports_the_same_callback_arrays(monkeypatch):
    """Everything FreeCallback records (state AND the derived arrays written on a launch's last step: accelerations,
    internal / external loads, lengths, dilatation, radius) agrees between the one-CTA and the cluster layout."""
    from br2.environment import Environment

    action = {"action1": np.array([15.0, 33.0]) * PSI, "action2": 30 * PSI, "action3": np.array([5.0, 0.0]) * PSI}
    histories = []
    for threads in (None, "100"):
        if threads is not None:
            monkeypatch.setenv("BR2_TILE_ONE_CTA_THREADS", threads)
        env = Environment("gpu_test", rendering_fps=250, batch=2, create_dirs=False)  # step_skip = 200
        env.reset(ROD_DB, ASSEMBLY["triple"], gravity=True)
        assert env.engine.kernel_info()["threads"] == 192
        env.run(action, duration=600 * 2e-5 - 1e-5, disable_progress_bar=True)
        histories.append([{key: np.array(sink[key]) for key in sink} for sink in env.data_rods])
        env.close()
    one_cta, cluster = histories
    for sink_a, sink_b in zip(one_cta, cluster):
        assert list(sink_a["step"]) == [200, 400, 600]
        for key in sink_a:
            scale = max(np.abs(sink_a[key]).max(), 1e-300)
            assert np.abs(sink_a[key] - sink_b[key]).max() <= 1e-9 * scale, key


def test_cluster_path_replays_out_of_range_instances_exactly(monkeypatch):
    """Cluster-sized assemblies have no one-CTA generic kernel to fall back on: the exact-arithmetic instantiation of the
    cluster kernel redoes flagged instances from the chunk in which they left the series' range.  Forced here on the
    sample triple assembly and compared with the generic kernel."""
    dt, n_steps = 2.0e-5, 400
    pressures = np.array([[30.0, 3.0e4, 10.0, 5.0e4], [20.0, 2.0e4, 5.0, 5.0e4], [10.0, 1.0e4, 0.0, 5.0e4]]) * PSI
    action = {name: pressures[i] for i, name in enumerate(NAMES)}
    generic = make_env("triple", 4, variant="generic", gravity=True)
    generic.simulator._callback_ops = []
    generic.run(action, duration=n_steps * dt - 0.5 * dt, disable_progress_bar=True)
    monkeypatch.setenv("BR2_TILE_ONE_CTA_THREADS", "100")
    cluster = make_env("triple", 4, gravity=True)
    cluster.simulator._callback_ops = []
    assert cluster.engine.kernel_info()["variant_name"] == "tile" and cluster.engine.kernel_info()["threads"] == 192
    cluster.run(action, duration=n_steps * dt - 0.5 * dt, disable_progress_bar=True)
    assert cluster.engine.replay_count() >= 2 and cluster.engine.unresolved_count() == 0
    assert sum(cluster.engine.replay_reasons().values()) >= 2
    for rc, rg in zip(cluster.simulator, generic.simulator):
        xc, xg = rc.position_collection, rg.position_collection
        assert np.array_equal(np.isfinite(xc), np.isfinite(xg))
        ok = np.isfinite(xg)
        assert np.abs(xc[ok] - xg[ok]).max() <= 1e-6 * max(1.0, np.abs(xg[ok]).max())
        for inst in (0, 2):
            assert np.abs(xc[inst] - xg[inst]).max() <= POS_TOL * 0.6
    cluster.close()
    generic.close()


def test_high_resolution_triple_baseline_config_4():
    """BASELINE config 4 at the parity sample SURVEY.md section 8(d) asks for: triple assembly with 200 elements per rod
    (1809 node slots) at dt = 5e-6 (dt = 2e-5 diverges at this resolution), instance 0 plus 64
    random instances, 10^4 steps -- a cluster of three 320-thread CTAs per instance."""
    worst, info = _cluster_case("triple", os.path.join(DATA, "rod_library_n200.json"), {"gravity": True}, 2, 10000, 5.0e-6, 65)
    print("cluster triple n=200", info, worst)
    assert info["threads"] == 320
    assert worst["x"] <= POS_TOL and worst["Q"] <= DIR_TOL and worst["v"] <= VEL_TOL and worst["w"] <= OMEGA_TOL


@pytest.mark.parametrize("variant", ["generic", "fast-uniform", "tile"])
def test_chunked_launches_equal_one_launch(variant):
    """Splitting a run into launches (as capture steps do) must not change the result: bit for bit with the
    generic kernel; to rounding noise with the fast kernel, which fuses the trailing and leading kinematic
    half steps INSIDE a launch into one rotation and so rounds differently at launch boundaries."""
    dt = 2.0e-5
    action = {"action1": 30 * PSI, "action2": 12 * PSI, "action3": 3 * PSI}
    a = make_env("single", 2, variant=variant, gravity=True)
    a.simulator._callback_ops = []
    a.run(action, duration=600 * dt - 0.5 * dt, disable_progress_bar=True)
    b = make_env("single", 2, variant=variant, gravity=True)
    b.simulator._callback_ops = []
    for _ in range(6):
        b.run(action, duration=100 * dt - 0.5 * dt, disable_progress_bar=True)
    assert a.time == b.time
    for ra, rb in zip(a.simulator, b.simulator):
        for attr in ("position_collection", "director_collection"):
            if variant == "generic":
                np.testing.assert_array_equal(getattr(ra, attr), getattr(rb, attr), err_msg=attr)
            else:
                # rounding-noise floor of the dynamics (~5e-13 / ~4e-12), two decades inside the parity budget
                assert np.abs(getattr(ra, attr) - getattr(rb, attr)).max() <= 1e-11, attr
    a.close()
    b.close()


def test_optional_ops_point_force_and_last_end_fixed():
    """PointForces / LastEndFixedRod on a lone rod, through the hooks, vs the faithful oracle."""
    from br2.custom_activation import PointForces
    from br2.custom_constraint import LastEndFixedRod
    from br2.free_simulator import BR2Simulator
    from br2.hooks import AnalyticalLinearDamper, CosseratRod, PositionVerlet, extend_stepper_interface
    from oracle import br2_port, mini_elastica

    spec = dict(n_elements=20, start=np.zeros(3), direction=np.array([0.0, 1.0, 0.0]), normal=np.array([0.0, 0.0, 1.0]),
                base_length=0.2, base_radius=0.005, density=1200.0, youngs_modulus=2e6, shear_modulus=1.2e6)
    dt, n_steps, force = 1.0e-5, 3000, np.array([0.02, 0.0, -0.01])

    def build(sim_cls, rod_cls, damper, point, pin, stepper_mod):
        sim = sim_cls()
        rod = rod_cls.straight_rod(**spec)
        sim.append(rod)
        sim.dampen(rod).using(damper, damping_constant=2.0, time_step=dt)
        sim.add_forcing_to(rod).using(point, force, 0.4, ramp_up_time=5e-3)
        sim.constrain(rod).using(pin, constrained_position_idx=(-1,), constrained_director_idx=(-1,))
        sim.finalize()
        return sim, rod

    sim, rod = build(BR2Simulator, CosseratRod, AnalyticalLinearDamper, PointForces, LastEndFixedRod, None)
    assert sim.engine.kernel_info()["variant_name"] == "tile"  # point forces ride the tile kernel as extra forces
    assert sim.kernel_plan() == "tile: 1 CTA(s) per instance of 11 threads (2 slots per thread), 1 tip-to-base joint(s) / point force(s)"
    do_step, stages = extend_stepper_interface(PositionVerlet(), sim)
    t = do_step(PositionVerlet(), stages, sim, 0.0, dt, n_steps=n_steps)
    osim, orod = build(br2_port.BR2Simulator, mini_elastica.CosseratRod, mini_elastica.AnalyticalLinearDamper,
                       br2_port.PointForces, br2_port.LastEndFixedRod, None)
    ot = mini_elastica.timestepper.integrate(mini_elastica.PositionVerlet(), osim, n_steps * dt, n_steps)
    assert abs(t - ot) < 1e-15
    assert np.abs(orod.position_collection[:, 8] - spec["start"][:, None][:, 0]).max() > 1e-6  # it moved
    scale = np.abs(orod.position_collection).max()
    assert np.abs(rod.position_collection - orod.position_collection).max() <= POS_TOL * scale
    assert np.abs(rod.director_collection - orod.director_collection).max() <= DIR_TOL
    np.testing.assert_array_equal(rod.position_collection[:, -1], orod.position_collection[:, -1])
    np.testing.assert_array_equal(rod.velocity_collection[:, -1], 0.0)


def test_baseline_config_1_examples_run_script(tmp_path, monkeypatch):
    """BASELINE configs[0]: what /root/reference/examples/run.py does -- single_br2_v1, gravity, 30 psi on action1 and
    action2, duration 1.0 s (50001 steps from the fp64 time accumulation, 25 captures), check_nan and steady-state check,
    save_data -- against the C oracle after all 50001 steps."""
    from br2.environment import Environment

    monkeypatch.chdir(tmp_path)
    action = {"action1": 30 * PSI, "action2": 30 * PSI}
    env = Environment(run_tag="c1")
    env.reset(rod_database_path=ROD_DB, assembly_config_path=ASSEMBLY["single"], gravity=True)
    status = env.run(action=action, duration=1.0, check_nan=True, check_steady_state=True, disable_progress_bar=True)
    assert status.end_status and not status.position_nan_status and not status.director_nan_status
    n_steps = 50001
    assert round(env.time / env.time_step) == n_steps and env.step_skip == 2000
    assert list(env.data_rods[0]["step"]) == list(range(2000, 50001, 2000))
    assert env.engine.replay_count() == 0
    env.save_data()
    saved = np.load(os.path.join("result_c1", "data", "br2_data.npz"))
    assert saved["positions"].shape == (3, 25, 3, 41) and saved["directors"].shape == (3, 25, 3, 3, 41)
    pressures = np.array([[30 * PSI], [30 * PSI], [0.0]])
    fused, W = fused_reference("single", {"gravity": True}, pressures, n_steps)
    x_scale = max(np.abs(fused.view(W, r, "x")).max() for r in range(3))
    for r, rod in enumerate(env.simulator):
        # 5e4 steps: the rounding-noise floor has had five times longer to grow than in the 1e4-step bar
        assert np.abs(rod.position_collection - fused.view(W, r, "x")[0]).max() <= 5 * POS_TOL * x_scale
        assert np.abs(rod.director_collection - fused.view(W, r, "Q")[0]).max() <= 5 * DIR_TOL
    tip = env.simulator[0].position_collection[:, -1]
    assert abs(tip[0]) > 1e-3  # the arm really bent
    env.close()


def test_lone_rod_with_odd_slot_count_on_the_tile_kernel():
    """Ragged case for the tile kernel: one rod, no joints (no partner / owner rows), 21 node slots = 10 full runs of two
    slots and one half-empty run; gravity, damper, soft-fixed base and a twist actuation through the hooks, vs the
    faithful oracle."""
    from br2.free_custom_systems import FreeBaseEndSoftFixed, FreeTwistActuation
    from br2.free_simulator import BR2Simulator
    from br2.hooks import AnalyticalLinearDamper, CosseratRod, GravityForces, PositionVerlet, extend_stepper_interface
    from oracle import br2_port, mini_elastica

    spec = dict(n_elements=20, start=np.zeros(3), direction=np.array([0.0, 1.0, 0.0]), normal=np.array([0.0, 0.0, 1.0]),
                base_length=0.2, base_radius=0.005, density=1200.0, youngs_modulus=2e6, shear_modulus=1.2e6)
    dt, n_steps = 1.0e-5, 3000

    def build(sim_cls, rod_cls, damper, gravity, fixed, twist):
        sim = sim_cls()
        rod = rod_cls.straight_rod(**spec)
        sim.append(rod)
        sim.dampen(rod).using(damper, damping_constant=2.0, time_step=dt)
        sim.add_forcing_to(rod).using(gravity, acc_gravity=np.array([0.3, 0.0, -9.80665]))
        pressure = [2.0e4]
        sim.add_forcing_to(rod).using(twist, actuation_ref=pressure, scale=1.25797e-6, ramp_up_time=5e-3)
        sim.constrain(rod).using(fixed, constrained_position_idx=(0,), constrained_director_idx=(0,), k=1e9, nu=0.0, kt=0.0)
        sim.finalize()
        return sim, rod

    sim, rod = build(BR2Simulator, CosseratRod, AnalyticalLinearDamper, GravityForces, FreeBaseEndSoftFixed, FreeTwistActuation)
    assert sim.engine.kernel_info()["variant_name"] == "tile"
    do_step, stages = extend_stepper_interface(PositionVerlet(), sim)
    do_step(PositionVerlet(), stages, sim, 0.0, dt, n_steps=n_steps)
    assert sim.engine.replay_count() == 0
    osim, orod = build(br2_port.BR2Simulator, mini_elastica.CosseratRod, mini_elastica.AnalyticalLinearDamper,
                       mini_elastica.GravityForces, br2_port.FreeBaseEndSoftFixed, br2_port.FreeTwistActuation)
    mini_elastica.timestepper.integrate(mini_elastica.PositionVerlet(), osim, n_steps * dt, n_steps)
    assert np.abs(orod.position_collection[:, -1] - np.array([0.0, 0.2, 0.0])).max() > 1e-5  # it moved
    scale = np.abs(orod.position_collection).max()
    assert np.abs(rod.position_collection - orod.position_collection).max() <= POS_TOL * scale
    assert np.abs(rod.director_collection - orod.director_collection).max() <= DIR_TOL
    assert np.abs(rod.velocity_collection - orod.velocity_collection).max() <= VEL_TOL
    assert np.abs(rod.omega_collection - orod.omega_collection).max() <= OMEGA_TOL


def test_last_end_fixed_rod_on_the_tile_kernel():
    """custom_constraint.LastEndFixedRod (/root/reference/br2/custom_constraint.py:35-88) pins the last node and element;
    without point forces the rod stays on the tile kernel.  A cantilever hanging from its far end under gravity."""
    from br2.custom_constraint import LastEndFixedRod
    from br2.free_simulator import BR2Simulator
    from br2.hooks import AnalyticalLinearDamper, CosseratRod, GravityForces, PositionVerlet, extend_stepper_interface
    from oracle import br2_port, mini_elastica

    spec = dict(n_elements=31, start=np.zeros(3), direction=np.array([0.0, 1.0, 0.0]), normal=np.array([0.0, 0.0, 1.0]),
                base_length=0.25, base_radius=0.004, density=1100.0, youngs_modulus=3e6, shear_modulus=2e6)
    dt, n_steps = 1.0e-5, 4000

    def build(sim_cls, rod_cls, damper, gravity, pin):
        sim = sim_cls()
        rod = rod_cls.straight_rod(**spec)
        sim.append(rod)
        sim.dampen(rod).using(damper, damping_constant=1.0, time_step=dt)
        sim.add_forcing_to(rod).using(gravity, acc_gravity=np.array([0.0, 0.0, -9.80665]))
        sim.constrain(rod).using(pin, constrained_position_idx=(-1,), constrained_director_idx=(-1,))
        sim.finalize()
        return sim, rod

    sim, rod = build(BR2Simulator, CosseratRod, AnalyticalLinearDamper, GravityForces, LastEndFixedRod)
    assert sim.engine.kernel_info()["variant_name"] == "tile"
    do_step, stages = extend_stepper_interface(PositionVerlet(), sim)
    do_step(PositionVerlet(), stages, sim, 0.0, dt, n_steps=n_steps)
    assert sim.engine.replay_count() == 0
    osim, orod = build(br2_port.BR2Simulator, mini_elastica.CosseratRod, mini_elastica.AnalyticalLinearDamper,
                       mini_elastica.GravityForces, br2_port.LastEndFixedRod)
    mini_elastica.timestepper.integrate(mini_elastica.PositionVerlet(), osim, n_steps * dt, n_steps)
    assert abs(orod.position_collection[2, 0]) > 1e-4  # the free end sagged
    scale = np.abs(orod.position_collection).max()
    assert np.abs(rod.position_collection - orod.position_collection).max() <= POS_TOL * scale
    assert np.abs(rod.director_collection - orod.director_collection).max() <= DIR_TOL
    np.testing.assert_array_equal(rod.position_collection[:, -1], orod.position_collection[:, -1])
    np.testing.assert_array_equal(rod.velocity_collection[:, -1], 0.0)
    np.testing.assert_array_equal(rod.omega_collection[:, -1], 0.0)


FULL_SIZE = [
    # BASELINE.json configs at their full batch sizes: (assembly, kwargs, batch, steps)
    ("single", {"gravity": True}, 4096, 2000),    # configs[1]
    ("double", {}, 16384, 500),                   # configs[2]
    ("single", {"gravity": True}, 65536, 250),    # configs[4], one GPU's worth is the whole sweep here
]


@pytest.mark.parametrize("assembly,kwargs,batch,n_steps", FULL_SIZE, ids=["single-4096", "double-16384", "single-65536"])
def test_full_size_batches_by_properties(assembly, kwargs, batch, n_steps):
    """At the BASELINE batch sizes the oracle is too slow to redo everything, so: (1) instances are independent and
    position-blind -- the second half of the batch repeats the first half's pressures and must reproduce it bit for bit;
    (2) a sample of instances spread over the batch matches the C oracle; (3) directors stay orthonormal; (4) nothing
    left the optimistic path, nothing is NaN."""
    dt = 2.0e-5
    rng = np.random.default_rng(batch)
    half = batch // 2
    first = rng.uniform(0.0, 40.0, size=(3, half)) * PSI
    if batch == 65536:
        first[0] = np.linspace(0.0, 40.0, half) * PSI       # configs[4]: action1 swept linearly over the instances
    pressures = np.concatenate([first, first], axis=1)
    env = make_env(assembly, batch, **kwargs)
    env.simulator._callback_ops = []
    assert env.engine.kernel_info()["variant_name"] == "tile"
    status = env.run({name: pressures[i] for i, name in enumerate(NAMES)}, duration=n_steps * dt - 0.5 * dt,
                     disable_progress_bar=True, check_nan=True)
    assert not any(status.nan_per_instance[k].any() for k in status.nan_per_instance)
    assert env.engine.replay_count() == 0
    torch = env.engine.torch
    state = env.engine.state
    assert torch.equal(state[:half], state[half:])
    sample = np.unique(np.concatenate([[0, 1, half - 1, half, batch - 1], rng.integers(0, batch, 11)]))
    fused, W = fused_reference(assembly, kwargs, pressures[:, sample], n_steps, dt)
    x_scale = max(np.abs(fused.view(W, r, "x")).max() for r in range(len(env.simulator)))
    idx = torch.as_tensor(sample, device=state.device)
    for r in range(len(env.simulator)):
        got_x = env.engine.view(r, "x")[idx].cpu().numpy()
        got_q = env.engine.view(r, "Q")[idx].cpu().numpy()
        assert np.abs(got_x - fused.view(W, r, "x")).max() <= POS_TOL * x_scale
        assert np.abs(got_q - fused.view(W, r, "Q")).max() <= DIR_TOL
        Q = env.engine.view(r, "Q")
        gram = torch.einsum("bijk,bljk->bilk", Q, Q)
        assert float((gram - torch.eye(3, dtype=Q.dtype, device=Q.device)[None, :, :, None]).abs().max()) < 1e-10
    env.close()


def test_device_side_captures_equal_host_snapshots():
    """FreeCallback histories recorded as device-side clones of the batch state (materialised on read) are identical to
    the ones recorded through the reference's make_callback protocol on host snapshots."""
    from br2.environment import Environment
    from br2.free_simulator import FreeCallback

    class EagerCallback(FreeCallback):
        make_callback_lazy = None  # forces the host-snapshot path

    action = {"action1": np.array([30.0, 5.0, 12.0]) * PSI, "action2": 12 * PSI, "action3": np.array([0.0, 20.0, 3.0]) * PSI}
    histories = []
    for callback in (None, EagerCallback):
        env = Environment("gpu_test", rendering_fps=250, batch=3, create_dirs=False)  # step_skip = 200
        env.reset(ROD_DB, ASSEMBLY["single"], gravity=True)
        if callback is not None:
            env.simulator._callback_ops = [(r, callback(op.every, op.callback_params, op.time_interval))
                                           for r, op in env.simulator._callback_ops]
        env.run(action, duration=1000 * 2e-5 - 1e-5, disable_progress_bar=True)
        if callback is None:
            series = env.data_rods[0]["position"]
            assert "still on the device" in repr(series) and len(series) == 5
        histories.append(env.data_rods)
        env.close()
    lazy, eager = histories
    for sink_l, sink_e in zip(lazy, eager):
        assert list(sink_l["step"]) == list(sink_e["step"]) == [200, 400, 600, 800, 1000]
        assert set(sink_l.keys()) == set(sink_e.keys())
        for key in sink_e:
            np.testing.assert_array_equal(np.array(sink_l[key]), np.array(sink_e[key]), err_msg=key)
        assert np.array(sink_l["position"]).shape == (5, 3, 3, 42)
        np.testing.assert_array_equal(sink_l["director"][-1], np.array(sink_e["director"])[-1])


def test_checkpoint_and_resume_reproduce_the_uninterrupted_run(tmp_path, monkeypatch):
    """save_state / load_state (reference free_simulator.py:93-99: one system_{idx}.npz per rod): a batch stopped after
    500 steps, saved, loaded into a fresh Environment and continued gives the uninterrupted run bit for bit."""
    from br2.environment import Environment

    monkeypatch.chdir(tmp_path)
    dt = 2.0e-5
    pressures = np.random.default_rng(11).uniform(0.0, 40.0, size=(3, 5)) * PSI
    action = {name: pressures[i] for i, name in enumerate(NAMES)}

    def fresh():
        env = Environment("ckpt", batch=5, create_dirs=False)
        env.reset(ROD_DB, ASSEMBLY["double"])
        env.simulator._callback_ops = []
        return env

    straight = fresh()
    straight.run(action, duration=500 * dt - 0.5 * dt, disable_progress_bar=True)
    straight.save_state(directory=str(tmp_path / "state"))
    saved_time = straight.time
    assert sorted(os.listdir(tmp_path / "state")) == ["system_%d.npz" % i for i in range(6)]
    straight.run(action, duration=700 * dt - 0.5 * dt, disable_progress_bar=True)

    resumed = fresh()
    resumed.load_state(directory=str(tmp_path / "state"))
    resumed.time = saved_time          # like the reference, load_state leaves the clock to the caller
    resumed.run(action, duration=700 * dt - 0.5 * dt, disable_progress_bar=True)
    assert resumed.time == straight.time
    for a, b in zip(straight.simulator, resumed.simulator):
        for attr in ("position_collection", "velocity_collection", "director_collection", "omega_collection"):
            np.testing.assert_array_equal(getattr(a, attr), getattr(b, attr), err_msg=attr)
    straight.close()
    resumed.close()


def test_step_host_entry_point_matches_device_path():
    """br2_step_host (host buffers in, host buffers out) == bind + step on device buffers."""
    import ctypes

    from br2 import _native

    env = make_env("single", 3, gravity=True)
    env.simulator._callback_ops = []
    pressures = np.random.default_rng(3).uniform(0, 40, size=(3, 3)) * PSI
    host_state = env.engine.state.cpu().numpy().copy()
    env.run({n: pressures[i] for i, n in enumerate(NAMES)}, duration=50 * 2e-5 - 1e-5, disable_progress_bar=True)
    t_end = ctypes.c_double()
    _native.check(env.engine.lib.br2_step_host(env.engine.handle, host_state.ctypes.data,
                                               np.ascontiguousarray(pressures).ctypes.data, 3, 50, 0.0, 2e-5,
                                               ctypes.byref(t_end)))
    assert t_end.value == env.time
    np.testing.assert_array_equal(host_state, env.engine.state.cpu().numpy())
    env.close()


@pytest.mark.parametrize("variant", ["fast-uniform", "tile"])
def test_fast_kernel_falls_back_to_exact_path_when_out_of_range(variant):
    """Absurd actuation drives strains / rotations outside the optimistic kernels' series range: the flagged
    instances must be redone by the exact-path kernel while the well-behaved instances of the same batch keep
    the optimistic kernel's result.  The one-slot-per-thread kernel replays the whole launch (result = the
    generic kernel's, bit for bit); the tile kernel replays from the start of the launch chunk in which the
    range was left (result = the generic kernel's up to the rounding noise of the chunks before)."""
    dt, n_steps = 2.0e-5, 400
    pressures = np.array([[30.0, 3.0e4, 10.0, 5.0e4], [20.0, 2.0e4, 5.0, 5.0e4], [10.0, 1.0e4, 0.0, 5.0e4]]) * PSI
    action = {name: pressures[i] for i, name in enumerate(NAMES)}
    fast = make_env("single", 4, variant=variant, gravity=True)
    fast.simulator._callback_ops = []
    fast.run(action, duration=n_steps * dt - 0.5 * dt, disable_progress_bar=True)
    replays = fast.engine.replay_count()
    assert replays >= 2, "instances 1 and 3 should have left the fast kernel's range"
    generic = make_env("single", 4, variant="generic", gravity=True)
    generic.simulator._callback_ops = []
    generic.run(action, duration=n_steps * dt - 0.5 * dt, disable_progress_bar=True)
    assert generic.engine.replay_count() == 0
    for rf, rg in zip(fast.simulator, generic.simulator):
        xf, xg = rf.position_collection, rg.position_collection
        for inst in (1, 3):
            if variant == "fast-uniform":  # replayed from the launch start by the same generic kernel (NaNs included)
                np.testing.assert_array_equal(xf[inst], xg[inst])
            else:
                assert np.array_equal(np.isfinite(xf[inst]), np.isfinite(xg[inst]))
                ok = np.isfinite(xg[inst])
                assert np.abs(xf[inst][ok] - xg[inst][ok]).max() <= 1e-6 * max(1.0, np.abs(xg[inst][ok]).max())
        for inst in (0, 2):  # stayed on the fast kernel: equal to rounding noise only
            assert np.abs(xf[inst] - xg[inst]).max() <= POS_TOL * 0.2
            assert np.isfinite(xf[inst]).all()
    fast.close()
    generic.close()


def test_benchmark_workload_never_leaves_the_fast_path():
    env = make_env("single", 256, gravity=True)
    env.simulator._callback_ops = []
    rng = np.random.default_rng(0)
    action = {name: rng.uniform(0.0, 40.0, size=256) * PSI for name in NAMES}
    env.run(action, duration=0.1, disable_progress_bar=True)
    assert env.engine.kernel_info()["variant_name"] == "tile"
    assert env.engine.replay_count() == 0
    env.close()


@pytest.mark.gpu
def test_cuda_path_reproduces_the_surveys_independent_probe_anchors():
    """RodBend tip after 5000 / 10000 steps of single_br2 (gravity, 30 psi on action1 and action2),
    from the survey's own mimic of the reference -- independent of this repository's oracle (see tests/test_oracle_fused.py)."""
    from br2.environment import Environment

    env = Environment(run_tag="anchors", create_dirs=False)
    env.reset(rod_database_path=ROD_DB, assembly_config_path=ASSEMBLY["single"], gravity=True)
    env.simulator._callback_ops = []
    action = {"action1": 30 * PSI, "action2": 30 * PSI}
    for n_steps, tip in ((5000, (-0.00470, 0.17958, 0.00057)), (10000, (-0.01113, 0.17870, 0.00216))):
        env.run(action=action, duration=5000 * env.time_step - 0.5 * env.time_step, disable_progress_bar=True)
        assert round(env.time / env.time_step) == n_steps
        got = env.simulator[0].position_collection[:, -1]
        assert np.abs(got - np.array(tip)).max() < 6e-6, (n_steps, got)
    env.close()


def _ragged_inputs(tmp_path):
    """A double assembly whose second segment uses shorter rods with fewer elements (30 on 0.12 m against 41 on 0.18 m):
    SURVEY.md section 8(f) row 4, per-rod heterogeneous n_elements across segments."""
    import json

    library = json.load(open(ROD_DB))
    for key in ("RodBend", "RodRightTwist", "RodLeftTwist"):
        library["Rods"][key + "Short"] = dict(library["Rods"][key], n_elements=30, base_length=0.12)
    assembly = json.load(open(ASSEMBLY["double"]))
    assembly["Segments"]["seg2"]["rod_order"] = ["RodRightTwistShort", "RodBendShort", "RodLeftTwistShort"]
    rod_db, path = tmp_path / "rod_library_ragged.json", tmp_path / "assembly_ragged.json"
    rod_db.write_text(json.dumps(library))
    path.write_text(json.dumps(assembly))
    return str(rod_db), str(path)


@pytest.mark.parametrize("variant", ["generic", "fast", "tile", "tile-as-cluster"])
def test_ragged_assembly_with_an_odd_batch_matches_oracle(tmp_path, variant, monkeypatch):
    from br2.environment import Environment

    if variant == "tile-as-cluster":  # one CTA per segment (63 and 48 threads), class tables in every CTA of the cluster
        monkeypatch.setenv("BR2_TILE_ONE_CTA_THREADS", "100")
        variant = "tile"
    from oracle import br2_port
    from oracle.fused import FusedOracle

    rod_db, assembly = _ragged_inputs(tmp_path)
    batch, n_steps, dt = 37, 3000, 2.0e-5
    pressures = np.random.default_rng(11).uniform(0.0, 40.0, size=(3, batch)) * PSI
    env = Environment("ragged", batch=batch, create_dirs=False)
    env.reset(rod_db, assembly, gravity=True)
    assert [rod.n_elems for rod in env.simulator] == [41, 41, 41, 30, 30, 30]
    # segments of 41 and 30 elements: two classes of rod constants; the tile kernel reads them from shared memory
    assert env.engine.kernel_info()["variant_name"] == "tile"
    env.engine.select_kernel(VARIANTS[variant])
    env.simulator._callback_ops = []
    status = env.run({name: pressures[i] for i, name in enumerate(NAMES)}, duration=n_steps * dt - 0.5 * dt,
                     disable_progress_bar=True, check_nan=True)
    assert round(env.time / dt) == n_steps and not status.position_nan_status
    if variant == "tile":  # the optimistic kernel's own result, not the exact-path replay's
        assert env.engine.replay_count() == 0 and env.engine.unresolved_count() == 0

    ref = br2_port.OracleEnvironment(time_step=dt)
    ref.reset(rod_db, assembly, gravity=True)
    fused = FusedOracle.from_simulator(ref.simulator, NAMES)
    W = fused.pack(batch)
    fused.run(W, pressures, n_steps, 0.0, dt)
    x_scale = max(np.abs(fused.view(W, r, "x")).max() for r in range(6))
    for r, rod in enumerate(env.simulator):
        assert np.abs(rod.position_collection - fused.view(W, r, "x")).max() <= POS_TOL * x_scale
        assert np.abs(rod.director_collection - fused.view(W, r, "Q")).max() <= DIR_TOL
    tips = env.simulator[3].position_collection[:, 1, -1]
    assert 0.25 < tips.mean() < 0.31 and tips.std() > 1e-5  # second segment sits on top of the first; instances differ
    env.close()


@pytest.mark.parametrize("assembly", ["single", "double", "triple-as-cluster"])
def test_work_queue_grouping_does_not_change_a_bit(assembly, monkeypatch):
    """The tile kernel's persistent CTAs pull (chunk, instance) items in groups of instances (csrc/br2_tile.cuh).  The order
    items are handed out in is scheduling only: tiny groups -- where a chunk IS handed out before the previous chunk of the
    same instance has been written back, so the progress wait really waits -- a short last group, and one group for the
    whole batch must produce identical bits."""
    batch, n_steps, dt = 200, 1000, 2.0e-5  # 8 chunks of 125 steps
    if assembly == "triple-as-cluster":  # three CTAs per instance: the cluster pulls the items
        monkeypatch.setenv("BR2_TILE_ONE_CTA_THREADS", "100")
        assembly = "triple"
    pressures = np.random.default_rng(21).uniform(0.0, 40.0, size=(3, batch)) * PSI
    action = {name: pressures[i] for i, name in enumerate(NAMES)}
    states = []
    for group in ("0", "64", "7"):
        monkeypatch.setenv("BR2_TILE_QUEUE_GROUP", group)
        env = make_env(assembly, batch, variant="tile", gravity=True)
        env.simulator._callback_ops = []
        env.run(action, duration=n_steps * dt - 0.5 * dt, disable_progress_bar=True)
        assert env.engine.replay_count() == 0
        states.append(env.engine.state.cpu().numpy().copy())  # the whole state block, derived arrays included
        env.close()
    assert np.isfinite(states[0]).all()
    np.testing.assert_array_equal(states[0], states[1])
    np.testing.assert_array_equal(states[0], states[2])


def test_terminal_metrics_kernel_matches_the_reference_formulas():
    """br2_metrics (one device pass over the batch state) against the reference's own numpy expressions
    (/root/reference/br2/environment.py:289-349) evaluated on host copies: steady-state mode 1 and mode 2, NaN flags,
    per instance; one instance is poisoned with a NaN to exercise numpy's max / nanmax semantics."""
    batch, dt = 5, 2.0e-5
    pressures = np.random.default_rng(5).uniform(0.0, 40.0, size=(3, batch)) * PSI
    action = {name: pressures[i] for i, name in enumerate(NAMES)}
    env = make_env("double", batch)
    env.simulator._callback_ops = []
    env.run(action, duration=200 * dt - 0.5 * dt, disable_progress_bar=True)
    fields = ("position_collection", "velocity_collection", "acceleration_collection", "director_collection",
              "omega_collection", "alpha_collection")
    before = {f: np.concatenate([getattr(rod, f) for rod in env.simulator], axis=-1) for f in fields}
    status = env.run(action, duration=150 * dt - 0.5 * dt, disable_progress_bar=True, check_nan=True, check_steady_state=2)
    after = {f: np.concatenate([getattr(rod, f) for rod in env.simulator], axis=-1) for f in fields}
    for short, f in zip(("x", "v", "a", "Q", "w", "alpha"), fields):
        axis = (1, 2) if short == "Q" else 1
        want = np.nanmax(np.linalg.norm(before[f] - after[f], axis=axis), axis=-1)  # per instance
        got = status.steady_state_deltas[short]
        assert got.shape == (batch,) and np.abs(got - want).max() <= 1e-13 * max(want.max(), 1e-300), short
    assert not any(status.nan_per_instance[k].any() for k in status.nan_per_instance)
    # mode 1 + NaN semantics: poison one node of instance 2
    x = env.simulator[3].position_collection
    x[2, 1, 7] = np.nan
    env.simulator[3].position_collection = x
    status = env.run(action, duration=None, disable_progress_bar=True, check_nan=True, check_steady_state=1)
    after_x = np.concatenate([rod.position_collection for rod in env.simulator], axis=-1)
    want = np.linalg.norm(after_x, axis=1).max(axis=-1)
    clean = [0, 1, 3, 4]
    assert np.isnan(status.max_velocity_per_instance[2]) and np.isnan(want[2]) and np.isnan(status.max_velocity)
    assert np.abs(status.max_velocity_per_instance[clean] - want[clean]).max() <= 1e-14
    assert list(status.nan_per_instance["position"]) == [False, False, True, False, False]
    assert status.position_nan_status
    env.close()
