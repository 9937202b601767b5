"""Per-instance material parameters: sweeps over E, G, damping constant and fibre angles inside
one batch.  CPU: every instance's constants equal, bit for bit, the ones a separate build from a modified rod library
lowers to.  GPU: the batch matches the oracle run once per parameter set."""
import json
import os

import numpy as np
import pytest

from conftest import DATA

ROD_DB = os.path.join(DATA, "rod_library.json")
PSI = 6895.0
NAMES = ["action1", "action2", "action3"]


def sweep(batch, seed=4):
    rng = np.random.default_rng(seed)
    return {"youngs_modulus": rng.uniform(0.7e7, 1.4e7, batch), "shear_modulus": rng.uniform(5.0e6, 8.0e6, batch),
            "damping_constant": rng.uniform(0.2, 1.5, batch), "alpha": rng.uniform(50.0, 70.0, batch),
            "beta": -rng.uniform(50.0, 70.0, batch), "gamma": rng.uniform(-20.0, 20.0, batch)}


def library_of(params, index, path):
    library = json.load(open(ROD_DB))
    for key in ("youngs_modulus", "shear_modulus", "damping_constant"):
        library["DefaultParams"][key] = float(params[key][index])
    for spec in library["Rods"].values():
        for kind in ("alpha", "beta", "gamma"):
            if kind in spec:
                spec[kind] = [float(params[kind][index])] * len(spec[kind])
    with open(path, "w") as handle:
        json.dump(library, handle)
    return str(path)


class _Env:
    time_step = 2.0e-5


def lowered(rod_db, assembly):
    from br2.free_simulator import FreeAssembly

    assy = FreeAssembly(_Env(), gravity=True)
    assy.build(rod_db, os.path.join(DATA, "assembly_%s.json" % assembly), verbose=False)
    return assy, assy.simulator.lower()


@pytest.mark.parametrize("assembly", ["single", "double"])
def test_instance_tables_equal_separate_builds(tmp_path, assembly):
    from br2 import _lowering as L
    from br2 import _sweep

    batch = 4
    params = sweep(batch)
    assy, program = lowered(ROD_DB, assembly)
    uni, act_d, joint_scale = _sweep.instance_tables(assy, program, params, batch, _Env.time_step)
    tip_rows = program.tables["joint_i"][:, 0] == L.JOINT_TIP_TO_BASE
    for b in range(batch):
        _, one = lowered(library_of(params, b, tmp_path / ("lib%d.json" % b)), assembly)
        np.testing.assert_array_equal(uni[b, :L.FC_PIN], one.tables["fast_c"][:L.FC_PIN, 0])
        np.testing.assert_array_equal(act_d[b], one.tables["act_d"])
        np.testing.assert_array_equal(one.tables["act_i"], program.tables["act_i"])
        if tip_rows.any():
            np.testing.assert_allclose(program.tables["joint_d"][tip_rows, 0] * joint_scale[b], one.tables["joint_d"][tip_rows, 0],
                                       rtol=3e-16, atol=0.0)
        # everything the batch shares really is independent of the swept parameters
        for name in ("fast_i", "slot_rod", "rod_i", "node_items", "overwrite_src"):
            np.testing.assert_array_equal(one.tables[name], program.tables[name])
        np.testing.assert_array_equal(one.tables["slot_c"][L.SC_MASS], program.tables["slot_c"][L.SC_MASS])
    with pytest.raises(KeyError):
        _sweep.instance_tables(assy, program, {"density": np.ones(batch)}, batch, _Env.time_step)
    with pytest.raises(ValueError):
        _sweep.instance_tables(assy, program, {"alpha": np.ones(batch + 1)}, batch, _Env.time_step)


@pytest.mark.gpu
@pytest.mark.parametrize("assembly,n_steps", [("single", 5000), ("double", 3000)])
def test_parameter_sweep_in_one_batch_matches_one_oracle_run_per_parameter_set(tmp_path, assembly, n_steps):
    from br2.environment import Environment
    from oracle import br2_port
    from oracle.fused import FusedOracle

    batch, dt = 5, 2.0e-5
    params = sweep(batch, seed=9)
    pressures = np.random.default_rng(10).uniform(5.0, 40.0, size=(3, batch)) * PSI
    env = Environment("sweep", batch=batch, create_dirs=False)
    env.reset(ROD_DB, os.path.join(DATA, "assembly_%s.json" % assembly), gravity=True, per_instance=params)
    env.simulator._callback_ops = []
    assert env.engine.kernel_info()["variant_name"] == "tile"
    status = env.run({name: pressures[i] for i, name in enumerate(NAMES)}, duration=n_steps * dt - 0.5 * dt,
                     disable_progress_bar=True, check_nan=True)
    assert not status.position_nan_status and env.engine.unresolved_count() == 0
    spread = 0.0
    for b in range(batch):
        ref = br2_port.OracleEnvironment(time_step=dt)
        ref.reset(library_of(params, b, tmp_path / ("lib%d.json" % b)), os.path.join(DATA, "assembly_%s.json" % assembly), gravity=True)
        fused = FusedOracle.from_simulator(ref.simulator, NAMES)
        W = fused.pack(1)
        fused.run(W, pressures[:, b:b + 1], n_steps, 0.0, dt)
        x_scale = max(np.abs(fused.view(W, r, "x")).max() for r in range(len(env.simulator)))
        for r, rod in enumerate(env.simulator):
            assert np.abs(rod.position_collection[b] - fused.view(W, r, "x")[0]).max() <= 1e-9 * x_scale, (b, r)
            assert np.abs(rod.director_collection[b] - fused.view(W, r, "Q")[0]).max() <= 1e-9, (b, r)
        spread = max(spread, np.abs(env.simulator[0].position_collection[b] - env.simulator[0].position_collection[0]).max())
    assert spread > 1e-4  # the parameter sets really lead to different motions
    env.close()


@pytest.mark.gpu
def test_run_raises_when_instances_froze_without_an_exact_path(tmp_path):
    """With per-instance material tables the generic replay cannot redo an instance (it reads the shared tables): one that
    leaves the fast kernel's validity range stays flagged and frozen.  ``Environment.run`` must say so instead of handing
    back its stale state."""
    from br2.environment import Environment

    batch, dt = 4, 2.0e-5
    env = Environment("sweep", batch=batch, create_dirs=False)
    env.reset(ROD_DB, os.path.join(DATA, "assembly_single.json"), gravity=True, per_instance=sweep(batch, seed=3))
    env.simulator._callback_ops = []
    pressures = np.array([[30.0, 5.0e4, 10.0, 20.0], [20.0, 5.0e4, 5.0, 20.0], [10.0, 5.0e4, 0.0, 20.0]]) * PSI
    with pytest.raises(RuntimeError, match="validity range"):
        env.run({name: pressures[i] for i, name in enumerate(NAMES)}, duration=400 * dt - 0.5 * dt, disable_progress_bar=True)
    assert env.engine.unresolved_count() >= 1
    env.close()
