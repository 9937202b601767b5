"""Generate tests/golden/* by executing the reference's REAL ``br2`` sources.

Run once, in the build container (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

How: PyElastica -- the reference's third-party dependency holding the Cosserat-rod arithmetic --
is neither vendored in /root/reference nor installable here, so ``oracle.mini_elastica`` (our
restatement of it) is aliased as ``elastica`` in ``sys.modules`` and the matplotlib-only
``br2.visualize`` modules are stubbed.  Everything else -- ``br2.environment.Environment``,
``br2.free_simulator.FreeAssembly``, the fibre actuation / soft-fixed base / tip-to-base joint
operators, the numba rotation helpers -- is the reference's own unmodified code, imported from
/root/reference.  The resulting vectors therefore pin (a) the reference's in-repo math exactly and
(b) the assembly/operator/run-loop semantics of ``oracle.br2_port`` and of the product's host
layer.  They do NOT pin the upstream PyElastica kernels (parity unpinned, see oracle/__init__.py).
"""
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache_golden")

from oracle import mini_elastica  # noqa: E402

mini_elastica.install_as_elastica()
for name in ("br2.visualize", "br2.visualize.post_processing", "br2.visualize.twist_angle"):
    sys.modules[name] = types.ModuleType(name)
sys.modules["br2.visualize.post_processing"].plot_video_with_surface = lambda *a, **k: None
sys.modules["br2.visualize.twist_angle"].visual_twist_with_surface = lambda *a, **k: None
sys.path.insert(0, REF)

import br2.environment as ref_env  # noqa: E402  (the reference's real module)
import br2.free_custom_systems as ref_ops  # noqa: E402
import br2.free_simulator as ref_sim  # noqa: E402
import br2.legacy_surface_connection_parallel_rod_numba as ref_legacy  # noqa: E402
import br2.custom_activation as ref_act  # noqa: E402
import br2.custom_constraint as ref_con  # noqa: E402

assert ref_env.__file__.startswith(REF), ref_env.__file__
sys.modules["br2.visualize"].__path__ = []

ROD_DB = os.path.join(REF, "sample_database/sample_rod_library.json")
ASSEMBLIES = {
    "single": os.path.join(REF, "sample_assembly/single_br2_v1.json"),
    "double": os.path.join(REF, "sample_assembly/double_br2_v1.json"),
    "triple": os.path.join(REF, "sample_assembly/triple_br2_v1.json"),
}
PSI = 6895.0
sys.path.insert(0, HERE)


def rot(axis, angle):
    axis = np.asarray(axis, dtype=np.float64)
    axis = axis / np.linalg.norm(axis)
    return ref_legacy._single_get_rotation_matrix(angle, axis)


# ----------------------------------------------------------------------------- helpers
def golden_helpers():
    rng = np.random.default_rng(1234)
    axes = rng.normal(size=(8, 3))
    axes /= np.linalg.norm(axes, axis=1)[:, None]
    # generic, tiny (near-identity branch) and near-pi (third branch) angles
    angles = np.array([0.3, -1.1, 2.5, 1e-4, 3e-7, np.pi - 2e-7, np.pi - 1e-9, 0.0])
    mats = np.stack([ref_legacy._single_get_rotation_matrix(a, ax) for a, ax in zip(angles, axes)])
    logs = np.stack([ref_legacy._single_inv_rotate(m.copy()) for m in mats])
    batch = np.ascontiguousarray(np.moveaxis(mats, 0, -1))
    logs_batch = ref_legacy._inv_rotate(batch)
    kat_R = ref_legacy._single_get_rotation_matrix(0.3, np.array([0.0, 0.0, 1.0]))
    kat_log = ref_legacy._single_inv_rotate(kat_R.copy())
    kat_batch = ref_legacy._inv_rotate(np.ascontiguousarray(np.stack([kat_R, kat_R @ kat_R], axis=-1)))
    np.savez(
        os.path.join(HERE, "helpers.npz"),
        axes=axes,
        angles=angles,
        mats=mats,
        logs=logs,
        logs_batch=logs_batch,
        kat_R=kat_R,
        kat_log=kat_log,
        kat_batch=kat_batch,
    )


# ----------------------------------------------------------------------------- operators
class _FakeRod:
    def __init__(self, n, rng):
        self.n_elems = n
        self.position_collection = rng.normal(size=(3, n + 1))
        self.velocity_collection = rng.normal(size=(3, n + 1))
        self.acceleration_collection = rng.normal(size=(3, n + 1))
        self.omega_collection = rng.normal(size=(3, n))
        self.alpha_collection = rng.normal(size=(3, n))
        q = np.stack([rot(rng.normal(size=3), rng.uniform(-2, 2)) for _ in range(n)], axis=-1)
        self.director_collection = np.ascontiguousarray(q)
        self.radius = rng.uniform(0.005, 0.01, size=n)
        self.tangents = rng.normal(size=(3, n))
        self.external_forces = rng.normal(size=(3, n + 1))
        self.external_torques = rng.normal(size=(3, n))


ROD_FIELDS = (
    "position_collection",
    "velocity_collection",
    "acceleration_collection",
    "omega_collection",
    "alpha_collection",
    "director_collection",
    "radius",
    "external_forces",
    "external_torques",
)


def _dump_rod(prefix, rod, out):
    for field in ROD_FIELDS:
        out[prefix + field] = getattr(rod, field).copy()


def golden_ops():
    rng = np.random.default_rng(99)
    out = {}
    # --- tip-to-base joint, forces then the overwrite "torques" (legacy...:364-478)
    r1, r2 = _FakeRod(5, rng), _FakeRod(7, rng)
    _dump_rod("tip_in_r1_", r1, out)
    _dump_rod("tip_in_r2_", r2, out)
    l1, l2 = rng.normal(size=3), rng.normal(size=3)
    out["tip_l1"], out["tip_l2"] = l1, l2
    out["tip_k"], out["tip_nu"] = 5763.67, 4.39e-6
    joint = ref_legacy.TipToTipStraightJoint(
        k=5763.67, nu=4.39e-6, kt=0.1, rod1_rd2_local=l1, rod2_rd2_local=l2
    )
    joint.apply_forces(r1, -1, r2, 0)
    joint.apply_torques(r1, -1, r2, 0)
    _dump_rod("tip_out_r1_", r1, out)
    _dump_rod("tip_out_r2_", r2, out)

    # --- soft-fixed base (free_custom_systems.py:104-185)
    rod = _FakeRod(6, rng)
    fixed_x = rod.position_collection[:, 0] + np.array([1e-4, -2e-4, 3e-5])
    fixed_Q = rot([0.2, 0.5, -1.0], 0.7)
    _dump_rod("pin_in_", rod, out)
    out["pin_fixed_x"], out["pin_fixed_Q"] = fixed_x, fixed_Q
    for tag, k, nu in (("a", 1e9, 0.0), ("b", 2.5e3, 0.125)):
        work = _FakeRod(6, np.random.default_rng(0))
        for field in ROD_FIELDS:
            setattr(work, field, getattr(rod, field).copy())
        bc = ref_ops.FreeBaseEndSoftFixed(fixed_x.copy(), fixed_Q.copy(), k=k, nu=nu, kt=0.0, _system=work)
        bc.constrain_values(work, 0.1)
        _dump_rod("pin_%s_values_" % tag, work, out)
        bc.constrain_rates(work, 0.1)
        _dump_rod("pin_%s_rates_" % tag, work, out)
    # exactly-on-target branch (|d| <= 1e-8)
    work = _FakeRod(6, np.random.default_rng(0))
    for field in ROD_FIELDS:
        setattr(work, field, getattr(rod, field).copy())
    bc = ref_ops.FreeBaseEndSoftFixed(
        work.position_collection[:, 0] + 5e-9, fixed_Q.copy(), k=1e9, nu=3.0, kt=0.0, _system=work
    )
    bc.constrain_values(work, 0.0)
    _dump_rod("pin_c_values_", work, out)

    # --- fibre actuation (free_custom_systems.py:20-101)
    for tag, t in (("ramp", 0.05), ("full", 0.7)):
        work = _FakeRod(9, np.random.default_rng(5))
        ref = [31.5 * PSI]
        ref_ops.FreeBendActuation(ref, z_angle=np.deg2rad(60.0), scale=1.00518e-6).apply_torques(work, t)
        out["bend_%s" % tag] = work.external_torques.copy()
        work = _FakeRod(9, np.random.default_rng(5))
        ref_ops.FreeTwistActuation(ref, scale=-1.25797e-6).apply_torques(work, t)
        out["twist_%s" % tag] = work.external_torques.copy()
    out["act_in_torques"] = _FakeRod(9, np.random.default_rng(5)).external_torques

    # --- optional ops (custom_activation.py / custom_constraint.py)
    work = _FakeRod(10, np.random.default_rng(6))
    out["pf_in"] = work.external_forces.copy()
    ref_act.PointForces(np.array([0.1, -0.2, 0.3]), 0.37, ramp_up_time=1e-3).apply_forces(work, 4e-4)
    out["pf_out"] = work.external_forces.copy()
    work = _FakeRod(10, np.random.default_rng(6))
    bc = ref_con.LastEndFixedRod(np.array([1.0, 2.0, 3.0]), fixed_Q.copy(), _system=work)
    bc.constrain_values(work, 0.0)
    bc.constrain_rates(work, 0.0)
    _dump_rod("lef_out_", work, out)
    np.savez(os.path.join(HERE, "ops.npz"), **out)


# ----------------------------------------------------------------------------- registration
def _jsonable(value):
    if isinstance(value, np.ndarray):
        return value.tolist()
    if isinstance(value, (np.floating, np.integer)):
        return value.item()
    if isinstance(value, (list, tuple)):
        return [_jsonable(v) for v in value]
    return value


def registration_log(simulator):
    """Everything ``FreeAssembly.build`` registered, in execution order, as plain data."""
    log = {"n_systems": len(simulator), "forcing": [], "constraints": [], "dampers": [], "connections": []}
    for sys_id, op in simulator._ext_forces_torques:
        entry = {"rod": sys_id, "cls": type(op).__name__}
        for key in ("acc_gravity", "z_angle", "magnitude_scale", "scale", "ramp_up_time"):
            if hasattr(op, key):
                entry[key] = _jsonable(getattr(op, key))
        log["forcing"].append(entry)
    for sys_id, op in simulator._constraints:
        log["constraints"].append(
            {
                "rod": sys_id,
                "cls": type(op).__name__,
                "k": _jsonable(op.k),
                "nu": _jsonable(op.nu),
                "kt": _jsonable(op.kt),
                "fixed_position": _jsonable(np.asarray(op.fixed_position)),
                "fixed_directors": _jsonable(np.asarray(op.fixed_directors)),
            }
        )
    for sys_id, op in simulator._dampers:
        log["dampers"].append(
            {
                "rod": sys_id,
                "cls": type(op).__name__,
                "translational": _jsonable(np.asarray(op.translational_damping_coefficient)),
                "rotational": _jsonable(np.asarray(op.rotational_damping_coefficient)),
            }
        )
    for first, second, idx1, idx2, op in simulator._connections:
        entry = {"rod_one": first, "rod_two": second, "idx_one": idx1, "idx_two": idx2,
                 "cls": type(op).__name__, "k": _jsonable(np.asarray(op.k)), "nu": _jsonable(np.asarray(op.nu))}
        if type(op).__name__ == "SurfaceJointSideBySide":
            entry["k_repulsive"] = _jsonable(np.asarray(op.k_repulsive))
            entry["dv1"] = _jsonable(np.asarray(op.rod_one_direction_vec_in_material_frame))
            entry["dv2"] = _jsonable(np.asarray(op.rod_two_direction_vec_in_material_frame))
            entry["offset"] = _jsonable(np.asarray(op.offset_btw_rods))
        else:
            entry["kt"] = _jsonable(op.kt)
            entry["l1"] = _jsonable(np.asarray(op.rod1_rd2_local))
            entry["l2"] = _jsonable(np.asarray(op.rod2_rd2_local))
        log["connections"].append(entry)
    return log


from cases import CALLBACK_KEYS, STATE_FIELDS, TRAJECTORIES  # noqa: E402


def golden_trajectories():
    cwd = os.getcwd()
    scratch = tempfile.mkdtemp(prefix="br2_golden_")
    os.chdir(scratch)  # Environment() creates result_<tag>/ under the cwd
    try:
        for name, assembly, kwargs, action, duration, fps in TRAJECTORIES:
            env = ref_env.Environment(run_tag=name, rendering_fps=fps)
            env.reset(rod_database_path=ROD_DB, assembly_config_path=ASSEMBLIES[assembly], **kwargs)
            with open(os.path.join(HERE, "registration_%s.json" % name), "w") as fh:
                json.dump(registration_log(env.simulator), fh)
            status = env.run(action=action, duration=duration, disable_progress_bar=True,
                             check_nan=True, check_steady_state=1)
            out = {"final_time": env.time, "step_skip": env.step_skip,
                   "max_velocity": status.max_velocity,
                   "nan_flags": np.array([status.position_nan_status, status.velocity_nan_status,
                                          status.director_nan_status, status.omega_nan_status])}
            for idx, rod in enumerate(env.shearable_rods.values()):
                for field in STATE_FIELDS:
                    out["rod%d_%s" % (idx, field)] = getattr(rod, field).copy()
                for key in CALLBACK_KEYS:
                    out["rod%d_cb_%s" % (idx, key)] = np.array(env.data_rods[idx][key])
            out["rod_names"] = np.array(list(env.shearable_rods.keys()))
            np.savez_compressed(os.path.join(HERE, "traj_%s.npz" % name), **out)
            print(name, "steps:", round(env.time / env.time_step), "captures:",
                  len(env.data_rods[0]["time"]), "tip:", list(env.shearable_rods.values())[0].position_collection[:, -1])
    finally:
        os.chdir(cwd)


if __name__ == "__main__":
    golden_helpers()
    golden_ops()
    golden_trajectories()
    print("golden vectors written to", HERE)
