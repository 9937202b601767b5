"""Run in a SUBPROCESS by tests/test_adapter.py (needs /root/reference, i.e. the build container).

Builds an assembly with the reference's REAL ``br2.free_simulator.FreeAssembly`` / operator classes on top of
``oracle.mini_elastica`` aliased as ``elastica`` (exactly as tests/golden/make_golden.py does), finalizes it through the
reference's own ``Environment.reset``, and lowers that simulator -- whose operators are the reference's class instances,
without any ``lower_*`` method -- through the product's ``br2._adapter.lower_simulator``.  The product package is loaded
under the alias ``br2_b200`` because the name ``br2`` belongs to the reference in this process.  Tables go to an .npz.
"""
import importlib.util
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join(ROOT, ".numba_cache"))

from oracle import mini_elastica  # noqa: E402

mini_elastica.install_as_elastica()
for name in ("br2.visualize", "br2.visualize.post_processing", "br2.visualize.twist_angle"):
    sys.modules[name] = types.ModuleType(name)
sys.modules["br2.visualize.post_processing"].plot_video_with_surface = lambda *a, **k: None
sys.modules["br2.visualize.twist_angle"].visual_twist_with_surface = lambda *a, **k: None
sys.path.insert(0, REF)
import br2.environment as ref_env  # noqa: E402  (the reference's real module)

assert ref_env.__file__.startswith(REF), ref_env.__file__
sys.modules["br2.visualize"].__path__ = []

pkg_dir = os.path.join(ROOT, "br2-simulator_b200", "br2")
spec = importlib.util.spec_from_file_location("br2_b200", os.path.join(pkg_dir, "__init__.py"),
                                              submodule_search_locations=[pkg_dir])
product = importlib.util.module_from_spec(spec)
sys.modules["br2_b200"] = product
spec.loader.exec_module(product)
from br2_b200 import _adapter  # noqa: E402


def main(rod_db, assembly, gravity, out_path):
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp(prefix="br2_adapter_"))  # Environment() creates result_<tag>/ under the cwd
    try:
        env = ref_env.Environment(run_tag="adapter")
        env.reset(rod_database_path=rod_db, assembly_config_path=assembly, gravity=bool(int(gravity)))
    finally:
        os.chdir(cwd)
    ops = [op for _, op in env.simulator._ext_forces_torques] + [op for _, op in env.simulator._constraints] + \
          [c[-1] for c in env.simulator._connections] + [op for _, op in env.simulator._dampers]
    assert ops and not any(hasattr(op, name) for op in ops for name in ("lower_forcing", "lower_constraint", "lower_joint", "lower_damper"))
    assert any(type(op).__module__.startswith("br2.") and type(op).__name__ == "FreeBaseEndSoftFixed" for op in ops)
    prog = _adapter.lower_simulator(env.simulator)
    payload = {"table_" + k: v for k, v in prog.tables.items()}
    payload.update({"count_" + k: np.array(v) for k, v in prog.counts.items()})
    payload.update({"fast_" + k: np.array(v) for k, v in prog.fast.items() if k != "why"})
    payload["words"] = np.array(prog.words)
    payload["n_actions_lists"] = np.array(len(prog.action_lists))
    np.savez(out_path, **payload)


if __name__ == "__main__":
    main(*sys.argv[1:5])
