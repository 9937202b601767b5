"""Integration path B: a simulator built with the reference's OWN classes lowers to the same device tables.

CPU only, and only where /root/reference exists (the build container): the reference's real ``br2`` package is imported
in a subprocess (tests/adapter/dump_reference_tables.py), its ``FreeAssembly`` builds single / double / triple on
``oracle.mini_elastica`` standing in for PyElastica, and ``br2._adapter.lower_simulator`` lowers the finalized simulator.
The tables must equal the ones the product's own builder lowers from the same JSON files."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import DATA

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "br2")), reason="needs the reference sources (build container only)")

ROD_DB = os.path.join(DATA, "rod_library.json")


def product_program(assembly, gravity):
    from br2.free_simulator import FreeAssembly

    class _Env:  # what FreeAssembly reads of its environment
        time_step = 2.0e-5

    assy = FreeAssembly(_Env(), gravity=gravity)
    assy.build(ROD_DB, os.path.join(DATA, "assembly_%s.json" % assembly))
    return assy.simulator.lower()


@pytest.mark.parametrize("assembly,gravity", [("single", True), ("double", False), ("triple", True)])
def test_reference_built_simulator_lowers_to_the_product_tables(tmp_path, assembly, gravity):
    out = tmp_path / "tables.npz"
    subprocess.check_call([sys.executable, os.path.join(HERE, "adapter", "dump_reference_tables.py"), ROD_DB,
                           os.path.join(DATA, "assembly_%s.json" % assembly), str(int(gravity)), str(out)])
    got = np.load(out)
    want = product_program(assembly, gravity)
    for key, value in want.counts.items():
        assert int(got["count_" + key]) == value, key
    assert int(got["words"]) == want.words and int(got["n_actions_lists"]) == len(want.action_lists)
    for key in ("ok", "uniform", "point_forces", "n_act", "ramp_up_time"):
        assert got["fast_" + key] == want.fast[key], key
    for name, table in want.tables.items():
        other = got["table_" + name]
        assert other.shape == table.shape and other.dtype == table.dtype, name
        if table.dtype.kind == "i":
            np.testing.assert_array_equal(other, table, err_msg=name)
        else:
            np.testing.assert_allclose(other, table, rtol=1e-13, atol=0.0, err_msg=name)


def test_subclass_that_overrides_the_arithmetic_is_refused():
    """No CPU fallback: a foreign forcing class whose apply_torques differs from the recognised parent's cannot be lowered."""
    from br2 import _adapter
    from br2.hooks.ops import NotLowerable

    class NoForces:
        pass

    class FreeTwistActuation(NoForces):
        def __init__(self):
            self.actuation_ref, self.scale, self.ramp_up_time = [0.0], 1.0, 0.2

        def apply_torques(self, system, time=0.0):
            pass

    class Mine(FreeTwistActuation):
        def apply_torques(self, system, time=0.0):
            raise RuntimeError("host arithmetic")

    class Unknown(NoForces):
        def apply_forces(self, system, time=0.0):
            pass

    assert type(_adapter.adapt(FreeTwistActuation(), _adapter._FORCING, "forcing")).__name__ == "FreeTwistActuation"
    with pytest.raises(NotLowerable, match="overrides apply_torques"):
        _adapter.adapt(Mine(), _adapter._FORCING, "forcing")
    with pytest.raises(NotLowerable, match="overrides apply_forces"):
        _adapter.adapt(Unknown(), _adapter._FORCING, "forcing")
