"""The optional fp32 mode of BASELINE.json's north_star, measured instead of argued.

`oracle/fused/br2_oracle.c` builds a second time with its arithmetic type set to float (time stays fp64).  That is the
reference's arithmetic, operation for operation, in single precision -- the best a straight fp32 port could do.  It does
not drift: it dies.  The guards the reference relies on (1e-10 inside the arccos of the curvature's log map, 1e-14 under
the sine and in the axis normalisation) are below fp32 resolution, so the first Voronoi cell whose rotation rounds to the
identity produces 0 / 0.  An fp32 mode therefore needs different arithmetic (other guards, small-angle branches), i.e.
it could not be held to "the reference's results under a looser bound" at all; fp64 is the only mode of this package."""
import os

import numpy as np

from conftest import DATA


def test_single_precision_build_of_the_reference_arithmetic_goes_non_finite():
    from oracle import br2_port
    from oracle.fused import FusedOracle

    dt, batch = 2.0e-5, 4
    pressures = np.random.default_rng(0).uniform(0.0, 40.0, size=(3, batch)) * 6895.0
    end = {}
    for real in (np.float64, np.float32):
        ref = br2_port.OracleEnvironment(time_step=dt)
        ref.reset(os.path.join(DATA, "rod_library.json"), os.path.join(DATA, "assembly_single.json"), gravity=True)
        fused = FusedOracle.from_simulator(ref.simulator, ["action1", "action2", "action3"], real=real)
        W = fused.pack(batch)
        assert W.dtype == real
        fused.run(W, pressures, 500, 0.0, dt)
        end[real] = W
    assert np.isfinite(end[np.float64]).all()
    assert not np.isfinite(end[np.float32]).all()   # measured: the first NaN appears after 102 steps
