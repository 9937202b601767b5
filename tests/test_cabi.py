"""The C-ABI library: loads, exports every symbol include/br2cuda.h declares, and its host-only
entry points behave (no compute calls here -- those are the -m gpu tests)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT

from br2 import _native


def header_symbols():
    text = open(os.path.join(ROOT, "include", "br2cuda.h")).read()
    return sorted(set(re.findall(r"\b(br2_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _native.load()
    names = header_symbols()
    assert set(names) == set(_native.EXPORTS)
    for name in names:
        assert getattr(lib, name) is not None
    assert lib.br2_abi_version() == _native.ABI_VERSION


def test_count_steps_reproduces_float_accumulation():
    # duration 1.0 at dt=2e-5 is 50001 steps, 0.2 is 10000
    assert _native.count_steps(0.0, 1.0, 2.0e-5)[0] == 50001
    n, t_end = _native.count_steps(0.0, 0.2, 2.0e-5)
    assert n == 10000 and t_end == 0.20000000000005924
    assert _native.count_steps(0.3, 2.0e-5, 2.0e-5)[0] == 1


def test_errors_are_reported_not_thrown():
    lib = _native.load()
    assert lib.br2_step(None, 1, 0.0, 1e-5, None, None) < 0
    assert b"NULL handle" in lib.br2_last_error()
    prog = _native.Br2Program()
    prog.abi_version = 99
    handle = ctypes.c_void_p()
    assert lib.br2_create(ctypes.byref(prog), 0, ctypes.byref(handle)) == -1
    assert b"abi_version" in lib.br2_last_error()


def test_engine_refuses_to_run_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from br2.environment import Environment
    from conftest import DATA

    env = Environment("nocuda", create_dirs=False)
    with pytest.raises(_native.Br2Error, match="no CPU fallback"):
        env.reset(os.path.join(DATA, "rod_library.json"), os.path.join(DATA, "assembly_single.json"))


def test_capture_plan_matches_the_reference_step_test():
    """br2_capture_plan == the reference's per-step test `round(time / dt) % step_skip == 0` on the fp64-accumulated time
    (/root/reference/br2/free_simulator.py:49-51), for one and for two collectors, from a non-zero start time."""
    from br2 import _native

    for t0, n_steps, dt, every in ((0.0, 50001, 2.0e-5, [2000]), (0.04, 4100, 2.0e-5, [2000, 300]), (0.0, 7, 1.0e-3, [None])):
        want, time = [], t0
        for ordinal in range(1, n_steps + 1):
            time += 0.5 * dt
            time += 0.5 * dt
            step = round(time / dt)
            if any(e is None or step % e == 0 for e in every):
                want.append((ordinal, time))
        assert _native.capture_plan(t0, n_steps, dt, every) == want
    assert len(_native.capture_plan(0.0, 50001, 2.0e-5, [2000])) == 25
