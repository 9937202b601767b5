"""ctypes binding of libbr2cuda.so (C ABI: include/br2cuda.h).

The library is built in-tree (``br2-simulator_b200/lib/libbr2cuda.so``) by
``__graft_entry__.build()``.  If it is missing or fails to load this module raises: the package has
no CPU fallback.
"""
import ctypes
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("BR2_CUDA_LIB") or os.path.join(PKG_DIR, "lib", "libbr2cuda.so")
ABI_VERSION = 2
METRICS_N = 12  # BR2_METRICS_N

_i32p = ctypes.POINTER(ctypes.c_int32)
_f64p = ctypes.POINTER(ctypes.c_double)


class Br2Program(ctypes.Structure):
    """``struct br2_program`` of include/br2cuda.h."""

    _fields_ = [
        ("abi_version", ctypes.c_int32),
        ("n_slots", ctypes.c_int32), ("n_rods", ctypes.c_int32), ("n_forcing", ctypes.c_int32),
        ("n_constraints", ctypes.c_int32), ("n_joints", ctypes.c_int32), ("n_actions", ctypes.c_int32),
        ("words_per_instance", ctypes.c_int64),
        ("slot_rod", ctypes.c_void_p), ("rod_i", ctypes.c_void_p), ("rod_d", ctypes.c_void_p),
        ("slot_c", ctypes.c_void_p), ("forcing_i", ctypes.c_void_p), ("forcing_d", ctypes.c_void_p),
        ("cons_i", ctypes.c_void_p), ("cons_d", ctypes.c_void_p), ("joint_i", ctypes.c_void_p),
        ("joint_d", ctypes.c_void_p), ("node_ptr", ctypes.c_void_p), ("node_items", ctypes.c_void_p),
        ("elem_ptr", ctypes.c_void_p), ("elem_items", ctypes.c_void_p), ("overwrite_src", ctypes.c_void_p),
        ("fast_ok", ctypes.c_int32), ("fast_uniform", ctypes.c_int32), ("n_act", ctypes.c_int32),
        ("fast_point_forces", ctypes.c_int32), ("act_ramp_up_time", ctypes.c_double),
        ("fast_i", ctypes.c_void_p), ("fast_c", ctypes.c_void_p), ("act_i", ctypes.c_void_p), ("act_d", ctypes.c_void_p),
    ]


class Br2Error(RuntimeError):
    pass


_lib = None

EXPORTS = (
    "br2_last_error", "br2_abi_version", "br2_create", "br2_destroy", "br2_words_per_instance",
    "br2_bind_state", "br2_set_actuation", "br2_step", "br2_step_host", "br2_count_steps",
    "br2_kernel_info", "br2_fp64_peak_tflops", "br2_select_kernel", "br2_selftest_math", "br2_replay_count",
    "br2_replay_reasons", "br2_unresolved_count", "br2_describe_plan", "br2_capture_plan", "br2_metrics", "br2_set_instance_params",
)


def load():
    """Load (once) and return the library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Br2Error(
            "CUDA library %s not found. Build it with `python -c \"import __graft_entry__ as g; g.build()\"` "
            "from the repo root. This package has no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    lib.br2_last_error.restype = ctypes.c_char_p
    lib.br2_abi_version.restype = ctypes.c_int
    lib.br2_create.argtypes = [ctypes.POINTER(Br2Program), ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
    lib.br2_destroy.argtypes = [ctypes.c_void_p]
    lib.br2_words_per_instance.argtypes = [ctypes.c_void_p]
    lib.br2_words_per_instance.restype = ctypes.c_int64
    lib.br2_bind_state.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
    lib.br2_set_actuation.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    lib.br2_step.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_double, ctypes.c_double, ctypes.c_void_p, _f64p]
    lib.br2_step_host.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                  ctypes.c_int64, ctypes.c_double, ctypes.c_double, _f64p]
    lib.br2_count_steps.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                    ctypes.POINTER(ctypes.c_int64), _f64p]
    lib.br2_capture_plan.argtypes = [ctypes.c_double, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p, ctypes.c_int32,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64)]
    lib.br2_set_instance_params.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.br2_metrics.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.br2_kernel_info.argtypes = [ctypes.c_void_p] + [_i32p] * 6
    lib.br2_select_kernel.argtypes = [ctypes.c_void_p, ctypes.c_int32]
    lib.br2_replay_count.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64)]
    lib.br2_describe_plan.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int64]
    lib.br2_unresolved_count.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64)]
    lib.br2_replay_reasons.argtypes = [ctypes.c_void_p] + [ctypes.POINTER(ctypes.c_int64)] * 3
    lib.br2_selftest_math.argtypes = [ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
    lib.br2_fp64_peak_tflops.argtypes = [ctypes.c_int, ctypes.c_int32, _f64p]
    if lib.br2_abi_version() != ABI_VERSION:
        raise Br2Error("libbr2cuda.so ABI %d != expected %d; rebuild" % (lib.br2_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise Br2Error("libbr2cuda error %d: %s" % (rc, load().br2_last_error().decode("utf-8", "replace")))


def make_program(prog):
    """ctypes view of a ``br2._lowering.Program`` (the numpy tables stay owned by ``prog``)."""
    t, c = prog.tables, prog.counts
    return Br2Program(
        ABI_VERSION, c["n_slots"], c["n_rods"], c["n_forcing"], c["n_constraints"], c["n_joints"], c["n_actions"],
        prog.words,
        *[t[name].ctypes.data for name in (
            "slot_rod", "rod_i", "rod_d", "slot_c", "forcing_i", "forcing_d", "cons_i", "cons_d", "joint_i",
            "joint_d", "node_ptr", "node_items", "elem_ptr", "elem_items", "overwrite_src")],
        int(prog.fast["ok"]), int(prog.fast["uniform"]), int(prog.fast["n_act"]), int(prog.fast["point_forces"]),
        float(prog.fast["ramp_up_time"]),
        *[t[name].ctypes.data for name in ("fast_i", "fast_c", "act_i", "act_d")])


def describe_plan(prog):
    """Which kernel the library would pick for a lowered ``Program`` and its layout (host only, no GPU needed)."""
    buf = ctypes.create_string_buffer(512)
    cprog = make_program(prog)
    check(load().br2_describe_plan(ctypes.byref(cprog), buf, len(buf)))
    return buf.value.decode()


def count_steps(t0, duration, dt):
    """(n_steps, t_end) of ``while time < t0 + duration`` with two ``+dt/2`` per step
    (/root/reference/br2/environment.py:278-286)."""
    n = ctypes.c_int64(0)
    t_end = ctypes.c_double(0.0)
    check(load().br2_count_steps(float(t0), float(duration), float(dt), ctypes.byref(n), ctypes.byref(t_end)))
    return n.value, t_end.value


def capture_plan(t0, n_steps, dt, every):
    """[(ordinal, time)] of the steps of a run after which a diagnostics collector fires (``every``: the collectors'
    step_skip values; ``None`` = every step): `round(time / dt) % every == 0` on the fp64-accumulated time,
    /root/reference/br2/free_simulator.py:49-51."""
    every = np.asarray([1 if e is None else int(e) for e in every], dtype=np.int64)
    if every.size == 0 or n_steps <= 0:
        return []
    smallest = max(int(every.min()), 1)
    capacity = n_steps // smallest + 2
    ordinals = np.zeros(capacity, dtype=np.int64)
    times = np.zeros(capacity, dtype=np.float64)
    count = ctypes.c_int64(0)
    check(load().br2_capture_plan(float(t0), int(n_steps), float(dt), every.ctypes.data, int(every.size),
                                  ordinals.ctypes.data, times.ctypes.data, capacity, ctypes.byref(count)))
    if count.value > capacity:  # (cannot happen: at most one stop per `smallest` steps, plus the ends)
        raise Br2Error("br2_capture_plan: %d stops, room for %d" % (count.value, capacity))
    return [(int(o), float(t)) for o, t in zip(ordinals[:count.value], times[:count.value])]
