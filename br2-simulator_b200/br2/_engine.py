"""Device engine: owns the batch state in HBM (a PyTorch tensor) and drives libbr2cuda through the
C ABI.  PyTorch is used for device memory, streams and copies only; every step is the hand-written
kernel behind ``br2_step``.
"""
import ctypes

import numpy as np

from . import _native
from .hooks.rod import DYNAMIC_FIELDS, STATE_IO

_SHAPES = {short: shape for short, shape in DYNAMIC_FIELDS.values()}


class Engine:
    def __init__(self, program, batch=1, device=None):
        import torch

        if not torch.cuda.is_available():
            raise _native.Br2Error(
                "br2 (B200 build) needs a CUDA device: rod integration runs only in the sm_100a kernels "
                "of libbr2cuda.so and there is no CPU fallback")
        self.torch = torch
        self.lib = _native.load()
        self.program = program
        self.batch = int(batch)
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        if self.device.index is None:  # "cuda": the current device, not device 0
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.words = program.words

        self._cprog = _native.make_program(program)
        handle = ctypes.c_void_p()
        _native.check(self.lib.br2_create(ctypes.byref(self._cprog), self.device.index, ctypes.byref(handle)))
        self.handle = handle

        template = np.zeros(self.words)
        self._template = template
        self.state = torch.zeros((self.batch, self.words), dtype=torch.float64, device=self.device)
        self.n_actions = program.counts["n_actions"]
        self.pressures = torch.zeros((max(self.n_actions, 1), self.batch), dtype=torch.float64, device=self.device)
        _native.check(self.lib.br2_bind_state(self.handle, self.state.data_ptr(), self.batch))
        _native.check(self.lib.br2_set_actuation(self.handle, self.pressures.data_ptr()))
        self.launches = 0
        self._variant = self.kernel_info()["variant"]

    # ------------------------------------------------------------------ state access
    def load_template(self, rods):
        """Initial state of every instance = the host rods' arrays."""
        for r, rod in enumerate(rods):
            for short, (off, size) in self.program.layout[r].items():
                self._template[off:off + size] = np.asarray(rod._host[short], dtype=np.float64).ravel()
        row = self.torch.from_numpy(self._template).to(self.device)
        self.state[:] = row.unsqueeze(0)

    def view(self, rod_index, short):
        """[batch, ...] device view of one field of one rod (no copy)."""
        off, size = self.program.layout[rod_index][short]
        shape = _SHAPES[short](self.program.n_elems[rod_index])
        return self.state[:, off:off + size].view((self.batch,) + tuple(shape))

    def fetch(self, rod_index, short):
        arr = self.view(rod_index, short).cpu().numpy()
        return arr[0] if self.batch == 1 else arr

    def upload(self, rod_index, short, value):
        view = self.view(rod_index, short)
        tensor = self.torch.from_numpy(np.ascontiguousarray(value, dtype=np.float64)).to(self.device)
        view.copy_(tensor.expand_as(view) if tensor.dim() < view.dim() else tensor)

    def snapshot(self):
        """Host copy of the whole batch state, split per rod / field (for diagnostics callbacks)."""
        host = self.state.cpu().numpy()
        out = []
        for r, fields in enumerate(self.program.layout):
            n = self.program.n_elems[r]
            arrays = {}
            for short, (off, size) in fields.items():
                a = host[:, off:off + size].reshape((self.batch,) + tuple(_SHAPES[short](n)))
                arrays[short] = a[0].copy() if self.batch == 1 else a.copy()
            out.append(arrays)
        return out

    # ------------------------------------------------------------------ actuation
    def set_pressures(self, values):
        """``values``: per action slot a scalar or a length-batch array.  Staged in pinned host memory and copied
        asynchronously on the current stream (the next launch is ordered behind it); the staging buffer is reused once
        its previous copy has completed."""
        if self.n_actions == 0:
            return
        torch = self.torch
        if getattr(self, "_pressure_stage", None) is None:
            self._pressure_stage = torch.empty((self.n_actions, self.batch), dtype=torch.float64).pin_memory()
            self._pressure_done = None
        if self._pressure_done is not None:
            self._pressure_done.synchronize()
        host = self._pressure_stage.numpy()
        for a, value in enumerate(values):
            host[a, :] = np.asarray(value, dtype=np.float64)
        self.pressures[: self.n_actions].copy_(self._pressure_stage, non_blocking=True)
        self._pressure_done = torch.cuda.Event()
        self._pressure_done.record(torch.cuda.current_stream(self.device))

    def set_instance_params(self, uni=None, act_d=None, joint_scale=None):
        """Material constants that differ between the instances (br2_set_instance_params, include/br2cuda.h): host arrays
        ``uni [batch, FC_COUNT]``, ``act_d [batch, n_act, 3]``, ``joint_scale [batch]``; ``None`` returns to the shared
        constants.  The device copies are kept alive here (the library borrows them)."""
        if uni is None:
            _native.check(self.lib.br2_set_instance_params(self.handle, None, None, None))
            self._instance_tables = None
            return
        torch = self.torch
        tables = []
        for array, shape in ((uni, (self.batch, -1)), (act_d, (self.batch, -1)), (joint_scale, (self.batch,))):
            host = np.asarray(array, dtype=np.float64)
            if host.size == 0:  # (an assembly without actuation rows)
                tables.append(torch.zeros(1, dtype=torch.float64, device=self.device))
                continue
            host = np.ascontiguousarray(host.reshape(shape))
            if host.shape[0] != self.batch:
                raise ValueError("per-instance table has %d rows, the batch %d" % (host.shape[0], self.batch))
            tables.append(torch.from_numpy(host).to(self.device))
        _native.check(self.lib.br2_set_instance_params(self.handle, *[t.data_ptr() for t in tables]))
        self._instance_tables = tables

    # ------------------------------------------------------------------ stepping
    def step(self, n_steps, t0, dt):
        """Advance all instances ``n_steps`` steps on the current stream; returns the end time."""
        t_end = ctypes.c_double(0.0)
        stream = self.torch.cuda.current_stream(self.device).cuda_stream
        with self.torch.cuda.device(self.device):
            _native.check(self.lib.br2_step(self.handle, int(n_steps), float(t0), float(dt),
                                            ctypes.c_void_p(stream), ctypes.byref(t_end)))
        if n_steps > 0:
            # the fast kernel is followed by the exact-path replay pass (br2_step launches both)
            self.launches += 2 if self._variant and getattr(self, "_instance_tables", None) is None else 1
        return t_end.value

    def metrics(self, previous=None):
        """``[batch, 12]`` device tensor of per-instance terminal metrics (br2_metrics, include/br2cuda.h): NaN mask,
        max |x|, max |v| and -- with ``previous``, a clone of the state block taken before the run -- the six steady-state
        deltas.  One kernel, one pass over the state; asynchronous on the current stream."""
        out = self.torch.empty((self.batch, _native.METRICS_N), dtype=self.torch.float64, device=self.device)
        stream = self.torch.cuda.current_stream(self.device).cuda_stream
        with self.torch.cuda.device(self.device):
            _native.check(self.lib.br2_metrics(self.handle, previous.data_ptr() if previous is not None else None,
                                               out.data_ptr(), ctypes.c_void_p(stream)))
        return out

    def kernel_info(self):
        vals = [ctypes.c_int32(0) for _ in range(6)]
        _native.check(self.lib.br2_kernel_info(self.handle, *[ctypes.byref(v) for v in vals]))
        keys = ("registers", "smem_bytes", "threads", "blocks_per_sm", "sm_count", "variant")
        info = dict(zip(keys, (v.value for v in vals)))
        info["variant_name"] = ("generic", "fast", "fast-uniform", "tile")[info["variant"]]
        return info

    def replay_count(self):
        """Instance-launches redone by the exact-path kernel behind the fast one (synchronises)."""
        count = ctypes.c_int64(0)
        _native.check(self.lib.br2_replay_count(self.handle, ctypes.byref(count)))
        return count.value

    def unresolved_count(self):
        """Instances frozen because they left the optimistic kernel's range and no exact-path kernel fits the
        assembly (cluster-sized assemblies only; synchronises)."""
        count = ctypes.c_int64(0)
        _native.check(self.lib.br2_unresolved_count(self.handle, ctypes.byref(count)))
        return count.value

    def replay_reasons(self):
        """Why instances left the optimistic kernel: counts per series (rotation angle, curvature angle, strain)."""
        vals = [ctypes.c_int64(0) for _ in range(3)]
        _native.check(self.lib.br2_replay_reasons(self.handle, *[ctypes.byref(v) for v in vals]))
        return dict(zip(("rotation", "curvature", "strain"), (v.value for v in vals)))

    def select_kernel(self, variant):
        """0 generic, 1 fast (per-slot constants), 2 fast (uniform constants), 3 tile (several slots per thread), -1 automatic."""
        _native.check(self.lib.br2_select_kernel(self.handle, int(variant)))
        self._variant = self.kernel_info()["variant"]

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            self.lib.br2_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
