"""Device-side diagnostics captures.

The reference's ``FreeCallback.make_callback`` (/root/reference/br2/free_simulator.py:49-73) copies thirteen numpy
arrays per rod every ``step_skip`` steps.  Doing the same for a batch means a device-to-host copy of the whole batch
state plus three host memcpys of it per capture -- for 4096 instances 160 ms, five times the 2000 integration steps
between two captures.  Here a capture is ONE device-to-device clone of the batch state (0.1 ms); the callback
dictionaries (``Environment.data_rods``) keep the reference's keys, order and shapes, but their entries are references
into the clone that turn into numpy arrays when somebody reads them (``sink["position"][3]``, ``np.array(sink[...])``,
``save_data``, ``gather_data``).  The first read of a capture moves it to the host once, for all rods and fields.

Falls back to an immediate host copy when device memory is short.
"""
from collections import defaultdict

import numpy as np


class StateCapture:
    """One snapshot of the batch state ``[batch, words]``: a device clone until first read, then a host array."""

    def __init__(self, engine, tensor=None, host=None, batch=None):
        self.layout = engine.program.layout
        self.n_elems = engine.program.n_elems
        self.batch = engine.batch if batch is None else int(batch)  # (a gathered capture holds the global batch)
        self._tensor = tensor
        self._host = host

    def host(self):
        if self._host is None:
            self._host = self._tensor.cpu().numpy()
            self._tensor = None  # the device copy is not needed any more
        return self._host

    def field(self, rod_index, short, shape):
        off, size = self.layout[rod_index][short]
        block = self.host()[:, off:off + size].reshape((self.batch,) + tuple(shape))
        return block[0].copy() if self.batch == 1 else block.copy()

    @property
    def on_device(self):
        return self._tensor is not None


class _Lazy:
    """Entry of a CaptureSeries that is still a reference into a StateCapture: one field of one rod, optionally passed
    through ``post`` (centre of mass).  ``rebind`` points the same entry at another capture (the gathered one)."""
    __slots__ = ("capture", "rod_index", "short", "shape", "post")

    def __init__(self, capture, rod_index, short, shape, post=None):
        self.capture, self.rod_index, self.short, self.shape, self.post = capture, rod_index, short, tuple(shape), post

    def make(self):
        value = self.capture.field(self.rod_index, self.short, self.shape)
        return value if self.post is None else self.post(value)

    def rebind(self, capture):
        return _Lazy(capture, self.rod_index, self.short, self.shape, self.post)


class CaptureSeries:
    """List-like history of one callback key.  Entries appended as ``_Lazy`` are materialised (and cached) on access."""

    def __init__(self):
        self._items = []

    def append(self, value):
        self._items.append(value)

    def _get(self, index):
        value = self._items[index]
        if isinstance(value, _Lazy):
            value = value.make()
            self._items[index] = value
        return value

    def __len__(self):
        return len(self._items)

    def __getitem__(self, index):
        if isinstance(index, slice):
            return [self._get(i) for i in range(*index.indices(len(self._items)))]
        if index < 0:
            index += len(self._items)
        return self._get(index)

    def __iter__(self):
        for i in range(len(self._items)):
            yield self._get(i)

    def __array__(self, dtype=None, copy=None):
        out = np.array(list(self))
        return out if dtype is None else out.astype(dtype, copy=False)

    def __eq__(self, other):
        return list(self) == list(other)

    def __repr__(self):
        return "CaptureSeries(%d entries, %d still on the device)" % (
            len(self._items), sum(isinstance(v, _Lazy) for v in self._items))

    def clear(self):
        self._items.clear()


class CaptureSink(defaultdict):
    """``callback_params`` of the reference (a ``defaultdict(list)``) whose lists can hold lazy entries."""

    def __init__(self):
        super().__init__(CaptureSeries)


def capture_state(engine):
    """Clone the batch state on the device if memory allows, else copy it to the host right away."""
    torch = engine.torch
    nbytes = engine.state.numel() * engine.state.element_size()
    try:
        free, _ = torch.cuda.mem_get_info(engine.device)
    except Exception:
        free = 0
    if free > 2 * nbytes + (1 << 30):
        return StateCapture(engine, tensor=engine.state.clone())
    return StateCapture(engine, host=engine.state.cpu().numpy())
