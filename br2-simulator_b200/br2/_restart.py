"""Checkpoint / resume of the batch state, in the layout of PyElastica's ``restart`` helpers that the
reference calls through ``FreeAssembly.save_state / load_state``
(/root/reference/br2/free_simulator.py:93-99): one ``system_{idx}.npz`` per rod holding ``time`` plus
the rod's arrays under their attribute names.  Loading copies the arrays back to the device, checks
that every file agrees on ``time`` and returns it (the reference drops the value).
"""
import os

import numpy as np

from .hooks.rod import DYNAMIC_FIELDS

_CONSTANT_ATTRS = ("rest_lengths", "rest_voronoi_lengths", "volume", "mass", "density", "shear_matrix",
                   "bend_matrix", "mass_second_moment_of_inertia", "inv_mass_second_moment_of_inertia",
                   "rest_sigma", "rest_kappa")


def save_state(simulator, directory="", time=0.0, verbose=False):
    os.makedirs(directory, exist_ok=True)
    for idx, rod in enumerate(simulator):
        arrays = {name: np.asarray(getattr(rod, name)) for name in DYNAMIC_FIELDS}
        arrays.update({name: np.asarray(getattr(rod, name)) for name in _CONSTANT_ATTRS})
        np.savez(os.path.join(directory, "system_%d.npz" % idx), time=time, **arrays)
    if verbose:
        print("Save complete: %s" % directory)


def load_state(simulator, directory="", verbose=False):
    """Every dynamic array of the file goes back into the rod, like upstream's ``load_state`` restores every non-time key
    in place; shapes are checked against the engine's batch first (a ``[batch, ...]`` checkpoint only fits the same
    batch; an unbatched one is broadcast to every instance on purpose -- that is how a batch is seeded from a
    reference checkpoint)."""
    times = []
    for idx, rod in enumerate(simulator):
        path = os.path.join(directory, "system_%d.npz" % idx)
        data = np.load(path)
        times.append(float(data["time"]))
        batch = rod._engine.batch if getattr(rod, "_engine", None) is not None else 1
        for name, (short, shape) in DYNAMIC_FIELDS.items():
            if name not in data.files:
                continue
            value = data[name]
            per_rod = tuple(shape(rod.n_elems))
            if value.shape != per_rod and value.shape != (batch,) + per_rod:
                raise ValueError("%s: %s has shape %s; this rod expects %s or %s (batch %d)"
                                 % (path, name, value.shape, per_rod, (batch,) + per_rod, batch))
            setattr(rod, name, value)
    if any(t != times[0] for t in times):
        raise ValueError("Restart time of loaded rods are different, check your inputs!")
    if verbose:
        print("Load complete: %s" % directory)
    return times[0]
