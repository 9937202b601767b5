"""Checkpoint helpers (oracle restatement of ``elastica.restart``; TEST INFRASTRUCTURE).

Called through /root/reference/br2/free_simulator.py:93-99: one ``system_{idx}.npz`` per rod
holding ``time`` plus every ndarray attribute; loading copies the arrays back in place,
checks that all files agree on ``time`` and re-applies the constraints."""
import os

import numpy as np


def save_state(simulator, directory="", time=0.0, verbose=False):
    os.makedirs(directory, exist_ok=True)
    for idx, rod in enumerate(simulator):
        path = os.path.join(directory, "system_{}.npz".format(idx))
        arrays = {k: v for k, v in rod.__dict__.items() if isinstance(v, np.ndarray)}
        np.savez(path, time=time, **arrays)
    if verbose:
        print("Save complete: {}".format(directory))


def load_state(simulator, directory="", verbose=False):
    time_list = []
    for idx, rod in enumerate(simulator):
        path = os.path.join(directory, "system_{}.npz".format(idx))
        data = np.load(path, allow_pickle=True)
        for key, value in data.items():
            if key == "time":
                time_list.append(value.item())
                continue
            if value.shape != ():
                getattr(rod, key)[:] = value
            else:
                setattr(rod, key, value)
    if not all(t == time_list[0] for t in time_list):
        raise ValueError("Restart time of loaded rods are different, check your inputs!")
    time = time_list[0]
    simulator.constrain_values(0)
    simulator.constrain_rates(0)
    if verbose:
        print("Load complete: {}".format(directory))
    return time
