"""Throughput of an assembly whose segments differ in element count (41 and 30 elements per rod) on the tile kernel
(classes of rod constants read from shared memory) and on the one-slot-per-thread kernel with per-slot constants.
usage (GPU box): python profiles/probes/ragged_kernels.py [batch] [steps]"""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "br2-simulator_b200")]
import numpy as np
import torch
from br2.environment import Environment

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
data = os.path.join(ROOT, "tests", "data")
library = json.load(open(os.path.join(data, "rod_library.json")))
for key in ("RodBend", "RodRightTwist", "RodLeftTwist"):
    library["Rods"][key + "Short"] = dict(library["Rods"][key], n_elements=30, base_length=0.12)
assembly = json.load(open(os.path.join(data, "assembly_double.json")))
assembly["Segments"]["seg2"]["rod_order"] = ["RodRightTwistShort", "RodBendShort", "RodLeftTwistShort"]
tmp = tempfile.mkdtemp()
rod_db, path = os.path.join(tmp, "lib.json"), os.path.join(tmp, "assembly.json")
json.dump(library, open(rod_db, "w"))
json.dump(assembly, open(path, "w"))
rng = np.random.default_rng(0)
action = {"action%d" % (i + 1): rng.uniform(0, 40, size=batch) * 6895.0 for i in range(3)}
dt = 2.0e-5
for variant, name in ((3, "tile"), (1, "fast")):
    env = Environment("ragged", batch=batch, time_step=dt, create_dirs=False)
    env.reset(rod_db, path, gravity=True)
    env.engine.select_kernel(variant)
    env.simulator._callback_ops = []
    env.run(action, duration=steps * dt - 0.5 * dt, disable_progress_bar=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        env.run(action, duration=steps * dt - 0.5 * dt, disable_progress_bar=True)
    torch.cuda.synchronize()
    elapsed = (time.perf_counter() - t0) / 3
    elements = sum(rod.n_elems for rod in env.simulator)
    print("%s: %.3e rod-element-steps/s (%d instances x %d elements x %d steps in %.1f ms); replays %d" % (
        name, batch * elements * steps / elapsed, batch, elements, steps, 1e3 * elapsed, env.engine.replay_count()))
    env.close()
