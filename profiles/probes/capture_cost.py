"""Cost of the diagnostics capture path: Environment.run with and without the FreeCallback collectors."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "br2-simulator_b200")); sys.path.insert(0, ROOT)
import numpy as np
import torch
from br2.environment import Environment

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rng = np.random.default_rng(0)
act = {"action%d" % (i + 1): rng.uniform(0, 40, B) * 6895 for i in range(3)}
for with_callbacks in (False, True):
    env = Environment("probe", batch=B, create_dirs=False)
    env.reset(os.path.join(ROOT, "tests/data/rod_library.json"), os.path.join(ROOT, "tests/data/assembly_single.json"), gravity=True)
    if not with_callbacks:
        env.simulator._callback_ops = []
    env.run(act, duration=2000 * 2e-5 - 1e-5, disable_progress_bar=True)  # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    env.run(act, duration=10000 * 2e-5 - 1e-5, disable_progress_bar=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("callbacks=%s  batch=%d  10000 steps (5 captures): %.3f s  -> %.3g el-steps/s" % (with_callbacks, B, dt, B * 123 * 10000 / dt))
    env.close()
