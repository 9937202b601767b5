import sys, os
sys.path.insert(0, 'br2-simulator_b200'); sys.path.insert(0, '.')
import numpy as np
from br2.environment import Environment
for asm, steps in (("double", 2000), ("triple", 500)):
    env = Environment("dbg", batch=8, create_dirs=False)
    env.reset("tests/data/rod_library.json", "tests/data/assembly_%s.json" % asm)
    env.simulator._callback_ops = []
    rng = np.random.default_rng(1)
    act = {"action%d" % (i + 1): rng.uniform(0, 40, 8) * 6895 for i in range(3)}
    env.run(act, duration=steps * 2e-5 - 1e-5, disable_progress_bar=True)
    print(asm, env.engine.kernel_info(), "replays", env.engine.replay_count(), env.engine.replay_reasons())
    env.close()
