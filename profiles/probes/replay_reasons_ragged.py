import sys, os, json, tempfile
sys.path.insert(0, 'br2-simulator_b200'); sys.path.insert(0, '.')
import numpy as np
from br2.environment import Environment
data = "tests/data"
library = json.load(open(os.path.join(data, "rod_library.json")))
for key in ("RodBend", "RodRightTwist", "RodLeftTwist"):
    library["Rods"][key + "Short"] = dict(library["Rods"][key], n_elements=30, base_length=0.12)
assembly = json.load(open(os.path.join(data, "assembly_double.json")))
assembly["Segments"]["seg2"]["rod_order"] = ["RodRightTwistShort", "RodBendShort", "RodLeftTwistShort"]
tmp = tempfile.mkdtemp()
rod_db, path = os.path.join(tmp, "lib.json"), os.path.join(tmp, "assembly.json")
json.dump(library, open(rod_db, "w")); json.dump(assembly, open(path, "w"))
for name, db, asm in (("ragged", rod_db, path), ("double", "tests/data/rod_library.json", "tests/data/assembly_double.json")):
    for variant in (1, 2):
        env = Environment("dbg", batch=8, create_dirs=False)
        env.reset(db, asm, gravity=True)
        try:
            env.engine.select_kernel(variant)
        except Exception as e:
            print(name, variant, "n/a", str(e)[:60]); env.close(); continue
        env.simulator._callback_ops = []
        rng = np.random.default_rng(1)
        act = {"action%d" % (i + 1): rng.uniform(0, 40, 8) * 6895 for i in range(3)}
        env.run(act, duration=2000 * 2e-5 - 1e-5, disable_progress_bar=True)
        print(name, "variant", variant, "replays", env.engine.replay_count(), env.engine.replay_reasons())
        env.close()
