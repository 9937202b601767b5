"""HBM bandwidth of the terminal-metrics kernel (br2_metrics, csrc/br2_metrics.cuh): one pass over the batch state (mode 1 /
NaN scan) or two (mode 2, with the pre-run copy).  CUDA events on the launching stream, L2 flushed between launches.
usage (GPU box): python profiles/probes/metrics_bandwidth.py [batch]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "br2-simulator_b200")); sys.path.insert(0, ROOT)
import numpy as np
import torch
from br2.environment import Environment

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
env = Environment("probe", batch=B, create_dirs=False)
env.reset(os.path.join(ROOT, "tests/data/rod_library.json"), os.path.join(ROOT, "tests/data/assembly_single.json"), gravity=True)
env.simulator._callback_ops = []
rng = np.random.default_rng(0)
env.run({"action%d" % (i + 1): rng.uniform(0, 40, B) * 6895 for i in range(3)}, duration=200 * 2e-5 - 1e-5, disable_progress_bar=True)
engine = env.engine
state_bytes = engine.state.numel() * 8
# what the kernel reads per instance: x, v, Q, omega (one pass); with the pre-run copy also acceleration and alpha, of both
n_of = [rod.n_elems for rod in env.simulator]
read_one = 8 * B * sum(6 * (n + 1) + 12 * n for n in n_of)
read_two = 2 * 8 * B * sum(9 * (n + 1) + 15 * n for n in n_of)
previous = engine.state.clone()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=engine.device)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6452.8
for label, prev, nbytes in (("NaN scan / steady-state mode 1 (x, v, Q, omega)", None, read_one), ("steady-state mode 2 (six fields of the state and of the pre-run copy)", previous, read_two)):
    times = []
    for _ in range(8):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = engine.metrics(prev)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) * 1e-3)
    best, med = min(times[2:]), sorted(times[2:])[len(times[2:]) // 2]
    gbs = nbytes / med / 1e9
    print("%s: batch %d, %.1f MB read (state block %.1f MB), median %.1f us (best %.1f) -> %.0f GB/s = %.0f %% of the measured HBM peak (%.1f GB/s)"
          % (label, B, nbytes / 1e6, state_bytes / 1e6, med * 1e6, best * 1e6, gbs, 100 * gbs / peak, peak))
assert torch.isfinite(out[:, 1]).all()
env.close()
