#!/usr/bin/env python
"""Benchmark of the rod-integration hot path (BASELINE.json metric: rod-element-steps/s).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

One bench "step" = ONE launch of the fused kernel advancing the whole batch `--substeps`
PositionVerlet steps (default 2000 = the reference's capture cadence `step_skip` at dt=2e-5, 25 fps).
Workload at every N: BASELINE.json configs[1], single_br2 assembly (3 rods x 41 elements), 4096 instances
PER GPU with per-instance random actuation pressures, gravity on, dt=2e-5 (weak scaling: the batch is
partitioned, instances never interact, no collective on the step path).

`value`   : element-steps/s with the state resident in HBM, CUDA-event timed per launch, L2 flushed
            between launches, max over ranks.
`e2e`     : same metric through the public API (`Environment.run` with HOST action arrays + a read-back
            of the full batch state into pinned host memory), wall clock between device synchronisations.
`roofline`: fp64 pipe -- 695 algorithmic flop per element-step against a DFMA
            micro-benchmark measured in this process (MEASURED_PEAKS.json has no fp64 entry); the
            `hbm` sub-object gives achieved algorithmic HBM GB/s against MEASURED_PEAKS.json.
`cpu_baseline`: the faithful oracle port (one numba dispatch per operator, like the reference) on one
            host core, on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "br2-simulator_b200"))
os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join(ROOT, ".numba_cache"))

RESULT_OUT = sys.stdout
DATA = os.path.join(ROOT, "tests", "data")
ROD_DB = os.path.join(DATA, "rod_library.json")
PSI = 6895.0
FLOP_PER_ELEMENT_STEP = 695.0  # SURVEY.md Appendix B (single-BR2-type assembly, gravity on)
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the step kernel on the default workload (4096 x
# single_br2, 2000 steps = 8 chunks), ncu capture profiles/r01_dram_2000step_launch.csv: 76.9 MB + 168.3 MB.  The
# algorithmic figure is 231.8 MB (state in, state + derived arrays out): the work queue hands an instance's next chunk
# out while the state the previous one wrote back is still in the L2 (groups of ~2048 instances), so only the last
# chunk's write-back reaches DRAM.  (Chunk-major over the whole batch it was 676 MB.)
NCU_DRAM_BYTES_PER_LAUNCH = {("single", 4096, 2000): 245184768.0}
WORKLOADS = {
    "single": dict(assembly="assembly_single.json", elements=123, gravity=True, seed=0,
                   label="single_br2_v1 x4096/GPU, random pressures U[0,40] psi, gravity on, dt=2e-5 (BASELINE configs[1])"),
    "double": dict(assembly="assembly_double.json", elements=246, gravity=False, seed=1,
                   label="double_br2_v1 with serial joints, random pressures, gravity off, dt=2e-5 (BASELINE configs[2])"),
    "triple": dict(assembly="assembly_triple.json", elements=369, gravity=True, seed=2,
                   label="triple_br2_v1 (41 elements per rod), random pressures, gravity on, dt=2e-5"),
    "triple200": dict(assembly="assembly_triple.json", elements=1800, gravity=True, seed=2, dt=5.0e-6,
                      rod_db="rod_library_n200.json",
                      label="triple_br2_v1 with 200 elements per rod (1809 node slots, cluster of 3 CTAs per instance), "
                            "random pressures, gravity on, dt=5e-6 (BASELINE configs[3])"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="single", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=4096, help="instances per GPU")
    ap.add_argument("--substeps", type=int, default=2000, help="integration steps per launch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--variant", type=int, default=-1, help="force a kernel variant (0 generic, 1/2 fast, 3 tile); -1 = automatic")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)), "MEASURED_PEAKS.json"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """`nvidia-smi -lms 100` in the background, started BEFORE the warm-up (its start-up takes longer than a short timed
    region); only the samples whose timestamps fall inside the timed region are reported."""
    QUERY = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None
        self.begin = self.end = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def mark_begin(self):
        self.begin = time.time()

    def mark_end(self):
        self.end = time.time()

    def stop(self):
        import datetime

        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)  # let the sample that covers the end of the region land
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        try:
            for line in open(self.path):
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 9:
                    continue
                try:
                    stamp = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    rows.append((stamp, float(parts[1]), float(parts[2]), parts[5:9]))
                except ValueError:
                    continue
            os.remove(self.path)
        except Exception:
            pass
        begin, end = self.begin or 0.0, self.end or float("inf")
        inside = [r for r in rows if begin <= r[0] <= end]
        window = "timed region"
        if not inside and rows:  # region shorter than the sampling period: the sample nearest to it (GPU busy since the warm-up)
            mid = 0.5 * (begin + min(end, begin + 3600.0))
            inside = [min(rows, key=lambda r: abs(r[0] - mid))]
            window = "nearest sample to the timed region (%.2f s away)" % abs(inside[0][0] - mid)
        if inside:
            clocks = sorted(r[1] for r in inside)
            reasons = set()
            for r in inside:
                for name, flag in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3]):
                    if flag.lower().startswith("active"):
                        reasons.add(name)
            out.update(sm_mhz=clocks[len(clocks) // 2], sm_max_mhz=inside[0][2], reasons=sorted(reasons),
                       samples=len(inside), window=window)
        return out


def dist_setup(args):
    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries the one JSON line only
        if args.impl == "ours" and torch.cuda.is_available():
            torch.cuda.set_device(local)
            dist.init_process_group(backend="nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    return world, rank, local


# ----------------------------------------------------------------------------------------------
# CPU legs (the only place bench.py touches oracle/)
# ----------------------------------------------------------------------------------------------
def _faithful_worker(payload):
    assembly, gravity, pressures, n_steps = payload
    from oracle import br2_port

    env = br2_port.OracleEnvironment()
    env.reset(ROD_DB, os.path.join(DATA, assembly), gravity=gravity)
    action = {"action%d" % (i + 1): float(p) for i, p in enumerate(pressures)}
    env.simulator._callback_list[:] = []
    env.run(action, n_steps=20)  # JIT / cache warm-up, untimed
    t0 = time.perf_counter()
    env.run(action, n_steps=n_steps)
    return time.perf_counter() - t0


def cpu_baseline_port(workload, n_steps=8000):
    """Faithful oracle port, ONE core, one instance: the reference's execution pattern
    (one Python->numba dispatch per operator / per joint element per step)."""
    w = WORKLOADS[workload]
    seconds = _faithful_worker((w["assembly"], w["gravity"], [30 * PSI, 30 * PSI, 0.0], n_steps))
    return {"value": w["elements"] * n_steps / seconds, "unit": "rod-element-steps/s", "cores": 1, "kind": "port",
            "sample": "1 instance x %d steps of the same assembly (faithful per-operator numba oracle; "
                      "the reference's PyElastica cannot be installed here)" % n_steps}


def run_reference_arm(args, world, rank):
    """`--impl reference`: the reference's CPU path (oracle port; PyElastica itself is not
    installable) on all host cores -- one independent instance per process, like a user of the
    reference running a sweep would."""
    if rank != 0:
        return
    import multiprocessing as mp

    w = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    sample_steps = 400
    import numpy as np

    rng = np.random.default_rng(w["seed"])
    payloads = [(w["assembly"], w["gravity"], list(rng.uniform(0, 40, size=3) * PSI), sample_steps) for _ in range(cores)]
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        for _ in range(max(args.warmup, 1)):
            pool.map(_faithful_worker, payloads)
        elapsed = 0.0  # stepping time only (slowest process per bench step); set-up / JIT is not charged
        for _ in range(args.steps):
            elapsed += max(pool.map(_faithful_worker, payloads))
    value = cores * w["elements"] * sample_steps * args.steps / elapsed
    line = {
        "impl": "reference", "metric": "rod-element-steps/sec", "value": value, "unit": "rod-element-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["label"], "batch_per_gpu": args.batch, "substeps_per_step": args.substeps},
        "cpu_baseline": {"value": value, "unit": "rod-element-steps/s", "cores": cores, "kind": "port",
                         "sample": "%d processes x 1 instance x %d steps per bench step (+20 untimed warm-up steps each); "
                                   "faithful per-operator numba oracle -- PyElastica 0.3.1 is not installable here" % (cores, sample_steps)},
        "e2e": {"value": value, "unit": "rod-element-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=RESULT_OUT, flush=True)


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def run_ours(args, world, rank, local):
    import numpy as np
    import torch

    from br2 import _native
    from br2.environment import Environment

    w = WORKLOADS[args.workload]
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    B, M, dt = args.batch, args.substeps, w.get("dt", 2.0e-5)
    rod_db = os.path.join(DATA, w.get("rod_db", "rod_library.json"))
    rng = np.random.default_rng(w["seed"] + 1000 * rank)
    pressures = rng.uniform(0.0, 40.0, size=(3, B)) * PSI

    # Environment's batch is the GLOBAL instance count; under torchrun it keeps this rank's contiguous share (B)
    env = Environment("bench", time_step=dt, batch=B * world, device=device, create_dirs=False)
    assert env.batch == B
    env.reset(rod_db, os.path.join(DATA, w["assembly"]), gravity=w["gravity"])
    env.simulator._callback_ops = []  # stepping only; capture is a separate (next-row) path
    engine = env.engine
    if args.variant >= 0:
        engine.select_kernel(args.variant)
    names = ["action1", "action2", "action3"]
    action_host = {n: torch.from_numpy(pressures[i].copy()).pin_memory().numpy() for i, n in enumerate(names)}
    env.assy.set_actuation(action_host)
    env._push_actuation()

    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=device)  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize(device)
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(device)

    # ---- device-resident throughput ------------------------------------------------------
    t_sim = 0.0
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        t_sim = engine.step(M, t_sim, dt)
    barrier()
    sampler.mark_begin()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    launches0 = engine.launches
    for i in range(args.steps):
        flush.fill_(float(i))  # L2 flush between timed launches (outside the event pair)
        starts[i].record()
        t_sim = engine.step(M, t_sim, dt)
        stops[i].record()
    barrier()
    sampler.mark_end()
    clocks = sampler.stop()
    launches = engine.launches - launches0
    per_launch_ms = [s.elapsed_time(e) for s, e in zip(starts, stops)]
    total_ms = torch.tensor([sum(per_launch_ms)], dtype=torch.float64, device=device)
    if world > 1:
        torch.distributed.all_reduce(total_ms, op=torch.distributed.ReduceOp.MAX)
    total_s = float(total_ms.item()) * 1e-3
    element_steps = world * B * w["elements"] * M * args.steps
    value = element_steps / total_s
    state_ok = bool(torch.isfinite(engine.state).all().item())

    # ---- end to end through the public API ---------------------------------------------------
    # Every step: H2D of the step's pressures (inside Environment.run), the integration launch, and a D2H read of the
    # whole batch state into pinned host memory.  The read is pipelined the way a user consuming results would do it:
    # a device-side clone of the state (0.1 ms) is copied out on a second stream while the next step's launch runs;
    # all copies complete inside the timed region (barrier() synchronises every stream of the device).
    pinned = [torch.empty((B, engine.words), dtype=torch.float64).pin_memory() for _ in range(2)]
    clones = [torch.empty_like(engine.state) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device)
    copied = [None, None]
    e2e_steps = max(2, args.steps)
    duration = M * dt - 0.5 * dt
    e2e_count = [0]

    from br2 import _dist

    def e2e_once():
        buf = e2e_count[0] % 2
        e2e_count[0] += 1
        env.run(action_host, duration=duration, disable_progress_bar=True)  # H2D of the pressures inside
        if copied[buf] is not None:
            copied[buf].synchronize()                                       # this buffer's previous read has landed
        clones[buf].copy_(engine.state)
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ready)
            pinned[buf].copy_(clones[buf], non_blocking=True)               # D2H of the batch state
            copied[buf] = torch.cuda.Event()
            copied[buf].record(copy_stream)
        if world > 1:
            # the only collective of the path: per-instance metrics of all ranks to rank 0 over NCCL
            tip = engine.view(0, "x")[:, :, -1].contiguous()
            _dist.gather_batch(tip, B * world, dst=0, device=device)

    e2e_once()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_once()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    if world > 1:
        torch.distributed.all_reduce(e2e_s, op=torch.distributed.ReduceOp.MAX)
    e2e_value = world * B * w["elements"] * M * e2e_steps / float(e2e_s.item())

    if rank != 0:
        return
    # ---- roofline denominators ---------------------------------------------------------------
    pk, pk_src = peaks()
    import ctypes

    tf = ctypes.c_double(0.0)
    _native.check(engine.lib.br2_fp64_peak_tflops(local, 3, ctypes.byref(tf)))
    kernel_s = sum(per_launch_ms) * 1e-3 / args.steps          # this rank's average launch duration
    el_steps_per_launch = B * w["elements"] * M
    achieved_tf = FLOP_PER_ELEMENT_STEP * el_steps_per_launch / kernel_s / 1e12
    io_words = sum(sum(sz for short, (off, sz) in fields.items() if short in ("x", "v", "Q", "w")) for fields in engine.program.layout)
    hbm_bytes = B * 8.0 * (io_words + engine.words)            # state read + state/derived write per launch
    info = engine.kernel_info()
    line = {
        "metric": "rod-element-steps/sec", "value": value, "unit": "rod-element-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_s / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["label"], "batch_per_gpu": B, "substeps_per_step": M, "dt": dt,
                   "elements_per_instance": w["elements"], "l2": "flushed between timed launches (256 MB fill)",
                   "parallelism": "batch partition x%d, no step-path collective" % world},
        "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": tf.value, "unit": "TFLOP/s",
                     "frac": achieved_tf / tf.value if tf.value else None,
                     "traffic": NCU_DRAM_BYTES_PER_LAUNCH.get((args.workload, B, M)),
                     "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/r01_dram_2000step_launch.csv)",
                     "flop_per_element_step": FLOP_PER_ELEMENT_STEP,
                     "peak_source": "DFMA micro-benchmark run in this process (br2_fp64_peak_tflops)",
                     "hbm": {"achieved_gbs": hbm_bytes / kernel_s / 1e9, "peak_gbs": pk["hbm_gbs"], "peak_source": pk_src,
                             "frac": hbm_bytes / kernel_s / 1e9 / pk["hbm_gbs"], "bytes_per_launch": hbm_bytes}},
        "e2e": {"value": e2e_value, "unit": "rod-element-steps/s", "h2d_bytes_per_step": int(3 * B * 8),
                "d2h_bytes_per_step": int(B * engine.words * 8), "steps": e2e_steps},
        "gpu_launches": launches,
        "kernel": dict(info, name=("br2_step_generic_kernel", "br2_step_fast_kernel", "br2_step_fast_kernel", "br2_step_tile_kernel")[info["variant"]],
                       launch_ms=per_launch_ms, exact_path_replays=engine.replay_count(),
                       launches_per_step="optimistic kernel + exact-path replay pass (exits at once unless flagged)"
                       if info["variant"] else "generic kernel"),
        "clocks": clocks,
        "state_finite": state_ok,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_port(args.workload)
    print(json.dumps(line), file=RESULT_OUT, flush=True)


def main():
    # stdout carries the one JSON line and nothing else: libraries that write to file descriptor 1 (NCCL prints its
    # version there when the box sets NCCL_DEBUG) are pointed at stderr for the whole run
    global RESULT_OUT
    RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    args = parse_args()
    world, rank, local = dist_setup(args)
    if args.impl == "reference":
        run_reference_arm(args, world, rank)
    else:
        run_ours(args, world, rank, local)
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
