#!/usr/bin/env python
"""Benchmark of the rod-integration hot path (BASELINE.json metric: rod-element-steps/s).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

One bench "step" = ONE launch of the fused kernel advancing this GPU's share of the batch `--substeps`
PositionVerlet steps (default 2000 = the reference's capture cadence `step_skip` at dt=2e-5, 25 fps).
Workload: on ONE GPU BASELINE.json configs[1] (single_br2, 3 rods x 41 elements, 4096 instances with per-instance
random actuation pressures, gravity on, dt=2e-5); on N > 1 GPUs configs[4] as written (65536 instances IN TOTAL,
action1 swept linearly over them, split contiguously over the ranks: strong scaling; instances never interact,
so there is no collective on the step path).  `extra_workloads` carries short runs of the other BASELINE
configurations (configs[2] double x16384, configs[3] triple @ 200 elements x8192, and at N = 1 the whole
configs[4] sweep) so that the driver's record holds them too; `gather` (N > 1) times the one real collective,
a capture block gathered onto rank 0's GPU.

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
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the step kernel, from an `ncu --metrics dram__bytes_*`
# capture of exactly this command; keyed by (workload, instances on the GPU, steps per launch) and by the capture file, so
# a line for any other workload carries null instead of a stale figure.  (Algorithmic bytes are computed live: `hbm`.)
NCU_DRAM_BYTES_PER_LAUNCH = {
    ("single", 4096, 2000): (223571200.0, "profiles/r02_dram_2000step_launch.csv (ncu capture of this command on the final round-2 kernel: 76.2 MB read + 147.3 MB written; part of the last write-back is still in the L2 when the kernel ends)"),
}
WORKLOADS = {
    # BASELINE.json configs[1]
    "single": dict(assembly="assembly_single.json", elements=123, gravity=True, seed=0, batch=4096, config="configs[1]",
                   what="single_br2_v1, per-instance random pressures U[0,40] psi on action1-3, gravity on, dt=2e-5"),
    # BASELINE.json configs[4]: the batch is the TOTAL over all GPUs, split contiguously
    "sweep": dict(assembly="assembly_single.json", elements=123, gravity=True, seed=3, batch=65536, total=True, sweep=True,
                  config="configs[4]",
                  what="single_br2_v1 actuation sweep: action1 swept linearly 0->40 psi over the instances, action2/3 "
                       "random U[0,40] psi, gravity on, dt=2e-5, contiguous batch split over the GPUs"),
    # BASELINE.json configs[2]
    "double": dict(assembly="assembly_double.json", elements=246, gravity=False, seed=1, batch=16384, total=True,
                   config="configs[2]",
                   what="double_br2_v1 (surface connections + 9 tip-to-base joints), random pressures, gravity off, dt=2e-5"),
    "triple": dict(assembly="assembly_triple.json", elements=369, gravity=True, seed=2, batch=4096, config="(sample triple, n=41)",
                   what="triple_br2_v1 (41 elements per rod), random pressures, gravity on, dt=2e-5"),
    # BASELINE.json configs[3]
    "triple200": dict(assembly="assembly_triple.json", elements=1800, gravity=True, seed=2, dt=5.0e-6, batch=8192, total=True,
                      rod_db="rod_library_n200.json", config="configs[3]",
                      what="triple_br2_v1 with 200 elements per rod (1809 node slots per instance), random pressures, "
                           "gravity on, dt=5e-6"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="default: `single` (BASELINE configs[1]) on one GPU, `sweep` (configs[4], 65536 instances split "
                         "over the GPUs) on several")
    ap.add_argument("--batch", type=int, default=None, help="instances (per GPU for `single` / `triple`; in total for the others)")
    ap.add_argument("--substeps", type=int, default=2000, help="integration steps per launch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra_workloads (configs[2], [3], [4]) lines")
    ap.add_argument("--n-elements", type=int, default=None, help="experiment: override the rod library's elements per rod")
    ap.add_argument("--dt", type=float, default=None, help="experiment: override the time step")
    ap.add_argument("--variant", type=int, default=-1, help="force a kernel variant (0 generic, 1/2 fast, 3 tile); -1 = automatic")
    return ap.parse_args()


def resolve_workload(args, world):
    """(name, spec, instances on each GPU, instances in total): what both arms run and describe."""
    name = args.workload or ("single" if world == 1 else "sweep")
    w = dict(WORKLOADS[name])
    if args.n_elements:  # experiments only (kernel studies): a rod library with another resolution
        import tempfile as _tmp

        library = json.load(open(os.path.join(DATA, w.get("rod_db", "rod_library.json"))))
        library["DefaultParams"]["n_elements"] = args.n_elements
        path = os.path.join(_tmp.mkdtemp(prefix="br2_bench_"), "rod_library.json")
        json.dump(library, open(path, "w"))
        rods = len(json.load(open(os.path.join(DATA, w["assembly"])))["Segments"]) * 3
        w.update(rod_db=path, elements=rods * args.n_elements, config="experiment: %d elements per rod" % args.n_elements)
    if args.dt:
        w["dt"] = args.dt
    batch = args.batch or w["batch"]
    if w.get("total"):
        total = batch
        if total < world:
            raise SystemExit("batch %d is smaller than the number of GPUs" % total)
    else:
        total = batch * world
    return name, w, total


def describe_config(name, w, total, world, substeps):
    """The `config` object of the JSON line -- identical in both arms (the driver compares them)."""
    per_gpu = "%d" % (total // world) if total % world == 0 else "%d or %d" % (total // world, total // world + 1)
    return {
        "workload": "%s x%d%s (BASELINE %s)" % (w["what"], total, "" if world == 1 else " in total, %s per GPU" % per_gpu, w["config"]),
        "name": name, "instances_total": total, "instances_per_gpu": per_gpu, "substeps_per_step": substeps,
        "dt": w.get("dt", 2.0e-5), "elements_per_instance": w["elements"],
        "l2": "flushed between timed launches (256 MB fill)",
        "parallelism": "contiguous batch partition x%d, no step-path collective" % world,
        "scaling_note": "fixed total batch split over the GPUs" if w.get("total") else "fixed batch per GPU",
    }


def global_pressures(w, total):
    """[3, total] pressures of the whole job (every rank draws the same stream and keeps its slice)."""
    import numpy as np

    rng = np.random.default_rng(w["seed"])
    pressures = rng.uniform(0.0, 40.0, size=(3, total)) * PSI
    if w.get("sweep"):
        pressures[0] = np.linspace(0.0, 40.0, total) * PSI
    return pressures


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
    assembly, rod_db, gravity, dt, pressures, n_steps = payload
    from oracle import br2_port

    env = br2_port.OracleEnvironment(time_step=dt)
    env.reset(os.path.join(DATA, rod_db), os.path.join(DATA, assembly), gravity=gravity)
    action = {"action%d" % (i + 1): float(p) for i, p in enumerate(pressures)}
    env.simulator._callback_list[:] = []
    env.run(action, n_steps=20)  # JIT / cache warm-up, untimed
    t0 = time.perf_counter()
    env.run(action, n_steps=n_steps)
    return time.perf_counter() - t0


def cpu_baseline_port(w, n_steps=8000):
    """Faithful oracle port, ONE core, one instance: the reference's execution pattern
    (one Python->numba dispatch per operator / per joint element per step)."""
    seconds = _faithful_worker((w["assembly"], w.get("rod_db", "rod_library.json"), w["gravity"], w.get("dt", 2.0e-5),
                                [30 * PSI, 30 * PSI, 0.0], n_steps))
    return {"value": w["elements"] * n_steps / seconds, "unit": "rod-element-steps/s", "cores": 1, "kind": "port",
            "sample": "1 instance x %d steps of the same assembly (faithful per-operator numba oracle; "
                      "the reference's PyElastica cannot be installed here)" % n_steps}


def run_reference_arm(args, world, rank):
    """`--impl reference`: the reference's CPU path (oracle port; PyElastica itself is not
    installable) on all host cores -- one independent instance per process, like a user of the
    reference running a sweep would.  A bounded RATE sample of the workload `config` names."""
    if rank != 0:
        return
    import multiprocessing as mp

    name, w, total = resolve_workload(args, world)
    cores = os.cpu_count() or 1
    sample_steps = 400
    pressures = global_pressures(w, total)
    picks = [int(i) for i in (range(cores) if total <= cores else
                              __import__("numpy").linspace(0, total - 1, cores).round())]
    payloads = [(w["assembly"], w.get("rod_db", "rod_library.json"), w["gravity"], w.get("dt", 2.0e-5),
                 list(pressures[:, i % total]), sample_steps) for i in picks]
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
        "higher_is_better": True, "scaling": "strong" if w.get("total") else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": describe_config(name, w, total, world, args.substeps),
        "cpu_baseline": {"value": value, "unit": "rod-element-steps/s", "cores": cores, "kind": "port",
                         "sample": "%d processes x 1 instance (spread over the batch's pressures) x %d steps per bench step "
                                   "(+20 untimed warm-up steps each); faithful per-operator numba oracle -- PyElastica 0.3.1 "
                                   "is not installable here" % (cores, sample_steps)},
        "e2e": {"value": value, "unit": "rod-element-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=RESULT_OUT, flush=True)


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
class Job:
    """One workload set up on this rank's GPU through the public API (`Environment`), ready to be stepped."""

    def __init__(self, name, w, total, world, rank, device, variant=-1):
        import torch

        from br2.environment import Environment

        self.name, self.w, self.total, self.world = name, w, total, world
        self.dt = w.get("dt", 2.0e-5)
        rod_db = os.path.join(DATA, w.get("rod_db", "rod_library.json"))
        # Environment's batch is the GLOBAL instance count; under torchrun it keeps this rank's contiguous share
        self.env = Environment("bench", time_step=self.dt, batch=total, device=device, create_dirs=False)
        self.env.reset(rod_db, os.path.join(DATA, w["assembly"]), gravity=w["gravity"])
        self.env.simulator._callback_ops = []  # stepping only; the capture path is measured separately (`gather`)
        self.engine = self.env.engine
        if variant >= 0:
            self.engine.select_kernel(variant)
        start, stop = self.env.batch_range
        self.local = stop - start
        pressures = global_pressures(w, total)[:, start:stop]
        self.action_host = {"action%d" % (i + 1): torch.from_numpy(pressures[i].copy()).pin_memory().numpy() for i in range(3)}
        self.env.assy.set_actuation(self.action_host)
        self.env._push_actuation()
        self.t_sim = 0.0

    def step(self, substeps):
        self.t_sim = self.engine.step(substeps, self.t_sim, self.dt)

    def close(self):
        self.env.close()


def time_device_resident(job, substeps, steps, warmup, flush, barrier, world, sampler=None):
    """W untimed + K timed launches, CUDA events around each launch (on torch's current stream, which is the stream
    br2_step launches on), L2 flushed between launches outside the event pairs; max over ranks."""
    import torch

    device = job.engine.device
    for _ in range(warmup):
        job.step(substeps)
    barrier()
    if sampler is not None:
        sampler.mark_begin()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    launches0 = job.engine.launches
    for i in range(steps):
        flush.fill_(float(i))
        starts[i].record()
        job.step(substeps)
        stops[i].record()
    barrier()
    if sampler is not None:
        sampler.mark_end()
    per_launch_ms = [s.elapsed_time(e) for s, e in zip(starts, stops)]
    total_ms = torch.tensor([sum(per_launch_ms)], dtype=torch.float64, device=device)
    if world > 1:
        torch.distributed.all_reduce(total_ms, op=torch.distributed.ReduceOp.MAX)
    total_s = float(total_ms.item()) * 1e-3
    return {"total_s": total_s, "per_launch_ms": per_launch_ms, "launches": job.engine.launches - launches0,
            "value": job.total * job.w["elements"] * substeps * steps / total_s}


def roofline_of(job, substeps, per_launch_ms, peak_tf, pk, pk_src):
    kernel_s = sum(per_launch_ms) * 1e-3 / len(per_launch_ms)  # this rank's average launch duration
    engine = job.engine
    achieved_tf = FLOP_PER_ELEMENT_STEP * job.local * job.w["elements"] * substeps / kernel_s / 1e12
    io_words = sum(sum(sz for short, (off, sz) in fields.items() if short in ("x", "v", "Q", "w")) for fields in engine.program.layout)
    hbm_bytes = job.local * 8.0 * (io_words + engine.words)  # state read + state/derived write per launch
    traffic = NCU_DRAM_BYTES_PER_LAUNCH.get((job.name, job.local, substeps))
    return {"bound": "fp64", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": achieved_tf / peak_tf if peak_tf else None,
            "traffic": traffic[0] if traffic else None,
            "traffic_source": traffic[1] if traffic else "no ncu DRAM capture of this exact workload",
            "flop_per_element_step": FLOP_PER_ELEMENT_STEP,
            "peak_source": "DFMA micro-benchmark run in this process (br2_fp64_peak_tflops); MEASURED_PEAKS.json has no fp64 entry",
            "hbm": {"achieved_gbs": hbm_bytes / kernel_s / 1e9, "peak_gbs": pk["hbm_gbs"], "peak_source": pk_src,
                    "frac": hbm_bytes / kernel_s / 1e9 / pk["hbm_gbs"], "bytes_per_launch": hbm_bytes}}


KERNEL_NAMES = ("br2_step_generic_kernel", "br2_step_fast_kernel", "br2_step_fast_kernel", "br2_step_tile_kernel")


def kernel_of(job, per_launch_ms):
    info = job.engine.kernel_info()
    return dict(info, name=KERNEL_NAMES[info["variant"]], plan=job.env.simulator.kernel_plan(), launch_ms=per_launch_ms,
                exact_path_replays=job.engine.replay_count(),
                launches_per_step="optimistic kernel + exact-path replay pass (exits at once unless flagged)"
                if info["variant"] else "generic kernel")


def measure_gather(job, world, rank, barrier, repeats=3):
    """One capture of this workload (a device clone of the batch state, what FreeCallback records:
    /root/reference/br2/free_simulator.py:56-72) gathered from every rank onto rank 0's GPU over NCCL, device to device."""
    import torch

    from br2 import _dist

    engine = job.engine
    block = engine.state.clone()
    _dist.gather_block(block, job.total, dst=0)  # warm-up (NCCL connection set-up)
    barrier()
    times = []
    for _ in range(repeats):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        out = _dist.gather_block(block, job.total, dst=0)
        e1.record()
        barrier()
        times.append(e0.elapsed_time(e1) * 1e-3)
        del out
    t = torch.tensor([min(times)], dtype=torch.float64, device=engine.device)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    seconds = float(t.item())
    received = (job.total - job.local) * engine.words * 8.0 if rank == 0 else 0.0
    return {"what": "one capture (batch state block [instances, %d words]) gathered to rank 0's GPU, NCCL send/recv, "
                    "device to device, best of %d" % (engine.words, repeats),
            "bytes_received_by_rank0": received, "seconds": seconds, "gbs": received / seconds / 1e9 if seconds else None,
            "nvlink_peak_gbs_per_direction": 900.0}


def run_ours(args, world, rank, local):
    import ctypes

    import torch

    from br2 import _native

    name, w, total = resolve_workload(args, world)
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    M = args.substeps
    job = Job(name, w, total, world, rank, device, args.variant)
    engine, env, B = job.engine, job.env, job.local

    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=device)  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize(device)
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(device)

    # ---- device-resident throughput ------------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    res = time_device_resident(job, M, args.steps, args.warmup, flush, barrier, world, sampler)
    clocks = sampler.stop()
    state_ok = bool(torch.isfinite(engine.state).all().item())

    # ---- end to end through the public API ---------------------------------------------------
    # Every step: H2D of the step's pressures (inside Environment.run), the integration launch, and a D2H read of the
    # whole batch state into pinned host memory (a superset of what the reference's FreeCallback captures every
    # `step_skip` = 2000 steps).  The read is pipelined the way a user consuming results would do it: a device-side clone
    # of the state is copied out on a second stream while the next step's launch runs; all copies complete inside the
    # timed region (barrier() synchronises every stream of the device).
    pinned = [torch.empty((B, engine.words), dtype=torch.float64).pin_memory() for _ in range(2)]
    clones = [torch.empty_like(engine.state) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device)
    copied = [None, None]
    e2e_steps = max(2, args.steps)
    duration = M * job.dt - 0.5 * job.dt
    e2e_count = [0]

    from br2 import _dist

    def e2e_once():
        buf = e2e_count[0] % 2
        e2e_count[0] += 1
        env.run(job.action_host, duration=duration, disable_progress_bar=True)  # H2D of the pressures inside
        if copied[buf] is not None:
            copied[buf].synchronize()                                           # this buffer's previous read has landed
        clones[buf].copy_(engine.state)
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ready)
            pinned[buf].copy_(clones[buf], non_blocking=True)                   # D2H of the batch state
            copied[buf] = torch.cuda.Event()
            copied[buf].record(copy_stream)
        if world > 1:
            # the only collective of the path: per-instance metrics of all ranks to rank 0 over NCCL
            tip = engine.view(0, "x")[:, :, -1].contiguous()
            _dist.gather_batch(tip, total, dst=0, device=device)

    e2e_once()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_once()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    if world > 1:
        torch.distributed.all_reduce(e2e_s, op=torch.distributed.ReduceOp.MAX)
    e2e_value = total * w["elements"] * M * e2e_steps / float(e2e_s.item())
    del pinned, clones

    gather = measure_gather(job, world, rank, barrier) if world > 1 else None

    # ---- roofline denominators ---------------------------------------------------------------
    pk, pk_src = peaks()
    tf = ctypes.c_double(0.0)
    _native.check(engine.lib.br2_fp64_peak_tflops(local, 3, ctypes.byref(tf)))
    line = {
        "metric": "rod-element-steps/sec", "value": res["value"], "unit": "rod-element-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * res["total_s"] / args.steps,
        "higher_is_better": True, "scaling": "strong" if w.get("total") else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": describe_config(name, w, total, world, M),
        "roofline": roofline_of(job, M, res["per_launch_ms"], tf.value, pk, pk_src),
        "e2e": {"value": e2e_value, "unit": "rod-element-steps/s", "h2d_bytes_per_step": int(3 * B * 8),
                "d2h_bytes_per_step": int(B * engine.words * 8), "steps": e2e_steps,
                "readback": "whole batch state (positions, velocities, directors, omegas and the derived callback arrays) "
                            "into pinned host memory every step; diagnostics collectors off"},
        "gpu_launches": res["launches"],
        "kernel": kernel_of(job, res["per_launch_ms"]),
        "clocks": clocks,
        "state_finite": state_ok,
    }
    if gather is not None:
        line["gather"] = gather
    job.close()
    del job, engine, env
    torch.cuda.empty_cache()

    # ---- the other BASELINE configurations, short runs (device-resident value + roofline only) ----------------
    extras = []
    if not args.no_extras and args.workload is None and args.batch is None:
        plan = [("double", 2000, 3, 2), ("triple200", 1000, 3, 2)]
        if world == 1:
            plan.insert(0, ("sweep", 2000, 3, 2))  # the N = 1 point of the configs[4] series
        for xname, xsub, xsteps, xwarm in plan:
            xw = WORKLOADS[xname]
            xtotal = xw["batch"]
            try:
                xjob = Job(xname, xw, xtotal, world, rank, device)
                xres = time_device_resident(xjob, xsub, xsteps, xwarm, flush, barrier, world)
                entry = {"config": describe_config(xname, xw, xtotal, world, xsub), "value": xres["value"],
                         "unit": "rod-element-steps/s", "steps": xsteps, "warmup": xwarm,
                         "ms_per_step": 1e3 * xres["total_s"] / xsteps, "scaling": "strong",
                         "roofline": roofline_of(xjob, xsub, xres["per_launch_ms"], tf.value, pk, pk_src),
                         "kernel": kernel_of(xjob, xres["per_launch_ms"]), "gpu_launches": xres["launches"],
                         "state_finite": bool(torch.isfinite(xjob.engine.state).all().item())}
                xjob.close()
                del xjob
                torch.cuda.empty_cache()
            except Exception as exc:  # an extra line must never take the headline down with it
                entry = {"config": {"name": xname}, "error": "%s: %s" % (type(exc).__name__, exc)}
            extras.append(entry)
    if rank != 0:
        return
    if extras:
        line["extra_workloads"] = extras
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_port(w)
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
