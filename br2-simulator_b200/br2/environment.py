"""User-facing environment, drop-in for /root/reference/br2/environment.py.

Same constructor, ``reset`` / ``run`` / ``save_state`` / ``load_state`` / ``save_data`` / ``render_video`` /
``close`` and the same attributes (``assy``, ``simulator``, ``shearable_rods``, ``data_rods``, ``time``,
``time_step``, ``step_skip``, thresholds).  Extensions: ``batch`` (number of independent assembly
instances integrated side by side on the GPU) and ``device``; with ``batch > 1`` an ``action`` value
may be a length-``batch`` array and per-instance results are returned next to the reference's scalars.

The reference's hot loop (``while time < self.time + duration: time = self.do_step(...)``,
:278-286) is replaced by a handful of kernel launches: the number of steps is fixed up front by the
same fp64 accumulation the loop performs, and launches are split only where a diagnostics callback
has to fire (every ``step_skip`` steps, ``current_step = round(time / dt)``).
"""
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _dist, _native
from .free_simulator import FreeAssembly
from .hooks import PositionVerlet, extend_stepper_interface


@dataclass
class DataPaths:
    """Result directories of one run tag: ``result_<tag>/{simulation_saves,renderings,data}``."""

    tag: str

    @property
    def paths(self) -> str:
        return "result_%s" % self.tag

    @property
    def simulation(self) -> str:
        return os.path.join(self.paths, "simulation_saves")

    @property
    def renderings(self) -> str:
        return os.path.join(self.paths, "renderings")

    @property
    def data(self) -> str:
        return os.path.join(self.paths, "data")

    def initialize(self):
        for directory in (self.paths, self.simulation, self.renderings, self.data):
            os.makedirs(directory, exist_ok=True)


@dataclass
class TerminalInfo:
    """Outcome of ``Environment.run``.  ``<x>_nan_status`` / ``<x>_steady_state_status`` attributes are
    attached only when the corresponding check was requested, as in the reference (:72-131)."""

    end_status: bool = False

    def _status_names(self):
        return [name for name in self.__dir__() if name.endswith("status")]

    def __str__(self):
        return "\n".join("%s = %s" % (name, getattr(self, name)) for name in self._status_names())

    @property
    def combined_nan_status(self) -> bool:
        # reference semantics (:113): all() over the individual NaN flags
        return all(getattr(self, n) for n in self._status_names() if "_nan_" in n and "combined_" not in n)

    @property
    def combined_steady_state_status(self) -> bool:
        # reference semantics (:131): any() over the individual steady-state flags
        return any(getattr(self, n) for n in self._status_names() if "_steady_state_" in n and "combined_" not in n)


def _capture_plan(collection, start_time, n_steps, dt):
    """Ordinals (1-based, within this run) of the steps after which some callback fires, using the
    reference's own step index ``round(time / dt)`` on the fp64-accumulated time (one C loop: br2_capture_plan)."""
    if not collection._callback_ops:
        return []
    return _native.capture_plan(start_time, n_steps, dt, [op.capture_steps() for _, op in collection._callback_ops])


class Environment:
    def __init__(self, run_tag: str, rendering_fps: int = 25, time_step: float = 2.0e-5,
                 final_time: Optional[float] = None, *, batch: int = 1, device=None, create_dirs: bool = True):
        self.StatefulStepper = PositionVerlet()
        self.paths = DataPaths(run_tag)
        if create_dirs:
            self.paths.initialize()
        self.final_time = final_time
        self.time_step = time_step
        self.rendering_fps = rendering_fps
        self.step_skip = int(1.0 / (self.rendering_fps * self.time_step))
        self.capture_interval = None
        self.shearable_rods = {}
        # ``batch`` is the GLOBAL number of instances.  Under torch.distributed (one process per GPU) the
        # batch is split contiguously over the ranks (br2._dist.partition); instances never interact, so
        # stepping needs no collective and only results are gathered (gather_data / TerminalInfo).
        self.global_batch = int(batch)
        self.world_size, self.rank = 1, 0
        try:
            import torch.distributed as dist

            if dist.is_available() and dist.is_initialized():
                self.world_size, self.rank = dist.get_world_size(), dist.get_rank()
        except ImportError:
            pass
        start, stop = _dist.partition(self.global_batch, self.world_size, self.rank)
        self.batch_range = (start, stop)
        self.batch = stop - start
        if self.batch < 1:
            raise ValueError("batch %d is smaller than the number of ranks %d" % (self.global_batch, self.world_size))
        self.device = device
        # steady-state thresholds (reference :176-181)
        self.position_threshold = 1.0e-7
        self.director_threshold = 1.0e-5
        self.velocity_threshold = 1.0e-4
        self.omega_threshold = 1.0e-3
        self.acceleration_threshold = 10 ** (0)
        self.alpha_threshold = 1.0e1

    # ------------------------------------------------------------------ reset (reference :183-224)
    def reset(self, rod_database_path: str, assembly_config_path: str, start_time: float = 0.0, **kwargs) -> None:
        assert os.path.exists(rod_database_path), "Rod database path does not exists."
        assert os.path.exists(assembly_config_path), "Assembly configuration does not exists."
        verbose = kwargs.pop("verbose", False)
        # extension: material parameters that differ between the instances of the batch (br2/_sweep.py), e.g.
        # per_instance={"youngs_modulus": np.linspace(0.8e7, 1.2e7, batch), "alpha": 60 + rng.normal(size=batch)}
        per_instance = kwargs.pop("per_instance", None)
        self.assy = FreeAssembly(self, **kwargs)
        self.shearable_rods = self.assy.build(rod_database_path, assembly_config_path, verbose=verbose)
        self.simulator = self.assy.simulator
        self.data_rods = self.assy.generate_callbacks(self.step_skip, time_interval=self.capture_interval)
        self.simulator.batch = self.batch
        self.simulator.device = self.device
        self.simulator.finalize()
        self.engine = self.simulator.engine
        if per_instance:
            from . import _sweep

            start, stop = self.batch_range
            local = {key: (value if np.ndim(value) == 0 else np.asarray(value)[start:stop] if len(value) == self.global_batch
                           else value) for key, value in per_instance.items()}
            self.engine.set_instance_params(*_sweep.instance_tables(self.assy, self.simulator.program, local, self.batch,
                                                                    self.time_step))
        self.do_step, self.stages_and_updates = extend_stepper_interface(self.StatefulStepper, self.simulator)
        self.time = start_time

    # ------------------------------------------------------------------ run (reference :226-352)
    def _push_actuation(self):
        program = self.simulator.program
        self.engine.set_pressures([shared[0] for shared in program.action_lists])

    def _state_stack(self, short):
        """[batch, 3 (or 3,3), total] device tensor: one field concatenated over rods (reference :259-265)."""
        torch = self.engine.torch
        return torch.cat([self.engine.view(r, short) for r in range(len(self.simulator))], dim=-1)

    def run(self, action: dict, duration: Optional[float] = None, disable_progress_bar: bool = False,
            check_nan: bool = False, check_steady_state: Optional[int] = None) -> TerminalInfo:
        status = TerminalInfo()
        if self.world_size > 1:
            # global-length action arrays are sliced to this rank's share; local-length ones pass through
            action = {k: (v if np.ndim(v) == 0 or len(v) == self.batch else
                          _dist.shard_action({k: v}, self.global_batch, self.world_size, self.rank)[k])
                      for k, v in action.items()}
        self.assy.set_actuation(action)
        self._push_actuation()
        engine, dt = self.engine, self.time_step

        previous = None
        if check_steady_state == 2:
            previous = engine.state.clone()  # the reference concatenates copies of six arrays per rod (:255-266)

        if not duration:
            duration = dt
        n_steps, _ = _native.count_steps(self.time, duration, dt)
        plan = _capture_plan(self.simulator, self.time, n_steps, dt)

        progress = None
        if not disable_progress_bar:
            from tqdm import tqdm

            progress = tqdm(total=self.time + duration, mininterval=0.5,
                            bar_format="{desc}: {percentage:.3f}%|{bar}| {n:.5f}/{total_fmt} [{elapsed}<{remaining}")
        time, done = self.time, 0
        stops = [ordinal for ordinal, _ in plan]
        if not stops or stops[-1] != n_steps:
            stops.append(n_steps)
        for stop in stops:
            chunk = stop - done
            if chunk > 0:
                time = self.do_step(self.StatefulStepper, self.stages_and_updates, self.simulator, time, dt,
                                    n_steps=chunk, callbacks=True, push_actuation=False)
                done = stop
                if progress is not None:
                    progress.update(chunk * dt)
        if progress is not None:
            progress.close()
        self.time = time

        # instances that left the optimistic kernel's validity range with no exact-path kernel to redo them (cluster-sized
        # assemblies, per-instance material tables) stop advancing: never hand back their stale state silently.  One
        # [batch]-int read-back; the verdict is OR-ed over the ranks so that all of them raise together instead of one
        # rank leaving the others inside the gather below.
        frozen = engine.unresolved_count()
        if _dist.reduce_flags(frozen > 0):
            raise RuntimeError(
                "%d instance(s) of this rank left the fast kernel's validity range (|rotation angle|, |curvature angle| or "
                "|damper exponent| beyond its series) and no exact-path kernel fits this assembly; their state stopped "
                "advancing at that step -- reduce the time step or the load, or run them on a one-CTA assembly" % frozen)

        # ---- terminal checks (reference :289-349): one device kernel reduces the batch state to a few numbers per
        # instance (br2_metrics); one gather brings the ranks' rows together
        if check_steady_state in (1, 2) or check_nan:
            metrics = engine.metrics(previous)
            metrics = _dist.gather_batch(metrics, self.global_batch, device=engine.device).cpu().numpy()
        if check_steady_state == 1:
            # the reference measures node POSITIONS here although it calls them velocities (:292-298)
            per_instance = metrics[:, 1]
            max_velocity = float(per_instance.max()) if not np.isnan(per_instance).any() else float("nan")
            status.velocity_steady_state_status = max_velocity < self.velocity_threshold
            status.max_velocity = max_velocity
            status.max_velocity_per_instance = per_instance
        elif check_steady_state == 2:
            # the reference computes these six deltas and discards them (:300-338); they are exposed
            # under a name its combined_* properties do not pick up.
            status.steady_state_deltas = {short: metrics[:, 3 + i] for i, short in enumerate(("x", "v", "a", "Q", "w", "alpha"))}
        if check_nan:
            mask = metrics[:, 0].astype(np.int64)
            flags = {}
            for bit, label in enumerate(("position", "velocity", "director", "omega")):
                flags[label] = (mask >> bit & 1).astype(bool)
                setattr(status, "%s_nan_status" % label, bool(flags[label].any()))
            status.nan_per_instance = flags
        status.end_status = True
        return status

    # ------------------------------------------------------------------ checkpoint (reference :354-392)
    def save_state(self, directory: Optional[str] = None, verbose: bool = False) -> None:
        self.assy.save_state(directory=directory or self.paths.simulation, time=self.time, verbose=verbose)

    def load_state(self, directory: str = "", clear_callback: bool = False, verbose: bool = False) -> None:
        self.assy.load_state(directory=directory or self.paths.simulation, verbose=verbose)
        if clear_callback:
            for sink in self.data_rods:
                sink.clear()
            if verbose:
                print("callback cleared")

    # ------------------------------------------------------------------ export (reference :394-464)
    def render_video(self, visualize_twist_angle: bool = False, **kwargs) -> None:
        """The reference renders with matplotlib / ffmpeg and then exports the positions (:394-429).  Rendering is host
        post-processing outside the path this package replaces: the export the call ends with is done FIRST (so the data
        the reference's own ``br2.visualize`` functions need is on disk), then ``br2.visualize`` says what to feed them."""
        self.save_data("position")
        from .visualize import plot_video_with_surface, visual_twist_with_surface

        save_folder = self.paths.renderings
        plot_video_with_surface(self.data_rods, video_name="br2_simulation", fps=self.rendering_fps, step=1,
                                save_folder=save_folder, **kwargs)
        if visualize_twist_angle:
            visual_twist_with_surface(self.data_rods, video_name="br2_simulation", fps=self.rendering_fps, step=1,
                                      save_folder=save_folder, **kwargs)

    def save_data(self, tag: Optional[str] = None) -> None:
        """``time``, element-centre ``positions [rod, T, (batch,) 3, n]`` and ``directors
        [rod, T, (batch,) 3, 3, n]`` to ``result_<tag>/data/br2_data[_tag].npz`` (reference :431-464)."""
        filename = "br2_data_%s.npz" % tag if tag else "br2_data.npz"
        os.makedirs(self.paths.data, exist_ok=True)
        positions, directors = [], []
        for sink in self.data_rods:
            nodes = np.array(sink["position"])
            positions.append(0.5 * (nodes[..., 1:] + nodes[..., :-1]))
            directors.append(np.array(sink["director"]))
        np.savez(os.path.join(self.paths.data, filename), time=np.array(self.data_rods[0]["time"]),
                 positions=np.asarray(positions), directors=np.asarray(directors))

    def gather_data(self, dst=0):
        """Callback histories of ALL instances on rank ``dst`` (same layout as ``data_rods``, batch axis =
        global batch); other ranks get ``None``.  Device-side captures are gathered device to device, ONE transfer per
        capture for all rods and fields (the payload of /root/reference/br2/free_simulator.py:56-72), and stay lazy on
        ``dst``; histories recorded on the host protocol are gathered per field."""
        if self.world_size == 1:
            return self.data_rods
        captures = _dist.gather_captures(self.data_rods, self.engine, self.global_batch, dst=dst)
        gathered = [_dist.gather_history(sink, self.global_batch, dst=dst, local_batch=self.batch, gathered_captures=captures)
                    for sink in self.data_rods]
        return gathered if self.rank == dst else None

    def gather_field(self, rod_index, attribute, dst=None):
        """One rod attribute (e.g. ``"position_collection"``) of ALL instances, ``[global_batch, ...]``."""
        local = np.asarray(getattr(self.simulator[rod_index], attribute))
        if self.batch == 1:
            local = local[None]
        return _dist.gather_batch(local, self.global_batch, dst=dst)

    def close(self):
        engine = getattr(self, "engine", None)
        if engine is not None:
            engine.close()
