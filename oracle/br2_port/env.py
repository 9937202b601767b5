"""Run loop of the reference, restated (TEST INFRASTRUCTURE).

Follows /root/reference/br2/environment.py:147-224 (construction / reset) and :226-352 (run):
the step count comes from fp64 accumulation of ``time`` (two ``+dt/2`` per step), not from
``duration/dt``.
"""
import numpy as np

from ..mini_elastica import PositionVerlet, extend_stepper_interface
from ..mini_elastica._calculus import _isnan_check
from .assembly import FreeAssembly


def count_steps(start_time, duration, dt):
    """Number of ``do_step`` calls ``while time < start+duration`` makes, and the final time
    (environment.py:278-286 with the stepper's two half-step time increments)."""
    time = start_time
    end = start_time + duration
    n = 0
    half = 0.5 * dt
    while time < end:
        time += half
        time += half
        n += 1
    return n, time


class OracleEnvironment:
    """Minimal ``Environment`` (no paths / rendering): ``reset`` + ``run`` + NaN / steady checks."""

    def __init__(self, rendering_fps=25, time_step=2.0e-5):
        self.StatefulStepper = PositionVerlet()
        self.time_step = time_step
        self.rendering_fps = rendering_fps
        self.step_skip = int(1.0 / (rendering_fps * time_step))
        self.capture_interval = None
        self.velocity_threshold = 1.0e-4

    def reset(self, rod_database_path, assembly_config_path, start_time=0.0, **kwargs):
        self.assy = FreeAssembly(self, **kwargs)
        self.shearable_rods = self.assy.build(rod_database_path, assembly_config_path)
        self.simulator = self.assy.simulator
        self.data_rods = self.assy.generate_callbacks(self.step_skip, time_interval=self.capture_interval)
        self.simulator.finalize()
        self.do_step, self.stages_and_updates = extend_stepper_interface(
            self.StatefulStepper, self.simulator
        )
        self.time = start_time

    def run(self, action, duration=None, n_steps=None, check_nan=False, check_steady_state=None):
        self.assy.set_actuation(action)
        time = self.time
        if n_steps is not None:
            for _ in range(n_steps):
                time = self.do_step(
                    self.StatefulStepper, self.stages_and_updates, self.simulator, time, self.time_step
                )
        else:
            if not duration:
                duration = self.time_step
            while time < self.time + duration:
                time = self.do_step(
                    self.StatefulStepper, self.stages_and_updates, self.simulator, time, self.time_step
                )
        self.time = time
        status = {"end_status": True}
        rods = list(self.shearable_rods.values())
        if check_steady_state == 1:
            # environment.py:290-298 measures POSITIONS here (labelled velocity upstream)
            max_velocity = max(
                np.linalg.norm(np.array(r.position_collection), axis=0).max() for r in rods
            )
            status["velocity_steady_state_status"] = bool(max_velocity < self.velocity_threshold)
            status["max_velocity"] = float(max_velocity)
        if check_nan:
            status["position_nan_status"] = any(_isnan_check(r.position_collection) for r in rods)
            status["velocity_nan_status"] = any(_isnan_check(r.velocity_collection) for r in rods)
            status["director_nan_status"] = any(_isnan_check(r.director_collection) for r in rods)
            status["omega_nan_status"] = any(_isnan_check(r.omega_collection) for r in rods)
        return status
