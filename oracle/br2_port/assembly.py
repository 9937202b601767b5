"""Assembly builder of the reference, restated (TEST INFRASTRUCTURE).

Follows /root/reference/br2/free_simulator.py:76-461 -- which operators get registered on which
rod, with which parameters, in which order -- including the behaviours listed in SURVEY.md
Appendix C (hollow scaling never runs, joint parameters leak from the segment's last rod,
ring of parallel pairs built with a negative index, all-pairs serial joints, ...).
"""
import json
from collections import defaultdict

import numpy as np

from ..mini_elastica import (
    AnalyticalLinearDamper,
    BaseSystemCollection,
    CallBackBaseClass,
    CallBacks,
    Connections,
    Constraints,
    CosseratRod,
    Damping,
    Forcing,
    GravityForces,
)
from ..mini_elastica.experimental.connection_contact_joint.parallel_connection import (
    SurfaceJointSideBySide,
    get_connection_vector_straight_straight_rod,
)
from ..mini_elastica.restart import load_state, save_state
from .ops import (
    FreeBaseEndSoftFixed,
    FreeBendActuation,
    FreeTwistActuation,
    TipToTipStraightJoint,
)


class BR2Simulator(BaseSystemCollection, Constraints, Connections, Forcing, Damping, CallBacks):
    """free_simulator.py:36-39 (same base order, hence same feature ordering)."""


CALLBACK_FIELDS = (
    "position",
    "velocity",
    "acceleration",
    "omega",
    "director",
    "external_forces",
    "external_torques",
    "internal_forces",
    "internal_torques",
    "lengths",
    "dilatation",
    "radius",
)
_FIELD_ATTR = {
    "position": "position_collection",
    "velocity": "velocity_collection",
    "acceleration": "acceleration_collection",
    "omega": "omega_collection",
    "director": "director_collection",
}


class FreeCallback(CallBackBaseClass):
    """free_simulator.py:42-73: every ``step_skip``-th step, append copies of 12 arrays + com."""

    def __init__(self, step_skip, callback_params, time_interval=None):
        super().__init__()
        self.every = step_skip
        self.time_interval = time_interval
        self.callback_params = callback_params

    def make_callback(self, system, time, current_step):
        if current_step % self.every != 0:
            return
        if self.time_interval is not None and not (
            self.time_interval[0] <= time <= self.time_interval[1]
        ):
            return
        out = self.callback_params
        out["time"].append(time)
        out["step"].append(current_step)
        for field in CALLBACK_FIELDS:
            out[field].append(getattr(system, _FIELD_ATTR.get(field, field)).copy())
        out["com"].append(system.compute_position_center_of_mass())


class FreeAssembly:
    def __init__(self, env, gravity=False, **kwargs):
        self.env = env
        self.simulator = BR2Simulator()
        self.actuation = defaultdict(list)
        self.simulator._actuation_lists = self.actuation  # for oracle.fused's extractor
        self.free = {}
        self.toggle_gravity = gravity
        # free_simulator.py:88-91 ("t_multiplier" is the reference's spelling)
        self.k_multiplier = kwargs.get("t_multiplier", 1)
        self.nu_multiplier = kwargs.get("nu_multiplier", 1) * 1.0e-3
        self.kt_multiplier = kwargs.get("kt_multiplier", 1)
        self.k_repulsive = kwargs.get("k_repulsive", 1)

    # -- checkpointing (free_simulator.py:93-99) ---------------------------------------
    def save_state(self, **kwargs):
        save_state(self.simulator, **kwargs)

    def load_state(self, **kwargs):
        load_state(self.simulator, **kwargs)

    # -- build (free_simulator.py:101-268) ---------------------------------------------
    def build(self, rod_info, connect_info, verbose=False, debug=False):
        with open(rod_info) as fh:
            library = json.load(fh)
        rod_specs = library["Rods"]
        defaults = library["DefaultParams"]
        for spec in rod_specs.values():
            for key, value in defaults.items():
                spec.setdefault(key, value)

        with open(connect_info) as fh:
            layout = json.load(fh)
        segments = layout["Segments"]
        activations = layout["Activations"]
        self.num_activation = len(activations)
        self.num_segment = len(segments)

        y_cursor = 0.0
        previous_names = None
        for seg_idx, (seg_name, seg) in enumerate(segments.items()):
            names, lengths, counts = [], [], []
            for rod_i, rod_type in enumerate(seg["rod_order"]):
                spec = rod_specs[rod_type].copy()
                spec["direction"] = np.array(spec["direction"])
                spec["normal"] = np.array(spec["normal"])
                lengths.append(spec["base_length"])
                counts.append(spec["n_elements"])
                axis = np.argmax(spec["direction"])
                spec["start"] = np.insert(np.array(seg["base_position"][rod_i]), axis, y_cursor)
                spec["is_first_segment"] = seg_idx == 0
                spec["base_radius"] = spec["outer_radius"]
                if "gamma" in spec:
                    spec["gamma"] = [g + seg["y-rotation"][rod_i] for g in spec["gamma"]]

                action = None
                for candidate, members in activations.items():
                    for member in members:
                        if member[0] == seg_name and member[1] == rod_i:
                            action = candidate
                    if action is not None:
                        break

                name = "%s_%d_%s" % (seg_name, rod_i, rod_type)
                self.free[name] = self.add_free(name, action, **spec)
                names.append(name)
            assert lengths.count(lengths[0]) == len(lengths)
            assert counts.count(counts[0]) == len(counts)
            y_cursor += lengths[0]

            # joint parameters come from the LAST rod's spec of the segment (Appendix C.9)
            k_joint = np.pi * spec["outer_radius"] * spec["youngs_modulus"] / spec["n_elements"] * self.k_multiplier
            nu_joint = spec["base_length"] / spec["n_elements"] * self.nu_multiplier
            kt_joint = spec["outer_radius"] / 2 * self.kt_multiplier

            if len(names) > 1:
                # ring (R-1,0),(0,1),...,(R-2,R-1) via python's negative index (Appendix C.8)
                for rod_i in range(len(names)):
                    self.add_parallel_connection(
                        names[rod_i - 1],
                        names[rod_i],
                        k=k_joint,
                        nu=nu_joint,
                        kt=kt_joint,
                        k_repulsive=self.k_repulsive,
                    )
            if seg_idx > 0:
                self.add_serial_connection(previous_names, names, k=k_joint, nu=nu_joint, kt=kt_joint)
            previous_names = list(names)

        self.shearable_rods = self.free
        return self.shearable_rods

    def generate_callbacks(self, step_skip, time_interval=None):
        # NB (free_simulator.py:273): time_interval lands in add_callback's ``callback`` slot
        return [self.add_callback(name, step_skip, time_interval) for name in self.free]

    def set_actuation(self, actuation):
        for key, value in actuation.items():
            self.actuation[key][0] = value

    def get_actuation_reference(self, actuation_name=None):
        if actuation_name is None:
            return None
        ref = self.actuation[actuation_name]
        ref.append(0.0)
        return ref

    def create_rod(self, name, is_first_segment=True, verbose=False, **rod_spec):
        rod = CosseratRod.straight_rod(**rod_spec)
        rod.outer_radius = rod_spec["outer_radius"]
        rod.inner_radius = rod_spec["inner_radius"]
        self.free[name] = rod
        self.simulator.append(rod)
        # free_simulator.py:301 does getattr(<dict>, "hollow", False) -> always False: the hollow
        # stiffness scaling never runs.  Restated as the no-op it is.
        if "damping_constant" in rod_spec:
            self.simulator.dampen(rod).using(
                AnalyticalLinearDamper,
                damping_constant=rod_spec["damping_constant"],
                time_step=self.env.time_step,
            )
        if is_first_segment:
            self.simulator.constrain(rod).using(
                FreeBaseEndSoftFixed,
                constrained_position_idx=(0,),
                constrained_director_idx=(0,),
                k=1e9,
                nu=0,
                kt=0.0,
            )
        if self.toggle_gravity:
            self.simulator.add_forcing_to(rod).using(
                GravityForces, acc_gravity=np.array([0.0, 9.80665, 0.0])
            )
        return rod

    def add_angled_fibers(self, rod, actuation_ref, fiber_angles):
        for alpha in fiber_angles:
            angle = alpha * np.pi / 180
            scale = (
                np.pi
                * (rod.inner_radius ** 3)
                * ((np.sin(angle) ** 2) + 2 * (np.cos(angle) ** 2))
                / (np.sin(2 * angle))
            )
            self.simulator.add_forcing_to(rod).using(FreeTwistActuation, actuation_ref, scale=scale)

    def add_straight_fibers(self, rod, actuation_ref, fiber_angles):
        scale = np.pi * rod.radius[-1] * (rod.inner_radius ** 2)
        for gamma in fiber_angles:
            self.simulator.add_forcing_to(rod).using(
                FreeBendActuation, actuation_ref, z_angle=gamma * np.pi / 180.0, scale=scale
            )

    def add_free(self, name, actuation_name, alpha=[], beta=[], gamma=[], **rod_spec):
        rod = self.create_rod(name, **rod_spec)
        ref = self.get_actuation_reference(actuation_name)
        if ref is not None:
            self.add_angled_fibers(rod, ref, alpha)
            self.add_angled_fibers(rod, ref, beta)
            self.add_straight_fibers(rod, ref, gamma)
        return rod

    def add_parallel_connection(self, name1, name2, **param):
        self.glue_rods_surface_connection(self.free[name1], self.free[name2], **param)

    def add_serial_connection(self, rod_list1, rod_list2, **param):
        for name1 in rod_list1:
            for name2 in rod_list2:
                self.tip_to_base_connection(self.free[name1], self.free[name2], **param)

    def add_callback(self, name, step_skip, callback=None, **kwargs):
        params = defaultdict(list)
        if callback is None:
            callback = FreeCallback
        self.simulator.collect_diagnostics(self.free[name]).using(
            callback, step_skip=step_skip, callback_params=params, **kwargs
        )
        return params

    def glue_rods_surface_connection(self, rod1, rod2, k, nu, kt, k_repulsive):
        # kt is accepted and dropped, as in free_simulator.py:402-435
        dv1, dv2, offset = get_connection_vector_straight_straight_rod(
            rod_one=rod1, rod_two=rod2, rod_one_idx=(0, rod1.n_elems), rod_two_idx=(0, rod2.n_elems)
        )
        for i in range(rod1.n_elems):
            self.simulator.connect(
                first_rod=rod1, second_rod=rod2, first_connect_idx=i, second_connect_idx=i
            ).using(
                SurfaceJointSideBySide,
                k=k,
                nu=nu,
                k_repulsive=k_repulsive,
                rod_one_direction_vec_in_material_frame=dv1[:, i],
                rod_two_direction_vec_in_material_frame=dv2[:, i],
                offset_btw_rods=offset[i],
            )

    def tip_to_base_connection(self, rod1, rod2, k, nu, kt):
        c1 = 0.5 * (rod1.position_collection[..., -1] + rod1.position_collection[..., -2])
        c2 = 0.5 * (rod2.position_collection[..., 0] + rod2.position_collection[..., 1])
        along = (c2 - c1) / np.linalg.norm(c2 - c1)
        self.simulator.connect(
            first_rod=rod1, second_rod=rod2, first_connect_idx=-1, second_connect_idx=0
        ).using(
            TipToTipStraightJoint,
            k=k,
            nu=nu,
            kt=kt,
            rod1_rd2_local=rod1.director_collection[..., -1] @ along,
            rod2_rd2_local=rod2.director_collection[..., 0] @ (-along),
            stability_check=False,
        )
