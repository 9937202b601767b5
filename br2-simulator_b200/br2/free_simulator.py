"""Assembly builder: rod library + assembly JSON -> rods and operators registered on a
``BR2Simulator`` through the PyElastica-style hooks.

Drop-in for /root/reference/br2/free_simulator.py (same class and method names / signatures):
``BR2Simulator`` (:36-39), ``FreeCallback`` (:42-73), ``FreeAssembly`` (:76-461).  What gets
registered, on which rod, with which parameters and in which order is the reference's, including the
behaviours SURVEY.md Appendix C lists:

* rods are solid cylinders of radius ``outer_radius``; the hollow stiffness scaling never runs in the
  reference (``getattr`` on a dict, :301) and is not applied here either; ``inner_radius`` only enters
  the actuation scales (:347, :356);
* a rod's action is the first ``Activations`` entry naming ``[segment, rod_index]`` (:170-176); rods
  without one get no fibre forcing (:371);
* parallel joints form the ring (R-1,0),(0,1),...,(R-2,R-1) -- python negative index (:202-204); one
  ``SurfaceJointSideBySide`` per element per pair (:417-435); ``kt`` is accepted and dropped;
* joint parameters come from the LAST rod's spec of the segment (:193-200, :220-228);
* from the second segment on, every rod of the previous segment is tip-to-base joined to every rod
  of the current one (:383-389), and only first-segment rods are base-constrained (:323).

The integration itself is not here: ``Environment.reset`` calls ``simulator.finalize()``, which lowers
everything registered below to device tables (br2/_lowering.py).
"""
import json
from collections import defaultdict

import numpy as np

from .free_custom_systems import FreeBaseEndSoftFixed, FreeBendActuation, FreeTwistActuation
from .hooks import (
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
    SurfaceJointSideBySide,
    get_connection_vector_straight_straight_rod,
)
from .legacy_surface_connection_parallel_rod_numba import TipToTipStraightJoint


class BR2Simulator(BaseSystemCollection, Constraints, Connections, Forcing, Damping, CallBacks):
    pass


_CALLBACK_ARRAYS = (
    ("position", "position_collection"),
    ("velocity", "velocity_collection"),
    ("acceleration", "acceleration_collection"),
    ("omega", "omega_collection"),
    ("director", "director_collection"),
    ("external_forces", "external_forces"),
    ("external_torques", "external_torques"),
    ("internal_forces", "internal_forces"),
    ("internal_torques", "internal_torques"),
    ("lengths", "lengths"),
    ("dilatation", "dilatation"),
    ("radius", "radius"),
)


class FreeCallback(CallBackBaseClass):
    """Appends copies of the rod's arrays to ``callback_params`` every ``step_skip`` steps; keys and
    shapes as in the reference (:56-72) so its post-processing keeps working.  For ``batch > 1`` every
    array gains a leading batch axis."""

    def __init__(self, step_skip, callback_params, time_interval=None):
        super().__init__()
        self.every = step_skip
        self.time_interval = time_interval
        self.callback_params = callback_params

    def make_callback(self, system, time, current_step):
        if current_step % self.every:
            return
        window = self.time_interval
        if window is not None and (time < window[0] or time > window[1]):
            return
        sink = self.callback_params
        sink["time"].append(time)
        sink["step"].append(current_step)
        for key, attribute in _CALLBACK_ARRAYS:
            sink[key].append(np.array(getattr(system, attribute), copy=True))
        sink["com"].append(system.compute_position_center_of_mass())

    def make_callback_lazy(self, capture, rod_index, rod, time, current_step):
        """Same record as ``make_callback`` from a device-side capture of the batch state (br2/_capture.py): nothing is
        copied to the host until an entry is read.  Used when the sink is the ``CaptureSink`` this package creates."""
        from ._capture import _Lazy
        from .hooks.rod import DYNAMIC_FIELDS

        if current_step % self.every:
            return
        window = self.time_interval
        if window is not None and (time < window[0] or time > window[1]):
            return
        sink = self.callback_params
        sink["time"].append(time)
        sink["step"].append(current_step)
        n = rod.n_elems
        for key, attribute in _CALLBACK_ARRAYS:
            short, shape = DYNAMIC_FIELDS[attribute]
            sink[key].append(_Lazy(capture, rod_index, short, shape(n)))
        mass = np.asarray(rod.mass)
        sink["com"].append(_Lazy(capture, rod_index, "x", (3, n + 1), post=lambda x: (x * mass).sum(axis=-1) / mass.sum()))


def _load_rod_library(path):
    with open(path) as handle:
        library = json.load(handle)
    defaults = library["DefaultParams"]
    rods = library["Rods"]
    for spec in rods.values():
        for key in defaults:
            if key not in spec:
                spec[key] = defaults[key]
    return rods


def _first_action_of(activations, segment_name, rod_position):
    """First action whose member list names this rod; within one action every member is scanned
    (reference :170-176)."""
    for action, members in activations.items():
        if any(m[0] == segment_name and m[1] == rod_position for m in members):
            return action
    return None


class FreeAssembly:
    def __init__(self, env, gravity=False, **kwargs):
        self.env = env
        self.simulator = BR2Simulator()
        self.actuation = defaultdict(list)
        self.free = {}  # "<segment>_<position>_<rod type>" -> rod
        self.rod_specs = {}  # same keys -> the rod spec it was built from (br2/_sweep.py recomputes constants from it)
        self.toggle_gravity = gravity
        # "t_multiplier" (sic) scales the joint stiffness in the reference (:88)
        self.k_multiplier = kwargs.get("t_multiplier", 1)
        self.nu_multiplier = kwargs.get("nu_multiplier", 1) * 1.0e-3
        self.kt_multiplier = kwargs.get("kt_multiplier", 1)
        self.k_repulsive = kwargs.get("k_repulsive", 1)

    # ------------------------------------------------------------------ checkpoint (reference :93-99)
    def save_state(self, **kwargs):
        from ._restart import save_state

        save_state(self.simulator, **kwargs)

    def load_state(self, **kwargs):
        from ._restart import load_state

        load_state(self.simulator, **kwargs)

    # ------------------------------------------------------------------ build (reference :101-268)
    def build(self, rod_info, connect_info, verbose=True, debug=False):
        say = print if verbose else (lambda *a, **k: None)
        rod_specs = _load_rod_library(rod_info)
        with open(connect_info) as handle:
            layout = json.load(handle)
        segments, activations = layout["Segments"], layout["Activations"]
        self.num_activation = len(activations)
        self.num_segment = len(segments)

        axial_offset = 0.0
        previous = None
        for seg_idx, seg_name in enumerate(segments):
            seg = segments[seg_name]
            members = []
            lengths, counts = [], []
            for position, rod_type in enumerate(seg["rod_order"]):
                spec = dict(rod_specs[rod_type])
                spec["direction"] = np.array(spec["direction"])
                spec["normal"] = np.array(spec["normal"])
                lengths.append(spec["base_length"])
                counts.append(spec["n_elements"])
                axis = int(np.argmax(spec["direction"]))
                spec["start"] = np.insert(np.array(seg["base_position"][position]), axis, axial_offset)
                spec["is_first_segment"] = seg_idx == 0
                spec["base_radius"] = spec["outer_radius"]
                if "gamma" in spec:
                    spec["gamma"] = [angle + seg["y-rotation"][position] for angle in spec["gamma"]]
                    spec["y_rotation"] = seg["y-rotation"][position]  # (kept for br2/_sweep.py)
                name = "%s_%d_%s" % (seg_name, position, rod_type)
                self.free[name] = self.add_free(name, _first_action_of(activations, seg_name, position), **spec)
                members.append(name)
            if len(set(lengths)) != 1:
                raise AssertionError("rods' lengths should all be the same within a segment")
            if len(set(counts)) != 1:
                raise AssertionError("rods' number of elements should all be the same within a segment")
            axial_offset += lengths[0]

            # joint parameters: taken from the spec of the segment's LAST rod, as in the reference
            k_joint = np.pi * spec["outer_radius"] * spec["youngs_modulus"] / spec["n_elements"] * self.k_multiplier
            nu_joint = spec["base_length"] / spec["n_elements"] * self.nu_multiplier
            kt_joint = spec["outer_radius"] / 2 * self.kt_multiplier
            if len(members) > 1:
                say("connecting in parallel...")
                for position in range(len(members)):
                    first, second = members[position - 1], members[position]
                    say("    connecting seg %d: %s || %s" % (seg_idx, first, second))
                    self.add_parallel_connection(first, second, k=k_joint, nu=nu_joint, kt=kt_joint,
                                                 k_repulsive=self.k_repulsive)
            if seg_idx > 0:
                say("connecting in serial...")
                say("  connecting seg-%d and seg-%d" % (seg_idx, seg_idx + 1))
                self.add_serial_connection(previous, members, k=k_joint, nu=nu_joint, kt=kt_joint)
            previous = list(members)

        self.shearable_rods = self.free
        return self.shearable_rods

    def generate_callbacks(self, step_skip, time_interval=None):
        # the reference passes ``time_interval`` positionally into add_callback's ``callback`` slot
        # (:273); with the default None that selects FreeCallback, which is kept.
        return [self.add_callback(name, step_skip, time_interval) for name in self.free]

    # ------------------------------------------------------------------ actuation (reference :276-289)
    def set_actuation(self, actuation):
        """``actuation[name]`` may be a scalar (all instances) or a length-batch array."""
        for name, value in actuation.items():
            self.actuation[name][0] = value  # IndexError on an unknown action, like the reference

    def get_actuation_reference(self, actuation_name=None):
        if actuation_name is None:
            return None
        shared = self.actuation[actuation_name]
        shared.append(0.0)
        return shared

    # ------------------------------------------------------------------ rods and their operators
    def create_rod(self, name, is_first_segment=True, verbose=False, **rod_spec):
        rod = CosseratRod.straight_rod(**rod_spec)
        rod.outer_radius = rod_spec["outer_radius"]
        rod.inner_radius = rod_spec["inner_radius"]
        self.free[name] = rod
        self.simulator.append(rod)
        if "damping_constant" in rod_spec:
            self.simulator.dampen(rod).using(AnalyticalLinearDamper, damping_constant=rod_spec["damping_constant"],
                                             time_step=self.env.time_step)
        if is_first_segment:
            self.simulator.constrain(rod).using(FreeBaseEndSoftFixed, constrained_position_idx=(0,),
                                                constrained_director_idx=(0,), k=1e9, nu=0, kt=0.0)
        if self.toggle_gravity:
            # +y, as in the reference (:337)
            self.simulator.add_forcing_to(rod).using(GravityForces, acc_gravity=np.array([0.0, 9.80665, 0.0]))
        return rod

    def add_angled_fibers(self, rod, actuation_ref, fiber_angles):
        for degrees in fiber_angles:
            angle = degrees * np.pi / 180
            scale = (np.pi * (rod.inner_radius ** 3) * ((np.sin(angle) ** 2) + 2 * (np.cos(angle) ** 2))
                     / (np.sin(2 * angle)))
            self.simulator.add_forcing_to(rod).using(FreeTwistActuation, actuation_ref, scale=scale)

    def add_straight_fibers(self, rod, actuation_ref, fiber_angles):
        scale = np.pi * rod.radius[-1] * (rod.inner_radius ** 2)
        for degrees in fiber_angles:
            self.simulator.add_forcing_to(rod).using(FreeBendActuation, actuation_ref,
                                                     z_angle=degrees * np.pi / 180.0, scale=scale)

    def add_free(self, name, actuation_name, alpha=[], beta=[], gamma=[], **rod_spec):
        rod = self.create_rod(name, **rod_spec)
        shared = self.get_actuation_reference(actuation_name)
        self.rod_specs[name] = dict(rod_spec, alpha=list(alpha), beta=list(beta), gamma=list(gamma), actuated=shared is not None)
        if shared is not None:
            self.add_angled_fibers(rod, shared, alpha)
            self.add_angled_fibers(rod, shared, beta)
            self.add_straight_fibers(rod, shared, gamma)
        return rod

    # ------------------------------------------------------------------ joints
    def add_parallel_connection(self, name1, name2, **param):
        self.glue_rods_surface_connection(self.free[name1], self.free[name2], **param)

    def add_serial_connection(self, rod_list1, rod_list2, **param):
        for upstream_name in rod_list1:
            for downstream_name in rod_list2:
                self.tip_to_base_connection(self.free[upstream_name], self.free[downstream_name], **param)

    def add_callback(self, name, step_skip, callback=None, **kwargs):
        from ._capture import CaptureSink

        sink = CaptureSink()  # a defaultdict like the reference's, whose lists may hold device-side captures
        self.simulator.collect_diagnostics(self.free[name]).using(
            FreeCallback if callback is None else callback, step_skip=step_skip, callback_params=sink, **kwargs)
        return sink

    def glue_rods_surface_connection(self, rod1, rod2, k, nu, kt, k_repulsive):
        toward_two, toward_one, gap = get_connection_vector_straight_straight_rod(
            rod_one=rod1, rod_two=rod2, rod_one_idx=(0, rod1.n_elems), rod_two_idx=(0, rod2.n_elems))
        for element in range(rod1.n_elems):
            self.simulator.connect(first_rod=rod1, second_rod=rod2, first_connect_idx=element,
                                   second_connect_idx=element).using(
                SurfaceJointSideBySide, k=k, nu=nu, k_repulsive=k_repulsive,
                rod_one_direction_vec_in_material_frame=toward_two[:, element],
                rod_two_direction_vec_in_material_frame=toward_one[:, element],
                offset_btw_rods=gap[element])

    def tip_to_base_connection(self, rod1, rod2, k, nu, kt):
        x1, x2 = rod1.position_collection, rod2.position_collection
        tip_centre = 0.5 * (x1[..., -1] + x1[..., -2])
        base_centre = 0.5 * (x2[..., 0] + x2[..., 1])
        along = (base_centre - tip_centre) / np.linalg.norm(base_centre - tip_centre)
        self.simulator.connect(first_rod=rod1, second_rod=rod2, first_connect_idx=-1, second_connect_idx=0).using(
            TipToTipStraightJoint, k=k, nu=nu, kt=kt,
            rod1_rd2_local=rod1.director_collection[..., -1] @ along,
            rod2_rd2_local=rod2.director_collection[..., 0] @ (-along),
            stability_check=False)
