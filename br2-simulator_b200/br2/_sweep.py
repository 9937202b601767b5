"""Per-instance material parameters: sweeps over Young's modulus, shear modulus, damping constant and fibre angles inside
ONE batch (the rod-spec keys of /root/reference/br2/free_simulator.py:108-120, :291-340).

With the reference such a sweep is one simulation per parameter set, each built from its own rod library.  Here the batch
shares geometry, masses and topology -- everything the tile plan is made of -- and every instance gets its own copy of the
few constants the parameters enter:

    youngs_modulus    E A (stretch), E I (bending), the surface joints' k = pi r_o E / n and the tip-to-base joints' k
    shear_modulus     alpha_c G A (shear), G I3 (twist)
    damping_constant  exp(-nu dt) and exp(-nu dt m / J)
    alpha, beta       angled fibres: twist couple per unit pressure  pi r_i^3 (sin^2 a + 2 cos^2 a) / sin 2a / n
    gamma             straight fibres: direction of the bending couple  pi r r_i^2 (cos g, sin g, 0)

The expressions below are the builder's own (hooks/rod.py ``straight_rod``, hooks/ops.py ``AnalyticalLinearDamper``,
free_simulator.py), evaluated for all instances at once and in the same order of operations, so an instance's constants
are bit for bit the ones a separate build from a modified rod library would produce (tests/test_sweep.py checks that).
"""
import numpy as np

from . import _lowering as L
from .hooks.rod import ALPHA_C

KEYS = ("youngs_modulus", "shear_modulus", "damping_constant", "alpha", "beta", "gamma")


def _column(value, batch, name):
    array = np.asarray(value, dtype=np.float64)
    if array.ndim == 0:
        return np.full(batch, float(array))
    if array.shape != (batch,):
        raise ValueError("per_instance[%r] has shape %s; expected a scalar or %d values" % (name, array.shape, batch))
    return array.copy()


def instance_tables(assy, program, per_instance, batch, dt):
    """(uni [batch, FC_COUNT], act_d [batch, n_act, 3], joint_scale [batch]) for ``Engine.set_instance_params``."""
    unknown = set(per_instance) - set(KEYS)
    if unknown:
        raise KeyError("per_instance keys %s are not supported (supported: %s)" % (sorted(unknown), ", ".join(KEYS)))
    if not (program.fast["ok"] and program.fast["uniform"]):
        raise NotImplementedError("per-instance parameters need an assembly of identical, uniform rods (the tile kernel's "
                                  "assemblies): " + (program.fast["why"] or "element constants differ along or between the rods"))
    names = list(assy.free.keys())
    specs = [assy.rod_specs[name] for name in names]
    rods = [assy.free[name] for name in names]
    first, rod = specs[0], rods[0]
    for key in ("youngs_modulus", "shear_modulus", "damping_constant", "outer_radius", "inner_radius", "n_elements", "density"):
        if any(spec.get(key) != first.get(key) for spec in specs):
            raise NotImplementedError("per-instance parameters need rods that share %r" % key)

    E = _column(per_instance.get("youngs_modulus", first["youngs_modulus"]), batch, "youngs_modulus")
    G = _column(per_instance.get("shear_modulus", first["shear_modulus"]), batch, "shear_modulus")
    uni = np.repeat(program.tables["fast_c"][:, 0][None, :], batch, axis=0)

    # ---- elastic constants (hooks/rod.py straight_rod, element 0 / Voronoi cell 0 of the first rod) -------------------
    radius = float(first["base_radius"])
    area = np.pi * radius * radius
    I1 = area * area / (4.0 * np.pi)
    I3 = 2.0 * I1
    l0, l1 = rod.rest_lengths[0], rod.rest_lengths[1] if rod.n_elems > 1 else rod.rest_lengths[0]
    uni[:, L.FC_SHEAR] = uni[:, L.FC_SHEAR + 1] = ALPHA_C * G * area
    uni[:, L.FC_SHEAR + 2] = E * area
    if rod.n_elems > 1:
        for axis, elem in enumerate((E * I1, E * I1, G * I3)):
            uni[:, L.FC_BEND + axis] = (elem * l1 + elem * l0) / (l1 + l0)

    # ---- damper (hooks/ops.py AnalyticalLinearDamper) -----------------------------------------------------------------
    if "damping_constant" in per_instance:
        if first.get("damping_constant") is None:
            raise NotImplementedError("the rods were built without a damper")
        nu = _column(per_instance["damping_constant"], batch, "damping_constant")
        mass = rod.mass
        element_mass = 0.5 * (mass[1] + mass[0]) + 0.5 * mass[0]
        if rod.n_elems == 1:
            element_mass += 0.5 * mass[-1]
        inv_j = np.diagonal(rod.inv_mass_second_moment_of_inertia[:, :, 0])
        uni[:, L.FC_CT] = np.exp(-nu * dt)
        for axis in range(3):
            ln_cr = np.log(np.exp(-nu * dt * element_mass * inv_j[axis]))
            uni[:, L.FC_LN_CR + axis] = ln_cr
            uni[:, L.FC_CR + axis] = np.exp(ln_cr)

    # ---- joints (free_simulator.py build: parameters from the segment's last rod) --------------------------------------
    k0 = np.pi * first["outer_radius"] * first["youngs_modulus"] / first["n_elements"] * assy.k_multiplier
    k = np.pi * first["outer_radius"] * E / first["n_elements"] * assy.k_multiplier
    has_joint = program.tables["fast_c"][L.FC_JK, 0] != 0.0
    if has_joint:
        uni[:, L.FC_JK] = k
    joint_scale = k / k0

    # ---- fibres: the rows of act_d in the order br2/_lowering.py emits them ----------------------------------------------
    n_act = int(program.fast["n_act"])
    act_d = np.zeros((batch, n_act, 3))
    row = 0
    angles = {kind: (_column(per_instance[kind], batch, kind) if kind in per_instance else None) for kind in ("alpha", "beta", "gamma")}
    for spec, this in zip(specs, rods):
        n = this.n_elems
        if not spec["actuated"]:
            continue
        twist, bend = [], []
        for kind in ("alpha", "beta"):
            for degrees in spec[kind]:
                deg = angles[kind] if angles[kind] is not None else np.full(batch, float(degrees))
                angle = deg * np.pi / 180
                scale = (np.pi * (this.inner_radius ** 3) * ((np.sin(angle) ** 2) + 2 * (np.cos(angle) ** 2)) / (np.sin(2 * angle)))
                twist.append(np.stack([np.zeros(batch), np.zeros(batch), scale / n], axis=1))
        scale = np.pi * this._host["rad"][-1] * (this.inner_radius ** 2)  # (the rest radius the builder read: host template)
        for degrees in spec["gamma"]:
            # a swept gamma replaces the library's angle; the segment's y-rotation is added on top, as the builder does
            deg = angles["gamma"] + spec.get("y_rotation", 0.0) if angles["gamma"] is not None else np.full(batch, float(degrees))
            z = deg * np.pi / 180.0
            bend.append(np.stack([scale * np.cos(z), scale * np.sin(z), np.zeros(batch)], axis=1))
        blocks = (twist if twist else []) + twist + bend  # rows of elements 0 .. n-2, then of the last element
        for coeff in blocks:
            act_d[:, row] = coeff
            row += 1
    if row != n_act:
        raise RuntimeError("internal: %d actuation rows rebuilt, the program holds %d" % (row, n_act))
    return uni, act_d, joint_scale
