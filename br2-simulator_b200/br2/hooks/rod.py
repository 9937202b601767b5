"""Cosserat rod description + device-backed attribute surface.

Replaces ``elastica.CosseratRod`` as the reference uses it: created through
``CosseratRod.straight_rod(**rod_spec)`` (/root/reference/br2/free_simulator.py:293) and then read
through ``position_collection``, ``director_collection``, ``radius``, ``n_elems``, ``mass`` ...
(SURVEY.md section 8b lists every attribute the reference touches).

Before ``finalize()`` the arrays are plain host numpy (the template every batch instance starts
from).  After ``finalize()`` the dynamic fields live in HBM; the same attribute names then return
host copies fetched from the device -- ``[3, n+1]``-style arrays for ``batch == 1`` (drop-in
shapes) and ``[batch, 3, n+1]`` otherwise.  There is no CPU integration path in this package.
"""
import numpy as np

# Kaneko shear-correction factor of PyElastica >= 0.3.0.
ALPHA_C = 27.0 / 28.0
LENGTH_GUARD = 1e-14

# name -> (short key in the device state layout, per-rod shape as a function of n)
DYNAMIC_FIELDS = {
    "position_collection": ("x", lambda n: (3, n + 1)),
    "velocity_collection": ("v", lambda n: (3, n + 1)),
    "director_collection": ("Q", lambda n: (3, 3, n)),
    "omega_collection": ("w", lambda n: (3, n)),
    "acceleration_collection": ("a", lambda n: (3, n + 1)),
    "alpha_collection": ("alpha", lambda n: (3, n)),
    "internal_forces": ("f_int", lambda n: (3, n + 1)),
    "internal_torques": ("t_int", lambda n: (3, n)),
    "external_forces": ("f_ext", lambda n: (3, n + 1)),
    "external_torques": ("t_ext", lambda n: (3, n)),
    "lengths": ("len", lambda n: (n,)),
    "dilatation": ("dil", lambda n: (n,)),
    "radius": ("rad", lambda n: (n,)),
}
# order of the fields inside one rod's block of the instance state (include/br2cuda.h)
STATE_ORDER = ("x", "v", "Q", "w", "a", "alpha", "f_int", "t_int", "f_ext", "t_ext", "len", "dil", "rad")
STATE_IO = ("x", "v", "Q", "w")


def field_sizes(n):
    return {short: int(np.prod(shape(n))) for short, shape in DYNAMIC_FIELDS.values()}


def rod_words(n):
    sizes = field_sizes(n)
    return sum(sizes[s] for s in STATE_ORDER)


class RodBase:
    """Marker base (``elastica.rod.RodBase``): what a system collection accepts."""


def _diag3(values_by_axis, n):
    out = np.zeros((3, 3, n))
    for axis in range(3):
        out[axis, axis, :] = values_by_axis[axis]
    return out


class CosseratRod(RodBase):
    def __init__(self, host):
        self._host = host
        self._engine = None
        self._index = None
        self.n_elems = int(host["n_elems"])

    # ------------------------------------------------------------------ construction
    @classmethod
    def straight_rod(cls, n_elements, start, direction, normal, base_length, base_radius, density,
                     nu=None, youngs_modulus=None, *args, **kwargs):
        """Straight, uniform, unstrained rod at rest.

        Keys the reference's rod spec carries along (``outer_radius``, ``inner_radius``,
        ``damping_constant``, ``hollow`` ...) are accepted and ignored exactly like upstream does;
        ``nu`` (internal dissipation) left PyElastica in 0.3.0 and is ignored too."""
        if youngs_modulus is None:
            raise TypeError("straight_rod() needs youngs_modulus")
        n = int(n_elements)
        if n < 1:
            raise ValueError("a rod needs at least one element")
        start = np.asarray(start, dtype=np.float64)
        direction = np.asarray(direction, dtype=np.float64)
        normal = np.asarray(normal, dtype=np.float64)
        shear_modulus = kwargs.get("shear_modulus", None)
        if shear_modulus is None:
            shear_modulus = youngs_modulus / (2.0 * (1.0 + 0.5))
        alpha_c = kwargs.get("alpha_c", ALPHA_C)

        tip = start + direction * base_length
        x = np.stack([np.linspace(start[c], tip[c], n + 1) for c in range(3)])
        dx = x[:, 1:] - x[:, :-1]
        rest_lengths = np.sqrt(dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2])
        tangents = dx / rest_lengths
        normal = normal / np.linalg.norm(normal)
        if np.abs(normal @ tangents).max() > 1e-12:
            raise ValueError("rod normal must be perpendicular to its direction")
        d1 = np.repeat(normal[:, None], n, axis=1)
        d2 = np.cross(tangents, d1, axis=0)
        # np.cross evaluates a1*b2 - a2*b1 etc., the same products as a hand-written cross
        Q = np.stack([d1, d2, tangents])  # rows are directors: [3, 3, n]

        radius = np.full(n, float(base_radius))
        rho = np.full(n, float(density))
        area = np.pi * radius * radius
        I1 = area * area / (4.0 * np.pi)
        I3 = 2.0 * I1
        rho_l = rho * rest_lengths
        J = _diag3([I1 * rho_l, I1 * rho_l, I3 * rho_l], n)
        Jinv = _diag3([1.0 / (I1 * rho_l), 1.0 / (I1 * rho_l), 1.0 / (I3 * rho_l)], n)
        shear = _diag3([alpha_c * shear_modulus * area, alpha_c * shear_modulus * area, youngs_modulus * area], n)
        bend_elem = _diag3([youngs_modulus * I1, youngs_modulus * I1, shear_modulus * I3], n)
        bend = (bend_elem[..., 1:] * rest_lengths[1:] + bend_elem[..., :-1] * rest_lengths[:-1]) / (
            rest_lengths[1:] + rest_lengths[:-1])
        volume = np.pi * radius ** 2 * rest_lengths
        mass = np.zeros(n + 1)
        mass[:-1] += 0.5 * rho * volume
        mass[1:] += 0.5 * rho * volume
        host = dict(
            n_elems=n,
            x=x, v=np.zeros((3, n + 1)), Q=Q, w=np.zeros((3, n)),
            a=np.zeros((3, n + 1)), alpha=np.zeros((3, n)),
            f_int=np.zeros((3, n + 1)), t_int=np.zeros((3, n)),
            f_ext=np.zeros((3, n + 1)), t_ext=np.zeros((3, n)),
            len=np.zeros(n), dil=np.zeros(n), rad=radius,
            rest_lengths=rest_lengths,
            rest_voronoi_lengths=0.5 * (rest_lengths[1:] + rest_lengths[:-1]),
            volume=volume, mass=mass, density=rho,
            shear_matrix=shear, bend_matrix=bend,
            mass_second_moment_of_inertia=J, inv_mass_second_moment_of_inertia=Jinv,
            rest_sigma=np.zeros((3, n)), rest_kappa=np.zeros((3, n - 1)),
        )
        rod = cls(host)
        rod._refresh_geometry()
        return rod

    def _refresh_geometry(self):
        """Upstream's constructor runs the strain kernels once, which *recomputes* ``radius`` from
        the volume and the guarded length; the reference later reads that value as
        ``rod.radius[-1]`` (/root/reference/br2/free_simulator.py:356) and inside
        ``get_connection_vector_straight_straight_rod`` (:404-413), so it must be reproduced."""
        h = self._host
        dx = h["x"][:, 1:] - h["x"][:, :-1]
        h["len"] = np.sqrt(dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2]) + LENGTH_GUARD
        h["rad"] = np.sqrt(h["volume"] / h["len"] / np.pi)
        h["dil"] = h["len"] / h["rest_lengths"]

    # ------------------------------------------------------------------ attribute surface
    def _bind(self, engine, index):
        self._engine = engine
        self._index = index

    def _get(self, short):
        if self._engine is None:
            return self._host[short]
        return self._engine.fetch(self._index, short)

    def _set(self, short, value):
        if self._engine is None:
            self._host[short][...] = value
        else:
            self._engine.upload(self._index, short, np.asarray(value, dtype=np.float64))

    @property
    def tangents(self):
        x = self.position_collection
        return (x[..., 1:] - x[..., :-1]) / self.lengths[..., None, :]

    # constants: host arrays, shared by every instance of the batch
    rest_lengths = property(lambda self: self._host["rest_lengths"])
    rest_voronoi_lengths = property(lambda self: self._host["rest_voronoi_lengths"])
    volume = property(lambda self: self._host["volume"])
    mass = property(lambda self: self._host["mass"])
    density = property(lambda self: self._host["density"])
    rest_sigma = property(lambda self: self._host["rest_sigma"])
    rest_kappa = property(lambda self: self._host["rest_kappa"])
    mass_second_moment_of_inertia = property(lambda self: self._host["mass_second_moment_of_inertia"])
    inv_mass_second_moment_of_inertia = property(lambda self: self._host["inv_mass_second_moment_of_inertia"])

    # stiffness matrices stay assignable until finalize (the reference's hollow scaling rebinds them,
    # /root/reference/br2/free_simulator.py:310-311 -- dead code there, but part of the surface)
    @property
    def shear_matrix(self):
        return self._host["shear_matrix"]

    @shear_matrix.setter
    def shear_matrix(self, value):
        self._require_unfinalized("shear_matrix")
        self._host["shear_matrix"] = np.asarray(value, dtype=np.float64)

    @property
    def bend_matrix(self):
        return self._host["bend_matrix"]

    @bend_matrix.setter
    def bend_matrix(self, value):
        self._require_unfinalized("bend_matrix")
        self._host["bend_matrix"] = np.asarray(value, dtype=np.float64)

    def _require_unfinalized(self, what):
        if self._engine is not None:
            raise RuntimeError("%s is baked into the device tables at finalize() and cannot be reassigned afterwards" % what)

    def compute_position_center_of_mass(self):
        x = self.position_collection
        return (x * self.mass).sum(axis=-1) / self.mass.sum()

    def compute_velocity_center_of_mass(self):
        v = self.velocity_collection
        return (v * self.mass).sum(axis=-1) / self.mass.sum()


def _dynamic_property(name, short):
    def getter(self):
        return self._get(short)

    def setter(self, value):
        self._set(short, value)

    return property(getter, setter, doc="%s (host array before finalize, device-backed after)" % name)


for _name, (_short, _shape) in DYNAMIC_FIELDS.items():
    setattr(CosseratRod, _name, _dynamic_property(_name, _short))


class HostRodView:
    """Read-only snapshot of one rod handed to diagnostics callbacks
    (``make_callback(system, time, current_step)``, /root/reference/br2/free_simulator.py:49-73)."""

    def __init__(self, rod, arrays):
        self._rod = rod
        self._arrays = arrays
        self.n_elems = rod.n_elems

    def __getattr__(self, name):
        if name in DYNAMIC_FIELDS:
            return self._arrays[DYNAMIC_FIELDS[name][0]]
        return getattr(self._rod, name)

    def compute_position_center_of_mass(self):
        x = self._arrays["x"]
        return (x * self._rod.mass).sum(axis=-1) / self._rod.mass.sum()

    def compute_velocity_center_of_mass(self):
        v = self._arrays["v"]
        return (v * self._rod.mass).sum(axis=-1) / self._rod.mass.sum()
