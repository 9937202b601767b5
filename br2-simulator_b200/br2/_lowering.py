"""Lower a finalized system collection to the flat tables of include/br2cuda.h.

The reference keeps Python operator objects and walks them every step
(/root/reference/br2/environment.py:279-285 -> PyElastica ``do_step``).  Here the same information --
which operator acts on which rod / element, with which parameters, in which order -- becomes a
handful of int32 / fp64 arrays that the CUDA kernel interprets (table formats: include/br2cuda.h).

Notes on semantics that the tables must carry:
* forcing rows are grouped per rod in registration order (upstream sorts stably by rod index);
* joint rows keep registration order; each element's external torque and each node's external
  force receive their joint contributions in that order (CSR gather lists);
* a tip-to-base joint *overwrites* the director and omega of rod two's first element with those of
  rod one's last element while ``synchronize`` runs (legacy_surface_connection_parallel_rod_numba.py:474-478);
  any later joint sees the overwritten director.  ``Q1/Q2`` of a joint row and ``overwrite_src``
  resolve these chains statically to "element whose start-of-synchronize value is used".
"""
import numpy as np

from .hooks import ops as _ops
from .hooks.rod import STATE_ORDER, field_sizes, rod_words

ABI_VERSION = 2
ROD_NI, ROD_ND = 8, 2
SC_COUNT = 25
SC_MASS, SC_REST_LEN, SC_VOLUME, SC_SHEAR, SC_J, SC_JINV, SC_LN_CR, SC_REST_SIGMA = 0, 1, 2, 3, 6, 9, 12, 15
SC_REST_VLEN, SC_BEND, SC_REST_KAPPA = 18, 19, 22
F_NI, F_ND = 4, 6
C_NI, C_ND = 2, 14
J_NI, J_ND = 8, 10
ROD_HAS_DAMPER = 1
MAX_SLOTS = 8 * 640  # a cluster of 8 CTAs x 320 threads x 2 slots; libbr2cuda decides whether the assembly really fits
# fast-path tables (include/br2cuda.h)
FAST_NI, FAST_NODE_ITEMS, FAST_MAX_EXTRA, FAST_EXTRA_ID = 18, 8, 64, 16
FAST_FLAGS, FAST_OWN_JOINT, FAST_EXTRA_JOINT, FAST_SOFTFIX_ROW, FAST_ACT_BEGIN, FAST_ACT_END = 0, 1, 2, 3, 4, 5
FAST_ELEM_ITEM0, FAST_NODE_ITEM0 = 6, 8
PIN_DIRECTOR, PIN_POSITION, ZERO_NODE_RATE, ZERO_ELEM_RATE, ZERO_REPORTED = 1, 2, 4, 8, 16
FC_REST_LEN, FC_INV_REST_LEN, FC_VOL_OVER_PI, FC_SHEAR, FC_J, FC_JINV, FC_LN_CR, FC_CR = 0, 1, 2, 3, 6, 9, 12, 15
FC_REST_VLEN, FC_INV_REST_VLEN, FC_BEND, FC_CT, FC_GRAVITY, FC_JK, FC_JNU, FC_JKREP, FC_PIN, FC_COUNT = (
    18, 19, 20, 23, 24, 27, 28, 29, 30, 42)
FORCE_GRAVITY, FORCE_TWIST, FORCE_BEND, FORCE_POINT = 0, 1, 2, 3
CONS_SOFT_FIXED_BASE, CONS_LAST_END_FIXED = 0, 1
JOINT_SURFACE, JOINT_TIP_TO_BASE = 0, 1
UNIFORM_RTOL = 1e-12  # linspace rounding: 1.4e-13 at 200 elements per rod over 0.54 m


def _diagonal_only(matrix, what):
    off = matrix.copy()
    for axis in range(3):
        off[axis, axis] = 0.0
    if off.any():
        raise _ops.NotLowerable("%s has off-diagonal entries; the device kernel supports diagonal "
                                "constitutive matrices only" % what)
    return np.stack([matrix[0, 0], matrix[1, 1], matrix[2, 2]])


class Program:
    """The lowered assembly: numpy tables + bookkeeping the engine needs."""

    def __init__(self):
        self.tables = {}
        self.layout = []       # per rod: {short: (offset, size)}
        self.n_elems = []
        self.words = 0
        self.n_slots = 0
        self.action_lists = []  # python lists, one per action slot (the reference's shared lists)
        self.action_names = []
        self.total_elements = 0


class LoweringContext:
    def __init__(self, rods):
        self.rods = rods
        self.slot0 = []
        slot = 0
        for rod in rods:
            self.slot0.append(slot)
            slot += rod.n_elems + 1
        self.n_slots = slot
        self.forcing = [[] for _ in rods]      # per rod: (type, action, node, d)
        self.constraints = [[] for _ in rods]  # per rod: (type, d)
        self.dampers = {}
        self.joints = []                       # (type, s1, s2, q1, q2, d, torque)
        self.director_source = {}              # (rod, elem) -> (rod, elem) after overwrites so far
        self.action_lists = []

    # -- helpers the operator descriptors call ---------------------------------------------
    def action_index(self, actuation_ref):
        if not isinstance(actuation_ref, list):
            raise TypeError("actuation_ref must be the shared python list handed out by "
                            "FreeAssembly.get_actuation_reference")
        for idx, known in enumerate(self.action_lists):
            if known is actuation_ref:
                return idx
        self.action_lists.append(actuation_ref)
        return len(self.action_lists) - 1

    def add_forcing(self, rod_index, kind, action=-1, node=0, d=()):
        row = np.zeros(F_ND)
        row[:len(d)] = d
        self.forcing[rod_index].append((kind, action, node, row))

    def add_constraint(self, rod_index, kind, fixed_position, fixed_directors, k=0.0, nu=0.0):
        row = np.zeros(C_ND)
        row[0:3] = np.asarray(fixed_position, dtype=np.float64).reshape(3)
        row[3:12] = np.asarray(fixed_directors, dtype=np.float64).reshape(3, 3).ravel()
        row[12], row[13] = k, nu
        self.constraints[rod_index].append((kind, row))

    def set_damper(self, rod_index, c_t, ln_c_r):
        if rod_index in self.dampers:
            raise _ops.NotLowerable("more than one damper on a rod is not supported")
        self.dampers[rod_index] = (c_t, np.asarray(ln_c_r, dtype=np.float64))

    def _element(self, rod_index, idx):
        n = self.rods[rod_index].n_elems
        if not -n <= idx < n:
            raise IndexError("connection index %d out of range for a rod with %d elements" % (idx, n))
        return idx % n

    def _resolve(self, rod_index, elem):
        return self.director_source.get((rod_index, elem), (rod_index, elem))

    def slot(self, rod_index, elem):
        return self.slot0[rod_index] + elem

    def add_joint(self, kind, rod_one, index_one, rod_two, index_two, d, overwrite, torque):
        e1 = self._element(rod_one, index_one)
        e2 = self._element(rod_two, index_two)
        q1 = self._resolve(rod_one, e1)
        q2 = self._resolve(rod_two, e2)
        row = np.zeros(J_ND)
        row[:len(d)] = d
        self.joints.append((kind, self.slot(rod_one, e1), self.slot(rod_two, e2), self.slot(*q1), self.slot(*q2), row, torque))
        if overwrite:
            self.director_source[(rod_two, e2)] = q1


def lower_collection(collection):
    rods = list(collection._systems)
    if not rods:
        raise ValueError("nothing to integrate: no rod was appended")
    ctx = LoweringContext(rods)
    if ctx.n_slots > MAX_SLOTS:
        raise _ops.NotLowerable(
            "assembly has %d node slots; one thread-block cluster integrates one assembly and holds at most %d"
            % (ctx.n_slots, MAX_SLOTS))

    # this package's descriptors lower themselves; instances of the reference's own classes (no ``lower_*`` method) are
    # recognised by class name and read through their public attributes (br2/_adapter.py); anything else is refused
    from . import _adapter

    for rod_index, op in collection._forcing:
        _adapter.adapt(op, _adapter._FORCING, "forcing").lower_forcing(ctx, rod_index)
    for rod_index, op in collection._constraint_ops:
        _adapter.adapt(op, _adapter._CONSTRAINTS, "constraint").lower_constraint(ctx, rod_index)
    for rod_index, op in collection._damper_ops:
        _adapter.adapt(op, _adapter._DAMPERS, "damper").lower_damper(ctx, rod_index)
    for rod_one, rod_two, idx_one, idx_two, op in collection._joints:
        _adapter.adapt(op, _adapter._JOINTS, "joint").lower_joint(ctx, rod_one, idx_one, rod_two, idx_two)

    prog = Program()
    S = ctx.n_slots
    prog.n_slots = S
    slot_rod = np.zeros(S, dtype=np.int32)
    rod_i = np.zeros((len(rods), ROD_NI), dtype=np.int32)
    rod_d = np.zeros((len(rods), ROD_ND))
    slot_c = np.zeros((SC_COUNT, S))
    slot_c[SC_MASS] = 1.0
    slot_c[SC_REST_LEN] = 1.0
    slot_c[SC_VOLUME] = 1.0
    slot_c[SC_REST_VLEN] = 1.0
    slot_c[SC_J:SC_J + 3] = 1.0

    forcing_rows, cons_rows = [], []
    w_off = 0
    for r, rod in enumerate(rods):
        n = rod.n_elems
        s0 = ctx.slot0[r]
        slot_rod[s0:s0 + n + 1] = r
        fields, off = {}, w_off
        sizes = field_sizes(n)
        for short in STATE_ORDER:
            fields[short] = (off, sizes[short])
            off += sizes[short]
        prog.layout.append(fields)
        prog.n_elems.append(n)
        flags = 0
        f_begin = len(forcing_rows)
        forcing_rows.extend(ctx.forcing[r])
        c_begin = len(cons_rows)
        cons_rows.extend(ctx.constraints[r])
        rod_d[r, 0] = 1.0
        if r in ctx.dampers:
            flags |= ROD_HAS_DAMPER
            c_t, ln_c_r = ctx.dampers[r]
            rod_d[r, 0] = c_t
            slot_c[SC_LN_CR:SC_LN_CR + 3, s0:s0 + n] = ln_c_r
        rod_i[r] = (n, s0, w_off, flags, f_begin, len(forcing_rows), c_begin, len(cons_rows))
        w_off = off

        slot_c[SC_MASS, s0:s0 + n + 1] = rod.mass
        slot_c[SC_REST_LEN, s0:s0 + n] = rod.rest_lengths
        slot_c[SC_VOLUME, s0:s0 + n] = rod.volume
        slot_c[SC_SHEAR:SC_SHEAR + 3, s0:s0 + n] = _diagonal_only(rod.shear_matrix, "shear_matrix")
        slot_c[SC_J:SC_J + 3, s0:s0 + n] = _diagonal_only(rod.mass_second_moment_of_inertia, "mass_second_moment_of_inertia")
        slot_c[SC_JINV:SC_JINV + 3, s0:s0 + n] = _diagonal_only(rod.inv_mass_second_moment_of_inertia, "inv_mass_second_moment_of_inertia")
        slot_c[SC_REST_SIGMA:SC_REST_SIGMA + 3, s0:s0 + n] = rod.rest_sigma
        if n > 1:
            slot_c[SC_REST_VLEN, s0:s0 + n - 1] = rod.rest_voronoi_lengths
            slot_c[SC_BEND:SC_BEND + 3, s0:s0 + n - 1] = _diagonal_only(rod.bend_matrix, "bend_matrix")
            slot_c[SC_REST_KAPPA:SC_REST_KAPPA + 3, s0:s0 + n - 1] = rod.rest_kappa
    prog.words = w_off
    assert w_off == sum(rod_words(rod.n_elems) for rod in rods)
    prog.total_elements = int(sum(rod.n_elems for rod in rods))

    forcing_i = np.zeros((len(forcing_rows), F_NI), dtype=np.int32)
    forcing_d = np.zeros((len(forcing_rows), F_ND))
    for f, (kind, action, node, row) in enumerate(forcing_rows):
        forcing_i[f, :3] = (kind, action, node)
        forcing_d[f] = row
    cons_i = np.zeros((len(cons_rows), C_NI), dtype=np.int32)
    cons_d = np.zeros((len(cons_rows), C_ND))
    for c, (kind, row) in enumerate(cons_rows):
        cons_i[c, 0] = kind
        cons_d[c] = row

    NJ = len(ctx.joints)
    joint_i = np.zeros((NJ, J_NI), dtype=np.int32)
    joint_d = np.zeros((NJ, J_ND))
    node_lists = [[] for _ in range(S)]
    elem_lists = [[] for _ in range(S)]
    for j, (kind, s1, s2, q1, q2, row, torque) in enumerate(ctx.joints):
        joint_i[j, :5] = (kind, s1, s2, q1, q2)
        joint_d[j] = row
        node_lists[s1].append((j << 1) | 0)
        node_lists[s1 + 1].append((j << 1) | 0)
        node_lists[s2].append((j << 1) | 1)
        node_lists[s2 + 1].append((j << 1) | 1)
        if torque:
            elem_lists[s1].append((j << 1) | 0)
            elem_lists[s2].append((j << 1) | 1)

    def csr(lists):
        ptr = np.zeros(S + 1, dtype=np.int32)
        ptr[1:] = np.cumsum([len(items) for items in lists])
        flat = [item for items in lists for item in items]
        return ptr, np.array(flat, dtype=np.int32).reshape(-1)

    node_ptr, node_items = csr(node_lists)
    elem_ptr, elem_items = csr(elem_lists)
    overwrite_src = np.full(S, -1, dtype=np.int32)
    for (rod_index, elem), src in ctx.director_source.items():
        overwrite_src[ctx.slot(rod_index, elem)] = ctx.slot(*src)

    prog.action_lists = ctx.action_lists
    fast = _lower_fast(ctx, rods, rod_i, slot_c, rod_d, forcing_rows, cons_rows, node_lists, elem_lists, overwrite_src)
    prog.tables = dict(
        slot_rod=slot_rod, rod_i=rod_i, rod_d=rod_d, slot_c=slot_c, forcing_i=forcing_i, forcing_d=forcing_d,
        cons_i=cons_i, cons_d=cons_d, joint_i=joint_i, joint_d=joint_d, node_ptr=node_ptr,
        node_items=node_items, elem_ptr=elem_ptr, elem_items=elem_items, overwrite_src=overwrite_src,
        fast_i=fast["fast_i"], fast_c=fast["fast_c"], act_i=fast["act_i"], act_d=fast["act_d"],
    )
    prog.fast = dict(ok=fast["ok"], uniform=fast["uniform"], point_forces=fast["point_forces"], why=fast["why"],
                     ramp_up_time=fast["ramp_up_time"],
                     n_act=len(fast["act_i"]))
    prog.tables = {k: np.ascontiguousarray(v) for k, v in prog.tables.items()}
    prog.counts = dict(n_slots=S, n_rods=len(rods), n_forcing=len(forcing_rows), n_constraints=len(cons_rows),
                       n_joints=NJ, n_actions=len(ctx.action_lists))
    return prog


def _lower_fast(ctx, rods, rod_i, slot_c, rod_d, forcing_rows, cons_rows, node_lists, elem_lists, overwrite_src):
    """Per-slot descriptors of the fast kernel (csrc/br2_fast.cuh), or ``ok = False`` with the reason
    when the assembly needs the generic kernel."""
    S = ctx.n_slots
    fast_i = np.full((S, FAST_NI), -1, dtype=np.int32)
    fast_i[:, FAST_FLAGS] = 0
    fast_i[:, FAST_ACT_BEGIN] = 0
    fast_i[:, FAST_ACT_END] = 0
    fast_c = np.zeros((FC_COUNT, S))
    act_i, act_d = [], []
    out = dict(ok=False, uniform=False, point_forces=False, why="", ramp_up_time=1.0, fast_i=fast_i, fast_c=fast_c,
               act_i=np.zeros(0, dtype=np.int32), act_d=np.zeros((0, 3)))

    def refuse(reason):
        out["why"] = reason
        return out

    if slot_c[SC_REST_SIGMA:SC_REST_SIGMA + 3].any() or slot_c[SC_REST_KAPPA:SC_REST_KAPPA + 3].any():
        return refuse("non-zero rest strains")
    ramps = {float(row[1]) for kind, _, _, row in forcing_rows if kind in (FORCE_TWIST, FORCE_BEND)}
    # point forces stay in the generic forcing table; the tile kernel serves them as "extra" forces, the
    # one-slot-per-thread fast kernel does not know them (libbr2cuda then picks the tile or the generic kernel)
    out["point_forces"] = any(kind == FORCE_POINT for kind, _, _, _ in forcing_rows)
    if len(ramps) > 1:
        return refuse("actuation rows with different ramp_up_time")
    out["ramp_up_time"] = ramps.pop() if ramps else 1.0

    # ---- per-slot constants (element values; undefined slots copy a defined neighbour so that a
    #      uniform assembly stays uniform) ------------------------------------------------------
    for r, rod in enumerate(rods):
        n, s0 = rod.n_elems, int(rod_i[r, 1])
        cols = slice(s0, s0 + n + 1)

        def spread(values, count):  # [.., count] element/voronoi values -> all n + 1 slots of the rod
            values = np.asarray(values, dtype=np.float64)
            if count == 0:
                return np.ones(values.shape[:-1] + (n + 1,))
            pad = np.repeat(values[..., -1:], n + 1 - count, axis=-1)
            return np.concatenate([values[..., :count], pad], axis=-1)

        fast_c[FC_REST_LEN, cols] = spread(rod.rest_lengths, n)
        fast_c[FC_INV_REST_LEN, cols] = 1.0 / fast_c[FC_REST_LEN, cols]
        fast_c[FC_VOL_OVER_PI, cols] = spread(rod.volume, n) / np.pi
        fast_c[FC_SHEAR:FC_SHEAR + 3, cols] = spread(slot_c[SC_SHEAR:SC_SHEAR + 3, s0:s0 + n], n)
        fast_c[FC_J:FC_J + 3, cols] = spread(slot_c[SC_J:SC_J + 3, s0:s0 + n], n)
        fast_c[FC_JINV:FC_JINV + 3, cols] = spread(slot_c[SC_JINV:SC_JINV + 3, s0:s0 + n], n)
        ln_cr = spread(slot_c[SC_LN_CR:SC_LN_CR + 3, s0:s0 + n], n)
        fast_c[FC_LN_CR:FC_LN_CR + 3, cols] = ln_cr
        fast_c[FC_CR:FC_CR + 3, cols] = np.exp(ln_cr)
        fast_c[FC_REST_VLEN, cols] = spread(rod.rest_voronoi_lengths, n - 1)
        fast_c[FC_INV_REST_VLEN, cols] = 1.0 / fast_c[FC_REST_VLEN, cols]
        fast_c[FC_BEND:FC_BEND + 3, cols] = spread(slot_c[SC_BEND:SC_BEND + 3, s0:s0 + n - 1], n - 1) if n > 1 else 0.0
        fast_c[FC_CT, cols] = rod_d[r, 0]

    # ---- forcing ----------------------------------------------------------------------------
    row_cursor = 0
    for r, rod in enumerate(rods):
        n, s0 = rod.n_elems, int(rod_i[r, 1])
        gravity = np.zeros(3)
        twist_rows, bend_rows = [], []
        for kind, action, node, row in ctx.forcing[r]:
            if kind == FORCE_GRAVITY:
                gravity += row[:3]
            elif kind == FORCE_TWIST:   # (P * scale * ramp) / n about d3, every element
                twist_rows.append((action, np.array([0.0, 0.0, row[0] / n])))
            elif kind == FORCE_BEND:    # P * scale * ramp * [cos, sin, 0], last element only
                bend_rows.append((action, np.array([row[0] * row[2], row[0] * row[3], 0.0])))
        fast_c[FC_GRAVITY:FC_GRAVITY + 3, s0:s0 + n + 1] = gravity[:, None]
        if twist_rows:
            begin = len(act_i)
            for action, coeff in twist_rows:
                act_i.append(action)
                act_d.append(coeff)
            fast_i[s0:s0 + n - 1, FAST_ACT_BEGIN] = begin
            fast_i[s0:s0 + n - 1, FAST_ACT_END] = len(act_i)
        # the last element sees the twist rows followed by the bend rows (registration order is
        # alpha, beta, gamma fibres -- free_simulator.py:371-374 -- but sums commute up to rounding)
        begin = len(act_i)
        for action, coeff in twist_rows + bend_rows:
            act_i.append(action)
            act_d.append(coeff)
        fast_i[s0 + n - 1, FAST_ACT_BEGIN] = begin
        fast_i[s0 + n - 1, FAST_ACT_END] = len(act_i)

    # ---- constraints --------------------------------------------------------------------------
    cons_cursor = 0
    for r, rod in enumerate(rods):
        n, s0 = rod.n_elems, int(rod_i[r, 1])
        c_begin = int(rod_i[r, 6])
        kinds = [kind for kind, _ in ctx.constraints[r]]
        if len(kinds) != len(set(kinds)):
            return refuse("two constraints of the same kind on one rod")
        for offset, (kind, row) in enumerate(ctx.constraints[r]):
            if kind == CONS_SOFT_FIXED_BASE:
                fast_i[s0, FAST_FLAGS] |= PIN_DIRECTOR | ZERO_NODE_RATE | ZERO_ELEM_RATE | ZERO_REPORTED
                fast_i[s0, FAST_SOFTFIX_ROW] = c_begin + offset
                fast_c[FC_PIN + 3:FC_PIN + 12, s0] = row[3:12]
            else:
                fast_i[s0 + n, FAST_FLAGS] |= PIN_POSITION | ZERO_NODE_RATE
                fast_c[FC_PIN:FC_PIN + 3, s0 + n] = row[0:3]
                if fast_i[s0 + n - 1, FAST_FLAGS] & PIN_DIRECTOR:
                    return refuse("two director pins on one element")
                fast_i[s0 + n - 1, FAST_FLAGS] |= PIN_DIRECTOR | ZERO_ELEM_RATE
                fast_c[FC_PIN + 3:FC_PIN + 12, s0 + n - 1] = row[3:12]

    # ---- joints -------------------------------------------------------------------------------
    # "own" joints: surface joints whose rod-two element evaluates them; each element may be rod two of
    # one and rod one of one such joint (the BR2 ring) -- their forces / torques are gathered implicitly.
    # Everything else (tip-to-base joints, surface joints outside that pattern) is an "extra" joint with
    # a compact id, spread over the threads, gathered through short per-slot lists.
    extra = []
    jparams = None
    is_rod_one = np.zeros(S, dtype=bool)
    for j, (kind, s1, s2, q1, q2, row, torque) in enumerate(ctx.joints):
        ownable = (kind == JOINT_SURFACE and q1 == s1 and q2 == s2 and s1 != s2
                   and fast_i[s2, FAST_OWN_JOINT] < 0 and not is_rod_one[s1])
        if ownable:
            fast_i[s2, FAST_OWN_JOINT] = j
            is_rod_one[s1] = True
            fast_c[FC_JK, s2], fast_c[FC_JNU, s2], fast_c[FC_JKREP, s2] = row[0], row[1], row[2]
            jparams = (row[0], row[1], row[2]) if jparams is None else jparams
        else:
            extra.append(j)
    if jparams is not None:  # slots without an own joint copy the common parameters (keeps uniformity)
        none = fast_i[:, FAST_OWN_JOINT] < 0
        fast_c[FC_JK, none], fast_c[FC_JNU, none], fast_c[FC_JKREP, none] = jparams
    if len(extra) > FAST_MAX_EXTRA:
        return refuse("more than %d joints outside the ring pattern" % FAST_MAX_EXTRA)
    # spread the extra joints over the threads, one warp after another
    warps = (S + 31) // 32
    candidates = [w * 32 + lane for lane in range(32) for w in range(warps) if w * 32 + lane < S]
    extra_id = {}
    for xid, (j, slot) in enumerate(zip(extra, candidates)):
        fast_i[slot, FAST_EXTRA_JOINT] = j
        fast_i[slot, FAST_EXTRA_ID] = xid
        extra_id[j] = xid
    x_node = [[] for _ in range(S)]
    x_elem = [[] for _ in range(S)]
    for j in extra:
        kind, s1, s2, q1, q2, row, torque = ctx.joints[j]
        xid = extra_id[j]
        for slot, sign in ((s1, 0), (s1 + 1, 0), (s2, 1), (s2 + 1, 1)):
            x_node[slot].append((xid << 1) | sign)
        if torque:
            x_elem[s1].append((xid << 1) | 0)
            x_elem[s2].append((xid << 1) | 1)
    for slot in range(S):
        if len(x_node[slot]) > FAST_NODE_ITEMS:
            return refuse("a node receives more than %d extra joint forces" % FAST_NODE_ITEMS)
        if len(x_elem[slot]) > 2:
            return refuse("an element receives more than 2 extra joint torques")
        fast_i[slot, FAST_NODE_ITEM0:FAST_NODE_ITEM0 + len(x_node[slot])] = x_node[slot]
        fast_i[slot, FAST_ELEM_ITEM0:FAST_ELEM_ITEM0 + len(x_elem[slot])] = x_elem[slot]

    # ---- uniformity: every slot's element constants equal slot 0's (kernel then uses the constant bank)
    ref = fast_c[:FC_PIN, :1]
    scale = np.maximum(np.abs(ref), 1e-300)
    out["uniform"] = bool((np.abs(fast_c[:FC_PIN] - ref) <= UNIFORM_RTOL * scale).all())
    out["ok"] = True
    out["act_i"] = np.array(act_i, dtype=np.int32).reshape(-1)
    out["act_d"] = np.array(act_d, dtype=np.float64).reshape(-1, 3)
    return out
