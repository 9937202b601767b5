"""Host logic of the product (no GPU): the assembly builder and the lowering to device tables,
checked against (a) the registration logs the reference's REAL builder produced
(tests/golden/registration_*.json) and (b) the oracle's own rods."""
import json
import os

import numpy as np
import pytest

from conftest import DATA, GOLDEN
from golden.cases import TRAJECTORIES

from br2 import _lowering as L
from br2.free_simulator import FreeAssembly
from oracle import br2_port

ASSEMBLY = {k: os.path.join(DATA, "assembly_%s.json" % k) for k in ("single", "double", "triple")}
ROD_DB = os.path.join(DATA, "rod_library.json")


class FakeEnv:
    time_step = 2.0e-5


def build_product(assembly, **kwargs):
    assy = FreeAssembly(FakeEnv(), **kwargs)
    assy.build(ROD_DB, ASSEMBLY[assembly], verbose=False)
    assy.generate_callbacks(2000)
    program = assy.simulator.lower()
    return assy, program


@pytest.mark.parametrize("case", TRAJECTORIES, ids=[c[0] for c in TRAJECTORIES])
def test_registration_matches_reference_builder(case):
    name, assembly, kwargs, action, duration, fps = case
    want = json.load(open(os.path.join(GOLDEN, "registration_%s.json" % name)))
    assy, program = build_product(assembly, **kwargs)
    sim = assy.simulator
    assert len(sim) == want["n_systems"]

    got_forcing = [(r, type(op).__name__) for r, op in sim._forcing]
    assert got_forcing == [(e["rod"], e["cls"]) for e in want["forcing"]]
    for (r, op), e in zip(sim._forcing, want["forcing"]):
        for key in ("acc_gravity", "z_angle", "magnitude_scale", "scale", "ramp_up_time"):
            if key in e:
                np.testing.assert_array_equal(np.asarray(getattr(op, key)), np.asarray(e[key]), err_msg=key)

    assert [(r, type(op).__name__) for r, op in sim._constraint_ops] == [(e["rod"], e["cls"]) for e in want["constraints"]]
    for (r, op), e in zip(sim._constraint_ops, want["constraints"]):
        assert (op.k, op.nu, op.kt) == (e["k"], e["nu"], e["kt"])
        np.testing.assert_array_equal(op.fixed_position, e["fixed_position"])
        np.testing.assert_array_equal(op.fixed_directors, e["fixed_directors"])

    assert [r for r, _ in sim._damper_ops] == [e["rod"] for e in want["dampers"]]
    for (r, op), e in zip(sim._damper_ops, want["dampers"]):
        np.testing.assert_array_equal(op.translational_damping_coefficient, e["translational"])
        np.testing.assert_array_equal(op.rotational_damping_coefficient, np.asarray(e["rotational"]))

    assert len(sim._joints) == len(want["connections"])
    for (r1, r2, i1, i2, op), e in zip(sim._joints, want["connections"]):
        assert (r1, r2, i1, i2, type(op).__name__) == (e["rod_one"], e["rod_two"], e["idx_one"], e["idx_two"], e["cls"])
        assert float(op.k) == e["k"] and float(op.nu) == e["nu"]
        if e["cls"] == "SurfaceJointSideBySide":
            assert float(op.k_repulsive) == e["k_repulsive"]
            np.testing.assert_array_equal(op.rod_one_direction_vec_in_material_frame, e["dv1"])
            np.testing.assert_array_equal(op.rod_two_direction_vec_in_material_frame, e["dv2"])
            assert float(op.offset_btw_rods) == e["offset"]
        else:
            np.testing.assert_array_equal(op.rod1_rd2_local, e["l1"])
            np.testing.assert_array_equal(op.rod2_rd2_local, e["l2"])


@pytest.mark.parametrize("assembly", ["single", "double", "triple"])
def test_slot_constants_match_oracle_rods(assembly):
    assy, program = build_product(assembly, gravity=True)
    ref = br2_port.OracleEnvironment()
    ref.reset(ROD_DB, ASSEMBLY[assembly], gravity=True)
    t = program.tables
    S = program.n_slots
    assert t["slot_c"].shape == (L.SC_COUNT, S)
    for r, rod in enumerate(ref.simulator):
        n = rod.n_elems
        assert tuple(t["rod_i"][r, :2]) == (n, sum(x.n_elems + 1 for x in list(ref.simulator)[:r]))
        s0 = t["rod_i"][r, 1]
        sc = t["slot_c"]
        np.testing.assert_array_equal(sc[L.SC_MASS, s0:s0 + n + 1], rod.mass)
        np.testing.assert_array_equal(sc[L.SC_REST_LEN, s0:s0 + n], rod.rest_lengths)
        np.testing.assert_array_equal(sc[L.SC_VOLUME, s0:s0 + n], rod.volume)
        np.testing.assert_array_equal(sc[L.SC_REST_VLEN, s0:s0 + n - 1], rod.rest_voronoi_lengths)
        for axis in range(3):
            np.testing.assert_array_equal(sc[L.SC_SHEAR + axis, s0:s0 + n], rod.shear_matrix[axis, axis])
            np.testing.assert_array_equal(sc[L.SC_BEND + axis, s0:s0 + n - 1], rod.bend_matrix[axis, axis])
            np.testing.assert_array_equal(sc[L.SC_J + axis, s0:s0 + n], rod.mass_second_moment_of_inertia[axis, axis])
            np.testing.assert_array_equal(sc[L.SC_JINV + axis, s0:s0 + n], rod.inv_mass_second_moment_of_inertia[axis, axis])
        # initial state template and the constructor-time geometry refresh (radius recomputed!)
        prod_rod = assy.simulator[r]
        np.testing.assert_array_equal(prod_rod.position_collection, rod.position_collection)
        np.testing.assert_array_equal(prod_rod.director_collection, rod.director_collection)
        np.testing.assert_array_equal(prod_rod.radius, rod.radius)
        np.testing.assert_array_equal(prod_rod.lengths, rod.lengths)
        np.testing.assert_array_equal(prod_rod.dilatation, rod.dilatation)
    damp = {r: d for r, d in ref.simulator._dampers}
    for r in range(len(ref.simulator)):
        n, s0 = t["rod_i"][r, 0], t["rod_i"][r, 1]
        assert t["rod_i"][r, 3] & L.ROD_HAS_DAMPER
        assert t["rod_d"][r, 0] == damp[r].translational_damping_coefficient
        np.testing.assert_array_equal(t["slot_c"][L.SC_LN_CR:L.SC_LN_CR + 3, s0:s0 + n],
                                      np.log(damp[r].rotational_damping_coefficient))


def test_single_tables_structure():
    assy, program = build_product("single", gravity=True)
    t, c = program.tables, program.counts
    assert c == dict(n_slots=126, n_rods=3, n_forcing=6, n_constraints=3, n_joints=123, n_actions=3)
    assert program.words == 3 * (15 * 42 + 24 * 41) and program.total_elements == 123
    # ring of pairs (2,0),(0,1),(1,2), element-major inside a pair
    ji = t["joint_i"]
    assert list(ji[0, :5]) == [0, 2 * 42, 0, 2 * 42, 0] and list(ji[41, :5]) == [0, 0, 42, 0, 42]
    assert list(ji[122, :5]) == [0, 42 + 40, 84 + 40, 42 + 40, 84 + 40]
    # every element is in exactly two joints; interior nodes receive four half-forces, end nodes two
    assert set(np.diff(t["elem_ptr"])[np.r_[0:41, 42:83, 84:125]]) == {2}
    counts = np.diff(t["node_ptr"])
    assert counts[0] == 2 and counts[41] == 2 and set(counts[1:41]) == {4}
    assert (t["overwrite_src"] == -1).all()
    # forcing: per rod gravity first, then the fibre; actions in first-seen order
    fi = t["forcing_i"]
    assert [tuple(row[:2]) for row in fi] == [(0, -1), (2, 0), (0, -1), (1, 1), (0, -1), (1, 2)]
    assert [assy.actuation[k] is lst for k, lst in zip(("action1", "action2", "action3"), program.action_lists)] == [True] * 3


def test_serial_joint_overwrite_chain_is_resolved():
    assy, program = build_product("triple")
    t = program.tables
    S0 = [int(x) for x in t["rod_i"][:, 1]]
    n = 41
    # segment 2/3 base elements end up with the director/omega of the LAST rod of the previous segment
    for seg, prev_last in ((1, 2), (2, 5)):
        for b in range(3):
            assert t["overwrite_src"][S0[3 * seg + b]] == S0[prev_last] + n - 1
    tips = t["joint_i"][t["joint_i"][:, 0] == 1]
    assert len(tips) == 18
    # pair (a, b) of the first interface: rod two's director is its own for a == 0, else prev rod a-1's tip
    first = tips[:9].reshape(3, 3, -1)
    for a in range(3):
        for b in range(3):
            row = first[a, b]
            assert row[1] == S0[a] + n - 1 and row[2] == S0[3 + b] and row[3] == S0[a] + n - 1
            assert row[4] == (S0[3 + b] if a == 0 else S0[a - 1] + n - 1)
    # fast path: the 18 tip joints are "extra" joints with compact ids, one per thread, spread over warps
    fi = t["fast_i"]
    assert program.fast["ok"] and program.fast["uniform"]
    extra_rows = fi[fi[:, L.FAST_EXTRA_JOINT] >= 0]
    assert sorted(extra_rows[:, L.FAST_EXTRA_ID]) == list(range(18))
    assert set(t["joint_i"][extra_rows[:, L.FAST_EXTRA_JOINT], 0]) == {1}
    assert (fi[:, L.FAST_OWN_JOINT] >= 0).sum() == 9 * n  # every element owns its ring joint
    # tip joints add no torque rows; unactuated rods of segment 3 have no fibre forcing
    assert np.diff(t["elem_ptr"]).max() == 2
    seg3 = t["rod_i"][6:9]
    assert [int(r[5] - r[4]) for r in seg3] == [0, 0, 1]


def test_unknown_operator_is_refused():
    from br2.hooks import NoForces, NotLowerable

    class Custom(NoForces):
        def apply_forces(self, system, time=0.0):
            system.external_forces[0] += 1.0

    assy = FreeAssembly(FakeEnv())
    assy.build(ROD_DB, ASSEMBLY["single"], verbose=False)
    assy.simulator.add_forcing_to(assy.simulator[0]).using(Custom)
    with pytest.raises(NotLowerable):
        assy.simulator.lower()


def test_high_resolution_triple_lowers_and_too_large_assembly_is_refused(tmp_path):
    """BASELINE config 4 (triple assembly, 200 elements per rod = 1809 node slots) lowers -- the library runs it
    as a cluster of three CTAs, one per segment; an assembly beyond a cluster's capacity is refused."""
    from br2.hooks import NotLowerable

    assy = FreeAssembly(FakeEnv())
    assy.build(os.path.join(DATA, "rod_library_n200.json"), ASSEMBLY["triple"], verbose=False)
    program = assy.simulator.lower()
    assert program.n_slots == 1809 and program.fast["ok"] and program.fast["uniform"]

    library = json.load(open(os.path.join(DATA, "rod_library_n200.json")))
    library["DefaultParams"]["n_elements"] = 1000
    path = tmp_path / "rod_library_n1000.json"
    path.write_text(json.dumps(library))
    assy = FreeAssembly(FakeEnv())
    assy.build(str(path), ASSEMBLY["triple"], verbose=False)
    with pytest.raises(NotLowerable, match="9009 node slots"):
        assy.simulator.lower()


def test_set_actuation_unknown_action_raises_like_reference():
    assy, _ = build_product("single")
    with pytest.raises(IndexError):
        assy.set_actuation({"no_such_action": 1.0})


def test_kernel_plan_of_the_baseline_assemblies(tmp_path):
    """Which kernel libbr2cuda picks and how it lays the assembly out (host-only entry point br2_describe_plan)."""
    from br2 import _native

    assert _native.describe_plan(build_product("single", gravity=True)[1]) == \
        "tile: 1 CTA(s) per instance of 63 threads (2 slots per thread), 0 tip-to-base joint(s) / point force(s)"
    assert _native.describe_plan(build_product("double")[1]) == \
        "tile: 1 CTA(s) per instance of 126 threads (2 slots per thread), 9 tip-to-base joint(s) / point force(s)"
    assert _native.describe_plan(build_product("triple")[1]) == \
        "tile: 1 CTA(s) per instance of 189 threads (2 slots per thread), 18 tip-to-base joint(s) / point force(s)"
    # BASELINE config 4: 200 elements per rod -> one CTA per segment, a cluster of three
    assy = FreeAssembly(FakeEnv())
    assy.build(os.path.join(DATA, "rod_library_n200.json"), ASSEMBLY["triple"], verbose=False)
    assert _native.describe_plan(assy.simulator.lower()) == \
        "tile: 3 CTA(s) per instance of 303, 303, 303 threads (2 slots per thread), 18 tip-to-base joint(s) / point force(s)"
    # a segment too long for one CTA of the cluster kernel is refused with the reason
    library = json.load(open(os.path.join(DATA, "rod_library_n200.json")))
    library["DefaultParams"]["n_elements"] = 250
    path = tmp_path / "rod_library_n250.json"
    path.write_text(json.dumps(library))
    assy = FreeAssembly(FakeEnv())
    assy.build(str(path), ASSEMBLY["triple"], verbose=False)
    with pytest.raises(_native.Br2Error, match="needs 378 threads"):
        _native.describe_plan(assy.simulator.lower())


def test_ragged_assembly_lowers_to_the_tile_kernel_with_classes_of_constants(tmp_path):
    """Segments whose rods differ in element count and length: rest lengths are no longer
    uniform over the assembly, but they are over every rod: the tile plan groups the rods into classes of equal constants
    (two here) and the kernels read the class tables from shared memory.  A rod whose elements differ among themselves
    still goes to the one-slot-per-thread kernel with per-slot constants."""
    from br2 import _native

    library = json.load(open(ROD_DB))
    for key in ("RodBend", "RodRightTwist", "RodLeftTwist"):
        library["Rods"][key + "Short"] = dict(library["Rods"][key], n_elements=30, base_length=0.12)
    assembly = json.load(open(ASSEMBLY["double"]))
    assembly["Segments"]["seg2"]["rod_order"] = ["RodRightTwistShort", "RodBendShort", "RodLeftTwistShort"]
    rod_db, path = tmp_path / "rod_library_ragged.json", tmp_path / "assembly_ragged.json"
    rod_db.write_text(json.dumps(library))
    path.write_text(json.dumps(assembly))
    assy = FreeAssembly(FakeEnv())
    assy.build(str(rod_db), str(path), verbose=False)
    program = assy.simulator.lower()
    assert program.n_slots == 3 * 42 + 3 * 31 and program.fast["ok"] and not program.fast["uniform"]
    assert _native.describe_plan(program) == ("tile: 1 CTA(s) per instance of 111 threads (2 slots per thread), 9 tip-to-base "
                                              "joint(s) / point force(s), 2 classes of rod constants (shared-memory copy)")
    # a rod that is not uniform in itself (a tapered rod): element constants vary along it
    tapered = assy.simulator.lower()
    tapered.tables["fast_c"][0, 5] *= 1.01  # rest length of rod 0's sixth element
    assert _native.describe_plan(tapered) == \
        "fast: one slot per thread (tile kernel refused: element constants vary along a rod)"
