/*
 * br2cuda.h -- C ABI of libbr2cuda.so: the B200 (sm_100a) replacement for the rod-integration hot
 * path of BR2-simulator.
 *
 * What it replaces.  In the reference the path is entered from exactly one place, the loop
 *     while time < self.time + duration:
 *         time = self.do_step(self.StatefulStepper, self.stages_and_updates, self.simulator, time, dt)
 * (/root/reference/br2/environment.py:278-285), where `do_step` comes from PyElastica's
 * `extend_stepper_interface(PositionVerlet(), simulator)` (environment.py:221-223) and `simulator`
 * is the system collection filled by `FreeAssembly.build` (/root/reference/br2/free_simulator.py:101-268)
 * through the hooks append / dampen / constrain / add_forcing_to / connect / collect_diagnostics /
 * finalize.  There is no FFI in the reference: the seam is the Python callable above.  The drop-in
 * therefore is (1) a `br2` Python package with the same classes and hooks whose `finalize()` lowers
 * the registered operators to the flat tables below, and (2) this library, which a maintainer binds
 * with ctypes (INTEGRATION.md shows the stub) -- or from any language with a C FFI.
 *
 * Conventions: every entry point returns 0 on success or a negative BR2_E* code; the message is
 * available from br2_last_error() (thread-local).  No exceptions or C++ types cross the boundary.
 * Device pointers are BORROWED (typically PyTorch-owned); the library never frees them.  One host
 * thread drives one handle; handles on different GPUs are independent.
 */
#ifndef BR2CUDA_H
#define BR2CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BR2_ABI_VERSION 2

enum {
    BR2_OK = 0,
    BR2_EINVAL = -1,   /* malformed program / argument */
    BR2_ECUDA = -2,    /* CUDA runtime error (message holds cudaGetErrorString) */
    BR2_ELIMIT = -3,   /* assembly does not fit the kernel's resource limits */
    BR2_ESTATE = -4    /* call order error (e.g. step before bind) */
};

/* ---- table layout (what finalize() lowers the registered operators to) ---------------------
 *
 * A "slot" is one CUDA thread's share of an assembly: node k and (for k < n) element k and
 * (for k < n-1) Voronoi cell k of one rod.  Rod r with n elements owns slots slot0 .. slot0+n.
 * All tables are host arrays; br2_create copies them to the device.
 */

/* rod_i: [n_rods][BR2_ROD_NI] int32 */
enum {
    BR2_ROD_N = 0,        /* elements */
    BR2_ROD_SLOT0 = 1,    /* first slot */
    BR2_ROD_WOFF = 2,     /* offset (doubles) of the rod's block inside one instance's state */
    BR2_ROD_FLAGS = 3,    /* BR2_ROD_HAS_DAMPER */
    BR2_ROD_F_BEGIN = 4,  /* forcing rows [begin, end) of this rod, in execution order */
    BR2_ROD_F_END = 5,
    BR2_ROD_C_BEGIN = 6,  /* constraint rows [begin, end) of this rod, in execution order */
    BR2_ROD_C_END = 7,
    BR2_ROD_NI = 8
};
enum { BR2_ROD_HAS_DAMPER = 1 };
/* rod_d: [n_rods][BR2_ROD_ND] double : translational damping factor exp(-nu dt) */
enum { BR2_ROD_CT = 0, BR2_ROD_ND = 2 };

/* slot_c: [BR2_SC_COUNT][n_slots] double, per-slot constants (PyElastica CosseratRod arrays,
 * SURVEY.md Appendix A.2; only the diagonals of the constitutive matrices are supported) */
enum {
    BR2_SC_MASS = 0,       /* node */
    BR2_SC_REST_LEN = 1,   /* element */
    BR2_SC_VOLUME = 2,
    BR2_SC_SHEAR = 3,      /* 3: diag(shear_matrix) */
    BR2_SC_J = 6,          /* 3: diag(mass_second_moment_of_inertia) */
    BR2_SC_JINV = 9,       /* 3: diag(inv_mass_second_moment_of_inertia) */
    BR2_SC_LN_CR = 12,     /* 3: ln(rotational_damping_coefficient) of AnalyticalLinearDamper */
    BR2_SC_REST_SIGMA = 15,/* 3 */
    BR2_SC_REST_VLEN = 18, /* Voronoi */
    BR2_SC_BEND = 19,      /* 3: diag(bend_matrix) */
    BR2_SC_REST_KAPPA = 22,/* 3 */
    BR2_SC_COUNT = 25
};

/* forcing_i: [n_forcing][BR2_F_NI] int32 ; forcing_d: [n_forcing][BR2_F_ND] double */
enum { BR2_F_TYPE = 0, BR2_F_ACTION = 1, BR2_F_NODE = 2, BR2_F_NI = 4, BR2_F_ND = 6 };
enum {
    BR2_FORCE_GRAVITY = 0, /* d = {g[3]}                       elastica GravityForces (free_simulator.py:334-338) */
    BR2_FORCE_TWIST = 1,   /* d = {scale, ramp_up_time}        FreeTwistActuation (free_custom_systems.py:96-101) */
    BR2_FORCE_BEND = 2,    /* d = {scale, ramp_up_time, cos z, sin z}  FreeBendActuation (free_custom_systems.py:49-57) */
    BR2_FORCE_POINT = 3    /* d = {F[3], ramp_up_time}, node   PointForces (custom_activation.py:46-81) */
};

/* cons_i: [n_constraints][BR2_C_NI] int32 ; cons_d: [n_constraints][BR2_C_ND] double
 * d = {fixed_position[3], fixed_directors[9] row-major, k, nu} */
enum { BR2_C_TYPE = 0, BR2_C_NI = 2, BR2_C_ND = 14 };
enum {
    BR2_CONS_SOFT_FIXED_BASE = 0, /* FreeBaseEndSoftFixed (free_custom_systems.py:104-185) */
    BR2_CONS_LAST_END_FIXED = 1   /* LastEndFixedRod (custom_constraint.py:9-88) */
};

/* joint_i: [n_joints][BR2_J_NI] int32 ; joint_d: [n_joints][BR2_J_ND] double, registration order.
 * S1/S2: element slots of rod one / rod two; Q1/Q2: element slots whose (start-of-synchronize)
 * directors the joint sees -- differ from S1/S2 only downstream of a tip-to-base joint's
 * director overwrite (legacy_surface_connection_parallel_rod_numba.py:474-478). */
enum { BR2_J_TYPE = 0, BR2_J_S1 = 1, BR2_J_S2 = 2, BR2_J_Q1 = 3, BR2_J_Q2 = 4, BR2_J_NI = 8, BR2_J_ND = 10 };
enum {
    BR2_JOINT_SURFACE = 0,    /* d = {k, nu, k_repulsive, dv1[3], dv2[3], offset}  elastica SurfaceJointSideBySide */
    BR2_JOINT_TIP_TO_BASE = 1 /* d = {k, nu, 0, l1[3], l2[3], 0}                   TipToTipStraightJoint (legacy...:346-478) */
};

/* ---- fast-path tables -------------------------------------------------------------------------
 * The BR2 assembly structure (every operator of SURVEY.md section 8a rows A5-A12) is served by a
 * specialised kernel that keeps per-slot descriptors in registers / shared memory instead of walking
 * the generic tables every step.  finalize() fills these when the assembly qualifies (fast_ok = 1):
 *   - forcing rows are gravity / twist / bend (all with the same ramp_up_time) and point forces (tile kernel only);
 *   - rest_sigma and rest_kappa are zero;
 *   - every element is "rod two" of at most one surface joint and "rod one" of at most one (those
 *     contributions are gathered implicitly); at most BR2_FAST_MAX_EXTRA further joints (tip-to-base
 *     joints, surface joints outside that pattern), of which a node receives at most
 *     BR2_FAST_NODE_ITEMS half-forces and an element at most two torques.
 * Otherwise the generic table-interpreting kernel runs.  Same results either way (tests compare them).
 */
/* fast_i: [n_slots][BR2_FAST_NI] int32 */
enum {
    BR2_FAST_FLAGS = 0,
    BR2_FAST_OWN_JOINT = 1,    /* joint row this slot's element evaluates as rod two, or -1 */
    BR2_FAST_EXTRA_JOINT = 2,  /* one more joint row evaluated by this thread from shared memory, or -1 */
    BR2_FAST_EXTRA_ID = 16,    /* compact id (< BR2_FAST_MAX_EXTRA) of that joint; gather items refer to it */
    BR2_FAST_SOFTFIX_ROW = 3,  /* constraint row of a soft-fixed base on this slot (node 0), or -1 */
    BR2_FAST_ACT_BEGIN = 4,    /* actuation rows [begin, end) acting on this slot's element */
    BR2_FAST_ACT_END = 5,
    BR2_FAST_ELEM_ITEM0 = 6,   /* 2 torque gather items ((extra id << 1) | side) of extra joints, -1 = none */
    BR2_FAST_NODE_ITEM0 = 8,   /* BR2_FAST_NODE_ITEMS force gather items ((extra id << 1) | sign) of extra joints */
    BR2_FAST_NI = 18
};
enum { BR2_FAST_NODE_ITEMS = 8, BR2_FAST_MAX_EXTRA = 64 };
enum {
    BR2_FAST_PIN_DIRECTOR = 1,  /* element director is reset to a constant every half step */
    BR2_FAST_PIN_POSITION = 2,  /* node position is reset to a constant every half step */
    BR2_FAST_ZERO_NODE_RATE = 4,/* node velocity (and reported acceleration) zeroed in constrain_rates */
    BR2_FAST_ZERO_ELEM_RATE = 8,/* element omega (and reported alpha) zeroed in constrain_rates */
    BR2_FAST_ZERO_REPORTED = 16 /* soft-fixed base also zeroes the reported a / alpha of slot 0 */
};
/* fast_c: [BR2_FC_COUNT][n_slots] double.  Element / Voronoi constants with reciprocals precomputed. */
enum {
    BR2_FC_REST_LEN = 0, BR2_FC_INV_REST_LEN = 1, BR2_FC_VOL_OVER_PI = 2,
    BR2_FC_SHEAR = 3, BR2_FC_J = 6, BR2_FC_JINV = 9, BR2_FC_LN_CR = 12, BR2_FC_CR = 15,
    BR2_FC_REST_VLEN = 18, BR2_FC_INV_REST_VLEN = 19, BR2_FC_BEND = 20,
    BR2_FC_CT = 23,          /* translational damping factor of the slot's rod (1 if none) */
    BR2_FC_GRAVITY = 24,     /* 3: summed gravity acceleration on the slot's rod */
    BR2_FC_JK = 27, BR2_FC_JNU = 28, BR2_FC_JKREP = 29,  /* parameters of the slot's own joint */
    BR2_FC_PIN = 30,         /* 12: pinned position (3) and director (9) for BR2_FAST_PIN_* slots */
    BR2_FC_COUNT = 42
};

typedef struct br2_program {
    int32_t abi_version;          /* BR2_ABI_VERSION */
    int32_t n_slots, n_rods, n_forcing, n_constraints, n_joints, n_actions;
    int64_t words_per_instance;   /* doubles of state + derived per instance */
    const int32_t *slot_rod;      /* [n_slots] rod of each slot */
    const int32_t *rod_i;
    const double *rod_d;
    const double *slot_c;
    const int32_t *forcing_i;
    const double *forcing_d;
    const int32_t *cons_i;
    const double *cons_d;
    const int32_t *joint_i;
    const double *joint_d;
    /* CSR gather lists, per slot, in joint registration order.  item = (joint << 1) | bit.
     * node list: bit 0 -> rod one (+F/2), bit 1 -> rod two (-F/2).  elem list: bit = which side's torque. */
    const int32_t *node_ptr;      /* [n_slots + 1] */
    const int32_t *node_items;
    const int32_t *elem_ptr;      /* [n_slots + 1] */
    const int32_t *elem_items;
    const int32_t *overwrite_src; /* [n_slots] element slot whose director/omega replace this slot's during
                                     synchronize (tip-to-base joint), or -1 */
    /* fast path */
    int32_t fast_ok;              /* 1: the fast tables below are valid */
    int32_t fast_uniform;         /* 1: every element slot has identical fast_c constants (up to 1e-13
                                     relative; kernel then reads them from the constant bank) */
    int32_t n_act;                /* actuation rows */
    int32_t fast_point_forces;    /* 1: the forcing table holds BR2_FORCE_POINT rows (served by the tile kernel as extra
                                     forces, or by the generic kernel; the one-slot-per-thread fast kernel refuses) */
    double act_ramp_up_time;
    const int32_t *fast_i;
    const double *fast_c;
    const int32_t *act_i;         /* [n_act] action index */
    const double *act_d;          /* [n_act][3] material-frame torque per unit pressure at full ramp */
} br2_program;

/* Instance state layout (doubles), per rod at BR2_ROD_WOFF, component-major like numpy [3, n]:
 *   position 3(n+1) | velocity 3(n+1) | director 9n | omega 3n            <- read at launch, written back
 *   acceleration 3(n+1) | alpha 3n | internal_forces 3(n+1) | internal_torques 3n |
 *   external_forces 3(n+1) | external_torques 3n | lengths n | dilatation n | radius n   <- written only
 * (exactly the 12 arrays FreeCallback copies, /root/reference/br2/free_simulator.py:56-72, plus alpha). */

typedef struct br2_handle_s *br2_handle;

const char *br2_last_error(void);
int br2_abi_version(void);

/* Lower-level life cycle ------------------------------------------------------------------- */
int br2_create(const br2_program *program, int device, br2_handle *out);
int br2_destroy(br2_handle h);
int64_t br2_words_per_instance(br2_handle h);

/* Borrow a device buffer of batch * words_per_instance doubles (instance-major). */
int br2_bind_state(br2_handle h, double *dev_state, int64_t batch);
/* Borrow per-instance actuation pressures, device, [n_actions][batch] doubles
 * (replaces FreeAssembly.set_actuation's shared python lists, free_simulator.py:276-289). */
int br2_set_actuation(br2_handle h, const double *dev_pressures);

/* Material parameters that differ BETWEEN the instances of the batch (sweeps over Young's / shear modulus, damping constant,
 * fibre angles: the rod-spec keys of free_simulator.py:108-120, 291-340), tile kernel only.  Borrowed device buffers:
 * dev_uni [batch][BR2_FC_COUNT] = per instance what column 0 of fast_c holds for the shared case; dev_act_d
 * [batch][n_act][3] = per instance act_d; dev_joint_scale [batch] = factor on the tip-to-base joints' stiffness (it is
 * proportional to Young's modulus, :220-228).  dev_uni == NULL returns to the shared constants.  Geometry, masses and the
 * joint direction vectors stay common to the batch.  The exact-path replay of one-CTA assemblies reads the shared tables,
 * so a per-instance run that leaves the optimistic kernel's range is reported by br2_unresolved_count instead of redone.
 * Defined for assemblies with ONE set of element constants; an assembly whose rods differ in their constants (segments of
 * different element count) already runs on class tables of its own and refuses per-instance ones (BR2_ESTATE). */
int br2_set_instance_params(br2_handle h, const double *dev_uni, const double *dev_act_d, const double *dev_joint_scale);

/* Advance every bound instance n_steps PositionVerlet steps starting at time t0 (replaces n_steps
 * calls of do_step, environment.py:279-285).  Asynchronous on `stream` (a cudaStream_t, e.g.
 * torch.cuda.current_stream().cuda_stream).  *t_end receives the fp64-accumulated end time. */
int br2_step(br2_handle h, int64_t n_steps, double t0, double dt, void *stream, double *t_end);

/* Same through HOST buffers: H2D of state and pressures, step, D2H of the state; synchronous. */
int br2_step_host(br2_handle h, double *host_state, const double *host_pressures, int64_t batch,
                  int64_t n_steps, double t0, double dt, double *t_end);

/* Number of do_step calls `while time < t0 + duration` makes, and the time it ends at
 * (environment.py:278: the count comes from fp64 accumulation, two +dt/2 per step). */
int br2_count_steps(double t0, double duration, double dt, int64_t *n_steps, double *t_end);

/* Where a run of n_steps steps from t0 must stop for diagnostics callbacks: the steps (1-based ordinals within the run)
 * after which `round(time / dt) % every[i] == 0` holds for some collector i, on the fp64-accumulated time -- the test
 * FreeCallback.make_callback performs every step (free_simulator.py:49-51; every[i] <= 1 = every step).  Writes up to
 * `capacity` ordinals / times, and the total number of stops to *count (call again with more room if count > capacity). */
int br2_capture_plan(double t0, int64_t n_steps, double dt, const int64_t *every, int32_t n_every, int64_t *ordinals,
                     double *times, int64_t capacity, int64_t *count);

/* Which kernel br2_create would pick for this program and how it would lay the assembly out ("tile: 3 CTA(s) per instance
 * of 303, 303, 303 threads ..."), or why the faster kernels refuse it.  Host only: needs no GPU.  BR2_ELIMIT if the
 * assembly fits no kernel. */
int br2_describe_plan(const br2_program *program, char *buf, int64_t buf_len);

/* Kernel facts for measurement: registers per thread, dynamic shared memory per block, threads per
 * block, resident blocks per SM as reported by the occupancy API, SM count of the device, and which
 * kernel br2_step launches (0 generic, 1 fast per-slot constants, 2 fast uniform constants). */
int br2_kernel_info(br2_handle h, int32_t *regs, int32_t *smem_bytes, int32_t *threads,
                    int32_t *blocks_per_sm, int32_t *sm_count, int32_t *variant);
/* Force the kernel variant (0 generic, 1 fast, 2 fast uniform, -1 automatic); for cross-checks. */
int br2_select_kernel(br2_handle h, int32_t variant);
/* Instance-launches the exact-path (generic) kernel had to redo because the fast kernel left the validity
 * range of its series on them (normally 0).  Synchronises the device. */
int br2_replay_count(br2_handle h, int64_t *count);
/* Instances still flagged after the replay pass, i.e. that left the optimistic kernel's range and could NOT be redone.
 * Expected to be 0: assemblies of up to 1024 node slots are replayed by the generic kernel, cluster-sized ones by the
 * exact-arithmetic instantiation of the cluster kernel.  A non-zero count means an instance has stopped advancing
 * (Environment.run(check_nan=True) raises on it).  Synchronises the device. */
int br2_unresolved_count(br2_handle h, int64_t *count);
/* Of the replays, how many left the range of the rotation-angle series (|omega| dt/2 too large), the curvature series
 * (neighbouring elements rotated too far apart) and the damper's exponential (axial strain too large); only
 * the tile kernel records reasons.  Synchronises the device. */
int br2_replay_reasons(br2_handle h, int64_t *rotation, int64_t *curvature, int64_t *strain);
/* Terminal metrics of a run, per instance, in ONE pass over the batch state on `stream` (replaces the host-side checks of
 * environment.py:289-349): dev_out [batch][BR2_METRICS_N] doubles =
 *   [0] NaN mask (bit 0 position, 1 velocity, 2 director, 3 omega; :343-348)   [1] max_nodes |x| (the reference's "velocity"
 *   steady-state test measures positions, :292-298; NaN if any is)             [2] max_nodes |v|
 *   [3..8] nanmax |previous - current| of position, velocity, acceleration, director (Frobenius norm per element), omega,
 *   alpha (:300-338) when dev_previous -- a copy of the state block taken before the run -- is given, else -inf.
 * Asynchronous; read dev_out after synchronising the stream. */
#define BR2_METRICS_N 12
int br2_metrics(br2_handle h, const double *dev_previous, double *dev_out, void *stream);

/* Evaluate the fast kernel's fp64 helpers on device arrays of n doubles (tests only):
 * which = 0 rcp, 1 rsqrt, 2 sinc(theta) [in: theta^2 <= 0.01], 3 cosc(theta) [in: theta^2 <= 0.01],
 *         4 log-map scale [in: cos theta >= 0.875], 5 exp(d) [in: |d| <= 0.02];
 *         tile kernel's shorter series: 6 sinc, 7 cosc [in: theta^2 <= 2e-3], 8 log-map scale [in: cos theta >= 0.97],
 *         9 exp(d) [in: |d| <= 0.09], 10 sqrt to 2^-40, 11 rotation angle of one half step of 1e-5 s about d3
 *         for omega = in (checks the 1e-14 guard of the reference's axis normalisation). */
int br2_selftest_math(int32_t which, const double *dev_in, double *dev_out, int64_t n);

/* fp64 roofline denominator: runs a register-resident DFMA chain kernel on `device` and returns the
 * best-of-`repeats` rate in TFLOP/s (FMA = 2 flop). */
int br2_fp64_peak_tflops(int device, int32_t repeats, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* BR2CUDA_H */
