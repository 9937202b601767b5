// br2_tile.cuh -- the shared-memory-lean performance kernel for ring-structured BR2 assemblies.
//
// The first fast kernel (br2_fast.cuh, one slot per thread) turned out to be bound by the SM's shared-memory
// crossbar (128 B/clk): 152 eight-byte accesses per slot and step = 77 % of the crossbar while the fp64 pipe
// sat at 57 % (profiles/r01_ncu_fast_kernel_c.txt).  This kernel gives every thread C CONSECUTIVE slots of one
// rod (node k, element k, Voronoi cell k for k = k0 .. k0+C-1):
//   * neighbour differences inside a thread's run come from registers; only the run's first node / element
//     is published for the previous thread, only the run's last shear force / bending couple for the next;
//   * a surface joint is evaluated once, by its rod-two element; the rod-one element publishes what the
//     joint needs from it (sum of its two nodes' positions and velocities, its world-frame lever arm and
//     arm length: 10 doubles) instead of its director, nodes and radius (23 doubles), and picks up the
//     force afterwards to form its own torque, so nobody reads a partner's director at all;
//   * forces go into the node velocities as soon as they exist (the update is linear in the force), so no
//     force accumulator lives across a barrier.
// Result: about half the shared-memory traffic per slot and step, and the C slots of a thread are independent
// instruction streams for the fp64 pipe (8-cycle dependent-issue latency, 2 cycles per warp instruction:
// profiles/probes/dfma_latency_probe.cu).
//
// Synchronisation (round 2).  A step has three exchange points (first node / element -> element quantities -> joint
// forces and couples -> dynamic step).  With one large CTA per SM -- the multi-segment and the cluster kernels -- a
// CTA-wide barrier at each of them left the SM idle while the last warp arrived (ncu, cluster kernel: 1.38 barrier +
// 0.40 membar stall cycles per issued instruction).  The barriers are now SPLIT-PHASE mbarriers (kSplit): a warp
// arrives as soon as it has published, goes on with work that needs nobody else's data (slot 0's element quantities
// after the first, all Voronoi couples after the second, the tip-to-base joints and the damper exponentials after the
// third) and waits only where it reads a neighbour's rows.  Tip-to-base joints between segments no longer use the
// hardware cluster barrier at all: the few elements involved PUSH director, omega, radius and node sums into the
// consuming CTAs' shared memory with asynchronous distributed-shared-memory stores that count their bytes off the
// consumer's mbarrier (st.async ... complete_tx: no cluster-scope fence anywhere); both CTAs of an interface evaluate its
// joints themselves (bit-identical arithmetic), so the forces never travel back.  The kernels with several CTAs per SM
// (single-segment: two warps per CTA, six CTAs per SM; double / triple on one CTA) keep __syncthreads: measured.
//
// Semantics are those of br2_fast.cuh / br2_generic.cuh; the optimistic-series
// protocol is the same: a CTA that leaves a series' validity range flags its instance and does not write back;
// the exact-path pass launched right behind on the stream redoes the instance from the chunk that failed (the generic
// kernel, or this kernel's kExact instantiation for assemblies too large for it).
//
// Work distribution: persistent CTAs (or clusters) pull (chunk, instance) items off one queue -- see the kernel.
// Template flags: kExtras = tip-to-base joints between segments (director overwrites included); kCluster = one CTA per
// group of ring-connected rods, a thread-block cluster per instance (assemblies beyond one SM's registers);
// kExact = IEEE / libm arithmetic without validity ranges; kSplit = split-phase mbarriers instead of __syncthreads.
//
// Applicability (decided on the host, br2_cuda.cu: make_tile_plan): fast-path assembly with uniform element
// constants, every surface joint connects equal element indices of two rods, all joints of a rod pair share one
// descriptor whose direction vectors have no component along the rod axis d3 (rods side by side), every other joint
// is a tip-to-base joint, and each group of ring-connected rods fits one CTA (320 threads = 640 slots).
#pragma once

#include <cooperative_groups.h>

#include "br2_math.cuh"
#include "br2_tables.cuh"

namespace br2 {

// CTA-wide or (kCluster) cluster-wide barrier, and the address of a peer CTA's copy of a shared-memory location
template <bool kCluster>
__device__ __forceinline__ void tile_sync() {
    if (kCluster)
        cooperative_groups::this_cluster().sync();
    else
        __syncthreads();
}
template <bool kCluster, typename T>
__device__ __forceinline__ T *tile_peer(T *local, int rank) {
    if (kCluster) return cooperative_groups::this_cluster().map_shared_rank(local, rank);
    return local;
}

// ---- split-phase barriers (mbarrier objects in shared memory) ------------------------------------------------------
// One elected lane per warp arrives (release at CTA scope; __syncwarp orders the other lanes' stores before it), every
// lane waits on the phase parity (of the step counter: every barrier completes once per step).  TRYWAIT suspends the warp in hardware instead of spinning.
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_warp(uint32_t bar) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// (the suspend-time hint keeps a waiting warp asleep until the phase completes: without it TRYWAIT returned after a few
// hundred cycles and the loop around it cost 120 issued instructions per warp and step on the cluster kernel)
#ifndef BR2_MBAR_SUSPEND_HINT
#define BR2_MBAR_SUSPEND_HINT 0x989680u
#endif
constexpr uint32_t kMbarSuspendHint = BR2_MBAR_SUSPEND_HINT;
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (kMbarSuspendHint != 0u)
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "WAIT_LOOP:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
            "@p bra WAIT_DONE;\n"
            "bra WAIT_LOOP;\n"
            "WAIT_DONE:\n"
            "}" ::"r"(bar), "r"(parity), "r"(kMbarSuspendHint) : "memory");
    else
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "WAIT_LOOP:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@p bra WAIT_DONE;\n"
            "bra WAIT_LOOP;\n"
            "WAIT_DONE:\n"
            "}" ::"r"(bar), "r"(parity) : "memory");
}
// ---- pushes into a peer CTA's shared memory (cluster kernels) -------------------------------------------------------
// st.async writes 8 bytes into the peer's shared memory and, when the write has landed, counts them off the peer's
// mbarrier (complete_tx).  The consumer arms the barrier with the byte count it expects (one arrive.expect_tx per step)
// and waits on it at CTA scope: no release / acquire fence at cluster scope anywhere -- those compile to MEMBAR.ALL.GPU
// on the producer and CCTL.IVALL (an L1 invalidation) on the consumer, which the first version of this path paid per step.
__device__ __forceinline__ uint32_t cluster_addr(uint32_t local, int rank) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(rank));
    return remote;
}
__device__ __forceinline__ void st_async_f64(uint32_t remote, double v, uint32_t remote_bar) {
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(remote),
                 "l"(__double_as_longlong(v)), "r"(remote_bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// the step's exchange points: arrive after publishing, wait before reading; kSplit = false degenerates to one
// __syncthreads at the wait
#ifndef BR2_TILE_SHUFFLE
#define BR2_TILE_SHUFFLE 0  // A/B: the next run's first node / element through warp shuffles where it is the next lane (measured: DESIGN.md)
#endif
#ifndef BR2_TILE_GATHER_MERGED
#define BR2_TILE_GATHER_MERGED 1  // tip-to-base forces: one warp-level branch, both slots' rows loaded together (A/B: 0)
#endif
#ifndef BR2_TILE_SCALAR_MASKS
#define BR2_TILE_SCALAR_MASKS 1  // flag masks applied to scalar coefficients instead of to vector results, max(x, 0) as one select (A/B: 0)
#endif
#ifndef BR2_TILE_LEAN_FMA
#define BR2_TILE_LEAN_FMA -1  // additions folded into neighbouring FMA chains: bit 0 rod-one joint torque, bit 1 tint + text; -1 = both on the
                              // kernels with tip-to-base joints (+0.5 ... +1 %), bit 0 on the single-segment kernel (bit 1: -1.5 % there)
#endif
#ifndef BR2_TILE_LEAN_SQRT
#define BR2_TILE_LEAN_SQRT 1  // radius as sqrt (cubic step from the rsqrt seed) instead of x * rsqrt(x); strain rate over |d|^2 (A/B: 0)
#endif
#ifndef BR2_TILE_AOS_PUB
#define BR2_TILE_AOS_PUB -1  // what an element publishes for its joints (node sums 3 + 3, lever arm 3, arm length) as ONE 80-byte row per
                             // slot, stored and loaded 16 bytes at a time, instead of one component-major row per double: 26 LDS / STS
                             // fewer per warp-step; -1 = cluster kernels only (measured: +1.7 % there, -1.2 % single, -1.8 % double:
                             // the wide accesses constrain ptxas' register allocation -- predicted by the issue model for single)
#endif
#ifndef BR2_TILE_HOIST_DAMPER
#define BR2_TILE_HOIST_DAMPER 0  // damper exponentials before the third wait (costs registers: measured)
#endif
// x > 0 ? x : 0 as ONE compare and one select (fmax(x, 0.0) compiles to DSETP.MAX plus a five-instruction NaN fix-up that
// returns the same values: the non-NaN operand, and +0 for -0)
__device__ __forceinline__ double sel_gt0(double x) {
    double r;
    asm("{\n.reg .pred p;\nsetp.gt.f64 p, %1, 0d0000000000000000;\nselp.f64 %0, %1, 0d0000000000000000, p;\n}" : "=d"(r) : "d"(x));
    return r;
}
struct TileBars {
    uint32_t b1, b2, b3, b4, xb;  // shared-memory addresses: three phase barriers, joint forces ready, pushed data (xb, xb + 8 by step parity)
};
template <bool kSplit>
__device__ __forceinline__ void tile_arrive(uint32_t bar) {
    if (kSplit) mbar_arrive_warp(bar);
}
// Where the per-thread constants live was measured (single_br2 x 4096, el-steps/s): joint direction vectors from the
// constant bank indexed by rod 3.05e10, from a per-rod shared-memory table 2.96e10; actuation torques in shared-memory
// rows 3.05e10, in registers 2.92e10 (every register kept across the step costs instruction-level parallelism).

// Arithmetic of the step: the optimistic series / hardware seeds (kExact = false), or IEEE division, sqrt and libm
// with no validity range (kExact = true).  The exact instantiation is the replay path of assemblies that are too large
// for the one-CTA generic kernel: it redoes, from the chunk in which it was flagged, an instance that left a series' range.
template <bool kExact>
struct TileMath {
    static __device__ __forceinline__ double rsqrt(double x) { return kExact ? 1.0 / sqrt(x) : fast_rsqrt(x); }
    static __device__ __forceinline__ double rcp(double x) { return kExact ? 1.0 / x : fast_rcp(x); }
    static __device__ __forceinline__ double rsqrt_lo(double x) { return kExact ? 1.0 / sqrt(x) : seed_rsqrt(x); }
    static __device__ __forceinline__ double rcp_lo(double x) { return kExact ? 1.0 / x : seed_rcp(x); }
    static __device__ __forceinline__ double sqrt_lo(double x) { return kExact ? sqrt(x) : sqrt_mid(x); }
    static __device__ __forceinline__ double sqrt(double x) { return kExact ? ::sqrt(x) : fast_sqrt(x); }
    static __device__ __forceinline__ double exp(double d) { return kExact ? ::exp(d) : exp_tiny(d); }
    // theta / sin(theta + 1e-14) from y = 1 - cos(theta)
    static __device__ __forceinline__ double logmap_ratio(double y) {
        if (!kExact) return logmap_ratio_tiny(y);
        const double theta = acos(1.0 - y);
        return theta / sin(theta + BR2_SIN_GUARD);
    }
    // Q <- R Q for one (hmult == hdt) or two fused (hmult == 2 hdt) kinematic half steps; true = left the series' range
    static __device__ __forceinline__ bool rotate(double (&Q)[9], const double (&w)[3], double hdt, double hmult) {
        if (!kExact) return rotate_directors_tiny(Q, w, hdt, hmult);
        const double u0 = hdt * w[0], u1 = hdt * w[1], u2 = hdt * w[2];
        const double theta = sqrt(u0 * u0 + u1 * u1 + u2 * u2);
        const double inv = 1.0 / (theta + BR2_ROTATION_NORM_GUARD);
        const double a0 = u0 * inv, a1 = u1 * inv, a2 = u2 * inv;   // the reference's guarded axis (Appendix A.4)
        double sn, cs;
        sincos((hmult > 1.5 * hdt) ? 2.0 * theta : theta, &sn, &cs);  // same axis both half steps: the angles add
        const double q = 1.0 - cs;
        const double R00 = 1.0 - q * (a1 * a1 + a2 * a2), R11 = 1.0 - q * (a0 * a0 + a2 * a2), R22 = 1.0 - q * (a0 * a0 + a1 * a1);
        const double q01 = q * a0 * a1, q02 = q * a0 * a2, q12 = q * a1 * a2;
        const double R01 = sn * a2 + q01, R10 = -sn * a2 + q01, R02 = -sn * a1 + q02, R20 = sn * a1 + q02;
        const double R12 = sn * a0 + q12, R21 = -sn * a0 + q12;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double q0 = Q[c], q1 = Q[3 + c], q2 = Q[6 + c];
            Q[c] = R00 * q0 + R01 * q1 + R02 * q2;
            Q[3 + c] = R10 * q0 + R11 * q1 + R12 * q2;
            Q[6 + c] = R20 * q0 + R21 * q1 + R22 * q2;
        }
        return false;
    }
};

template <int C, int NT>
struct TileSmem {
    // rows of every exchange array: thread t owns row t + 1; row 0 is all zeros and never written.  A run without a
    // predecessor reads row (own - 1): row 0, or the previous rod's last run, whose "last element" entries are zero because
    // its last slot is the node-only slot.  A run without a successor reads its own row (the result is masked).  Both
    // keep a half-warp's addresses contiguous, i.e. free of bank conflicts.
    static constexpr int R = NT + 1;
    static constexpr int XF = 0;                  // [3][R]    position of the run's first node
    static constexpr int VF = XF + 3 * R;         // [3][R]    velocity of the run's first node
    static constexpr int QF = VF + 3 * R;         // [9][R]    director of the run's first element
    static constexpr int LF = QF + 9 * R;         // [R]       length of the run's first element
    static constexpr int SKL = LF + R;            // [3][R]    Q^T n_L / e of the run's last element
    static constexpr int ML = SKL + 3 * R;        // [3][R]    C/2 - A of the run's last Voronoi cell (A = tau_L / E^3, C = (kappa x tau_L) D / E^3)
    static constexpr int PC = ML + 3 * R;         // [C][3][R] x_k + x_{k+1} of each element (twice its centre)
    static constexpr int PR = PC + 3 * C * R;     // [C][3][R] world-frame lever arm of the element as rod one
    static constexpr int PV = PR + 3 * C * R;     // [C][3][R] v_k + v_{k+1}
    static constexpr int PA = PV + 3 * C * R;     // [C][R]    arm length as rod one (radius + its share of the rest gap)
    static constexpr int HF = PA + C * R;         // [C][3][R] half force of the joint the element owns (it is rod two)
    static constexpr int FS = HF + 3 * C * R;     // [C][3][R] spring part k d of that joint (the torques use only this)
    static constexpr int ZERO_END = FS + 3 * C * R;  // everything above is zeroed at launch
    static constexpr int TA = ZERO_END;           // [C][3][R] actuation torque at full ramp (material frame)
    static constexpr int ROWS_END = TA + 3 * C * R;
    // after the rows (host: TileSmem<C, 0>::ROWS_END * (threads + 1) words), the extra-joint area of a kExtras kernel;
    // every CTA has its own complete copy, filled by the elements' owners (pushes through distributed shared memory):
    //   EXD [2][n_xe][20]  by step parity, per published element: director 9 | omega 3 | radius | x_k + x_{k+1} 3 | v_k + v_{k+1} 3
    //   NF  [n_nr][items][4] node rows: the signed half forces a node adds, item by item (scattered there by the threads
    //                      that evaluate the joints, in the order the reference applies them)
    //   XJD [n_xj][8]      k, nu, l1[3], l2[3]   (point force: F[3], ramp_up_time)
    //   XJI [n_xj][8] int  published element of: rod one's element, rod two's element, director one, director two; type
    //   XJT [n_xj][4] int  where in NF the joint's half force goes: rod one's two nodes (+), rod two's (-); -1 = another CTA's
    //   XPM [n_xe]    int  bit r set: the CTA of cluster rank r consumes the element's data
    static constexpr int EX = ROWS_END;
    static constexpr int EXW = 20;                // doubles per published element (row stride)
    static constexpr int EXN = 19;                // ... of which in use
};

constexpr int TILE_F_VALID = 1, TILE_F_ELEM = 2, TILE_F_VOR = 4, TILE_F_ZERO_V = 8, TILE_F_ZERO_W = 16, TILE_F_OWN = 32,
              TILE_F_REPORT0 = 64, TILE_F_SOFTFIX = 128;

template <int C>
struct TileThread {
    double x[C][3], v[C][3], Q[C][9], w[C][3];
    double dtmc[C];  // dt c_t / node mass
    int fl[C];
    int t, tn, t1, to;  // rows: own, next run (or own), partner as rod one of my joints (or 0), owner of the joints in which I am rod one (or 0)
    int rod;
    int uoff;  // kPerInst kernels: where my rod's class of constants starts in the CTA's shared-memory copy
    // extra joints (kExtras kernels): per slot published xe + 1 | overwrite source xe + 1 << 8 | node items << 16 |
    // node row << 20; the extra joint component this thread evaluates (joint << 2 | component) or -1
    int xinfo[C], xjoint;
    int bad;  // bit 0: rotation angle, bit 1: curvature angle, bit 2: strain left its series' range (or NaN)
};

// Where a thread's run sits in the assembly and in the instance's state block.  Recomputed (behind an opaque
// copy of the thread index, so the compiler cannot keep it alive) wherever it is needed outside the step loop:
// these five values would otherwise occupy registers for the whole launch.
struct TileWhere {
    int rod, n, k0, s0;
    double *inst;
};

template <int C>
__device__ __forceinline__ TileWhere tile_where(const StepParams &p, int t, long long b, int rank) {
    asm volatile("" : "+r"(t));
    const DevTables &tb = p.t;
    const TilePlan &tp = p.tile;
    TileWhere wh;
    int rod = -1, first = 0;
    for (int r = tb.n_rods - 1; r >= 0; --r) {
        if (tp.rod_cta[r] != rank) continue;
        first = r;
        if (rod < 0 && t >= tp.rod_t0[r]) rod = r;
    }
    if (rod < 0) rod = first;
    const int *ri = tb.rod_i + rod * BR2_ROD_NI;
    wh.rod = rod;
    wh.n = ri[BR2_ROD_N];
    wh.k0 = (t < tp.cta_threads[rank]) ? C * (t - tp.rod_t0[rod]) : wh.n + 1;
    wh.s0 = ri[BR2_ROD_SLOT0] + wh.k0;  // table slot of my first node (when valid)
    wh.inst = p.state + b * tb.words + ri[BR2_ROD_WOFF];
    return wh;
}

// A/B experiment (BR2_TILE_SHUFFLE): what the next run published, from the next lane's registers when that run lives in
// the next lane of this warp, from its shared-memory row otherwise (last lane of a warp, last run of a rod)
__device__ __forceinline__ double tile_next(double mine, const double *row, bool via_row) {
    double up = __shfl_down_sync(0xffffffffu, mine, 1);
    if (via_row) up = *row;
    return up;
}

// exp(strain ln c_r) of the rotational damper (Appendix A.5): one exponential per element for a round section
template <bool kExact, bool kPerInst>
__device__ __forceinline__ void tile_damper_exponentials(const StepParams &p, const double *smu, double len, int flags,
                                                         double (&out)[3], int &bad) {
    using M = TileMath<kExact>;
#define UNI(idx) (kPerInst ? smu[(idx)] : p.uni[(idx)])
    const double dil = len * UNI(BR2_FC_INV_REST_LEN);
    const double strain = (flags & TILE_F_ELEM) ? dil - 1.0 : 0.0;
    if (p.tile.round_section) {
        const double e2 = strain * UNI(BR2_FC_LN_CR + 2);
        if (!kExact) bad |= (e2 * e2 <= BR2_E2_MAX) ? 0 : 4;
        out[0] = M::exp(e2);
    } else {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const double e = strain * UNI(BR2_FC_LN_CR + i);
            if (!kExact) bad |= (e * e <= BR2_E2_MAX) ? 0 : 4;
            out[i] = M::exp(e);
        }
    }
#undef UNI
}

// One PositionVerlet step of a thread's run.  kEnd: last step of a chunk (ends with a single kinematic half step
// instead of the fused pair, so the state is a whole-step state); kLast: last step of the launch (also reports
// the callback quantities).  gstep: steps this CTA has executed since the kernel started (phase parity of the
// barriers; its low bit selects the buffer of the pushed tip-to-base data).
template <int C, int NT, bool kEnd, bool kLast, bool kExtras, bool kCluster, bool kExact, bool kSplit, bool kPerInst>
__device__ __forceinline__ void tile_step(const StepParams &p, double *sm, const double *smu, TileThread<C> &th, double time_mid,
                                          long long b, int rank, int gstep, const TileBars &bars) {
    using L = TileSmem<C, NT>;
    constexpr int R = L::R;
    static_assert(C >= 2, "slot 0's element must lie inside the thread's run");
    static_assert(kExtras || !kCluster, "the cluster kernels carry the extra-joint code");
    using M = TileMath<kExact>;
    constexpr int kLeanFma = BR2_TILE_LEAN_FMA < 0 ? (kExtras ? 3 : 1) : BR2_TILE_LEAN_FMA;
    constexpr bool kAosPub = BR2_TILE_AOS_PUB < 0 ? kCluster : (BR2_TILE_AOS_PUB != 0);
    const DevTables &tb = p.t;
    const int t = th.t;
    const double dt = p.dt, hdt = p.kd[11];
    // loop-invariant products of the uniform constants and dt come from the constant bank (StepParams::kd, filled by br2_step
    // with the very expressions quoted here) instead of being hoisted into registers that then spill
    // kPerInst (material parameters that differ between the instances of a batch): both tables come from this CTA's copy
    // in shared memory (smu: uni[BR2_FC_COUNT] then kd[BR2_KD_COUNT]), read as broadcasts, instead of from the constant bank
#define KD(idx, expr) (kPerInst ? smu[BR2_FC_COUNT + (idx)] : p.kd[(idx)])
    TileWhere wh{};
    if (kLast) wh = tile_where<C>(p, th.t - 1, b, rank);
    const int n = wh.n, np1 = wh.n + 1;
    double *const inst = wh.inst;
#define UNI(idx) (kPerInst ? smu[(idx)] : p.uni[(idx)])
#define JDV(i) p.tile.jd[th.rod][(i)]
#define TAV(j, i) sm[L::TA + (3 * (j) + (i)) * R + t]
    double (&x)[C][3] = th.x;
    double (&v)[C][3] = th.v;
    double (&Q)[C][9] = th.Q;
    double (&w)[C][3] = th.w;

    // reported quantities (kLast only)
    double r_fint[C][3], r_fext[C][3];

    // extra joints (tip-to-base joints between segments; kExtras kernels only): a handful of threads involved
    const int n_xe = kExtras ? p.tile.n_xe : 0, n_xj = kExtras ? p.tile.n_xj : 0;
    const int buf = gstep & 1;
    double *const EXD = sm + L::EX + buf * (L::EXW * n_xe);           // this step's buffer of pushed element data
    double *const NF = sm + L::EX + 2 * L::EXW * n_xe;
    const int nf_row = 4 * (kExtras ? p.tile.nf_items : 0);
    const double *const XJD = NF + nf_row * (kExtras ? p.tile.n_nr : 0);
    const int *const XJI = (const int *)(XJD + 8 * n_xj), *const XJT = XJI + 8 * n_xj, *const XPM = XJT + 4 * n_xj;
    const uint32_t xbar = bars.xb + 8u * (uint32_t)buf, xparity = ((uint32_t)gstep >> 1) & 1u;

    // =============================== P1: publish the run's first node / element ==============
    // (split-phase kernels: the first element's length with them -- the previous run's last Voronoi cell needs it, and this
    // way every Voronoi cell is computable before the second exchange point is waited for)
    double d0[3], len0, rs0;
    if (kSplit) {
#pragma unroll
        for (int i = 0; i < 3; ++i) d0[i] = x[1][i] - x[0][i];
        double d2 = fma(d0[0], d0[0], fma(d0[1], d0[1], d0[2] * d0[2]));
        d2 = (th.fl[0] & TILE_F_ELEM) ? d2 : 1.0;  // keeps the node-only slot finite
        rs0 = M::rsqrt(d2);
        len0 = fma(d2, rs0, BR2_EPS_LEN);
        sm[L::LF + t] = len0;
    }
    const bool next_via_row = BR2_TILE_SHUFFLE && ((threadIdx.x & 31) == 31 || th.tn != t + 1);
    double x0_start[3], v0_start[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        sm[L::XF + i * R + t] = x[0][i];
        sm[L::VF + i * R + t] = v[0][i];
        x0_start[i] = x[0][i];
        v0_start[i] = v[0][i];
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) sm[L::QF + i * R + t] = Q[0][i];
    if (kCluster && threadIdx.x == 0 && p.tile.xb_count[rank] > 0)  // arm this step's buffer: bytes my peers will push into it
        mbar_expect_tx(xbar, (uint32_t)(L::EXN * 8 * p.tile.xb_count[rank]));
    tile_arrive<kSplit>(bars.b1);
    if (!kSplit) __syncthreads();  // B1

    // =============================== P2: element quantities ==================================
    // slot 0 .. C-2 need nobody else's data; the run's last slot reads the next run's first node
    double len[C], ra2[C];
    double tint[C][3];   // internal couple on element k (complete except for the previous run's last cell)
    {
        const double dtg[3] = {KD(0, dt * UNI(BR2_FC_CT) * UNI(BR2_FC_GRAVITY)), KD(1, dt * UNI(BR2_FC_CT) * UNI(BR2_FC_GRAVITY + 1)),
                               KD(2, dt * UNI(BR2_FC_CT) * UNI(BR2_FC_GRAVITY + 2))};
        double skp[3];  // the previous element's Q^T n_L / e
        double radx[C];
#pragma unroll
        for (int j = 0; j < C; ++j) {
            if (kSplit && j == C - 1) mbar_wait(bars.b1, (uint32_t)gstep & 1u);  // B1
            const bool elem = th.fl[j] & TILE_F_ELEM;
            double d[3], dv[3], pcs[3], pvs[3];
            double2 *const pub = (double2 *)(sm + L::PC + (j * R + t) * 10);  // (AOS_PUB) my slot's published row
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const double xn = (j + 1 < C) ? x[(j + 1) % C][i]
                                  : BR2_TILE_SHUFFLE ? tile_next(x0_start[i], &sm[L::XF + i * R + th.tn], next_via_row) : sm[L::XF + i * R + th.tn];
                const double vn = (j + 1 < C) ? v[(j + 1) % C][i]
                                  : BR2_TILE_SHUFFLE ? tile_next(v0_start[i], &sm[L::VF + i * R + th.tn], next_via_row) : sm[L::VF + i * R + th.tn];
                d[i] = (kSplit && j == 0) ? d0[i] : xn - x[j][i];
                dv[i] = vn - v[j][i];
                pcs[i] = xn + x[j][i];
                pvs[i] = vn + v[j][i];
                if (!kAosPub) {
                    sm[L::PC + (3 * j + i) * R + t] = pcs[i];
                    sm[L::PV + (3 * j + i) * R + t] = pvs[i];
                }
            }
            if (kAosPub) {
                pub[0] = make_double2(pcs[0], pcs[1]);
                pub[1] = make_double2(pcs[2], pvs[0]);
                pub[2] = make_double2(pvs[1], pvs[2]);
            }
            const double rest_len = UNI(BR2_FC_REST_LEN), inv_rest_len = UNI(BR2_FC_INV_REST_LEN);
            double rs;
            if (kSplit && j == 0) {
                rs = rs0;
                len[j] = len0;
            } else {
                double d2 = fma(d[0], d[0], fma(d[1], d[1], d[2] * d[2]));
                d2 = elem ? d2 : 1.0;  // keeps the node-only slot finite
                rs = M::rsqrt(d2);
                len[j] = fma(d2, rs, BR2_EPS_LEN);
            }
            const double inv_len = rs * fma(-BR2_EPS_LEN, rs, 1.0);  // 1/(|d| + 1e-14), second order 1e-23
            const double r2 = UNI(BR2_FC_VOL_OVER_PI) * inv_len;
            const double rad = BR2_TILE_LEAN_SQRT ? M::sqrt(r2) : r2 * M::rsqrt(r2);
            const double dil = len[j] * inv_rest_len;
            const double inv_dil = rest_len * inv_len;
            // only thins the ~1e-8 m rest gap of a joint: 2^-20 relative is ample.  (1 / sqrt(e) is also radius x a constant;
            // folding it into a per-rod arm factor saves the seed and two instructions per slot and was measured SLOWER,
            // 3.12e10 against 3.17e10 el-steps/s: ptxas' schedule, not the instruction count, sets this kernel's pace.)
            const double rse = M::rsqrt_lo(dil);
            ra2[j] = fma(JDV(3), rse, rad);
            if (kLast && elem) {
                inst[15 * np1 + 21 * n + wh.k0 + j] = len[j];
                inst[15 * np1 + 22 * n + wh.k0 + j] = dil;
                inst[15 * np1 + 23 * n + wh.k0 + j] = rad;
            }
            // sigma = e Q t - d3 = Q d / l_hat - d3 (the 1/|d| of the tangent cancels against the dilatation);
            // nL holds n_L / e, which is what both the nodal force Q^T n_L / e and the couple (Q t x n_L) l_hat need
            double qd[3], nL[3];
#if BR2_TILE_SCALAR_MASKS
            const double inv_dil_s = elem ? inv_dil : 0.0;
#endif
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                qd[i] = fma(Q[j][3 * i], d[0], fma(Q[j][3 * i + 1], d[1], Q[j][3 * i + 2] * d[2]));
                const double sigma = fma(inv_rest_len, qd[i], (i == 2) ? -1.0 : 0.0);
#if BR2_TILE_SCALAR_MASKS
                // 0 in the node-only slot (sigma is finite there: d2 = 1); the two shear components, whose sigma is a plain
                // product, take the rest length from the constant (S_i / l_hat) instead of from sigma: one FMA less each
                if (BR2_TILE_LEAN_SQRT && i < 2)
                    nL[i] = (KD(12 + i, UNI(BR2_FC_SHEAR + i) * inv_rest_len) * inv_dil_s) * qd[i];
                else
                    nL[i] = (UNI(BR2_FC_SHEAR + i) * inv_dil_s) * sigma;
#else
                nL[i] = (UNI(BR2_FC_SHEAR + i) * inv_dil) * sigma;
#endif
            }
            // damping and gravity first (the update is linear), then the shear forces of elements k and k-1
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const double raw = fma(Q[j][i], nL[0], fma(Q[j][3 + i], nL[1], Q[j][6 + i] * nL[2]));
                const double sk = BR2_TILE_SCALAR_MASKS ? raw : (elem ? raw : 0.0);
                const double fi = (j == 0) ? sk : sk - skp[i];
                skp[i] = sk;
                v[j][i] = fma(th.dtmc[j], fi, fma(v[j][i], UNI(BR2_FC_CT), dtg[i]));
                if (kLast) r_fint[j][i] = fi;
            }
            {
                // d(e)/dt = (x_{k+1} - x_k).(v_{k+1} - v_k) / l / l_hat  (expanded form in the reference)
#if BR2_TILE_LEAN_SQRT
                // e_dot / e = (d . dv) / |d|^2: the rest length cancels between the strain rate and the dilatation
                const double ue = fma(d[0], dv[0], fma(d[1], dv[1], d[2] * dv[2])) * (inv_len * inv_len);
#else
                const double edot = fma(d[0], dv[0], fma(d[1], dv[1], d[2] * dv[2])) * inv_len * inv_rest_len;
#endif
                double jw[3];
#pragma unroll
                for (int i = 0; i < 3; ++i) jw[i] = UNI(BR2_FC_J + i) * w[j][i] * inv_dil;
#if !BR2_TILE_LEAN_SQRT
                const double ue = edot * inv_dil;
#endif
                tint[j][0] = fma(jw[1], w[j][2], fma(-jw[2], w[j][1], fma(jw[0], ue, fma(qd[1], nL[2], -qd[2] * nL[1]))));
                tint[j][1] = fma(jw[2], w[j][0], fma(-jw[0], w[j][2], fma(jw[1], ue, fma(qd[2], nL[0], -qd[0] * nL[2]))));
                tint[j][2] = fma(jw[0], w[j][1], fma(-jw[1], w[j][0], fma(jw[2], ue, fma(qd[0], nL[1], -qd[1] * nL[0]))));
            }
            // what the joint in which this element is rod one needs from it
            const double ra1 = fma(JDV(7), rse, rad);
            const double a0 = JDV(4) * ra1, a1 = JDV(5) * ra1;  // planar arms: no d3 component (tile plan requirement)
            double prs[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) prs[i] = fma(Q[j][i], a0, Q[j][3 + i] * a1);  // Q^T dv1 (r + o)
            if (kAosPub) {
                pub[3] = make_double2(prs[0], prs[1]);
                pub[4] = make_double2(prs[2], ra1);
            } else {
                sm[L::PA + j * R + t] = ra1;
#pragma unroll
                for (int i = 0; i < 3; ++i) sm[L::PR + (3 * j + i) * R + t] = prs[i];
            }
            if (kExtras) radx[j] = rad;  // (an element of a tip-to-base joint pushes it below)
        }
        if (!kSplit) sm[L::LF + t] = len[0];
#pragma unroll
        for (int i = 0; i < 3; ++i) sm[L::SKL + i * R + t] = skp[i];
        if (kExtras) {
            // What the tip-to-base joints need from my elements (start-of-step director and omega: the overwrite in P4 comes
            // later) goes into this CTA's buffer of this step's parity with plain stores -- ordered before their readers by
            // B2 -- and, in a cluster, from there into every OTHER consuming CTA's buffer, one asynchronous store per double
            // that counts itself off the consumer's barrier.  The remote stores are issued by the WARP, lane i moving double
            // i: one thread issuing its 19 (or 38) `st.async` in a row took 1400 (2500) cycles per step -- ~70 cycles each,
            // measured with clock64 -- and the whole cluster ran at that thread's pace.  After B1 on purpose: every warp of
            // this CTA has then finished the previous step, so a peer that races ahead on my data cannot overwrite a buffer
            // still being read.
            int pubs[C];
#pragma unroll
            for (int j = 0; j < C; ++j) {
                pubs[j] = (th.xinfo[j] & 0xFF) - 1;
                if (pubs[j] >= 0) {
                    double val[L::EXN];
#pragma unroll
                    for (int i = 0; i < 9; ++i) val[i] = Q[j][i];
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        val[9 + i] = w[j][i];
                        val[13 + i] = kAosPub ? sm[L::PC + (j * R + t) * 10 + i] : sm[L::PC + (3 * j + i) * R + t];
                        val[16 + i] = kAosPub ? sm[L::PC + (j * R + t) * 10 + 3 + i] : sm[L::PV + (3 * j + i) * R + t];
                    }
                    val[12] = radx[j];
#pragma unroll
                    for (int i = 0; i < L::EXN; ++i) EXD[pubs[j] * L::EXW + i] = val[i];
                }
            }
            if (kCluster) {
                const int lane = threadIdx.x & 31;
                __syncwarp();  // the stores above are visible to the other lanes of the warp
#pragma unroll
                for (int j = 0; j < C; ++j) {
                    unsigned owners = __ballot_sync(0xffffffffu, pubs[j] >= 0);
                    while (owners) {
                        const int src = __ffs(owners) - 1;
                        owners &= owners - 1;
                        const int pub = __shfl_sync(0xffffffffu, pubs[j], src);
                        const int mask = XPM[pub] & ~(1 << rank);
                        const double *row = EXD + pub * L::EXW;
                        for (int r = 0; r < p.tile.n_ctas; ++r) {
                            if (!((mask >> r) & 1)) continue;
                            if (lane < L::EXN)
                                st_async_f64(cluster_addr(smem_addr(row + lane), r), row[lane], cluster_addr(xbar, r));
                        }
                    }
                }
            }
        }
    }
    tile_arrive<kSplit>(bars.b2);
    if (!kSplit) __syncthreads();  // B2

    // =============================== P3a: Voronoi couples (own registers + rows published before B1) ==========
    double tw[C][3];  // world-frame joint torque on my element (own joint now, rod-one joint in P4)
    // range checks of both slots merged into th.bad once per step (one-CTA kernels: +0.5 ... +1 % together with the folded
    // joint torque; the cluster kernel measured 1 % slower with it and keeps one update per check)
    constexpr bool kMergeFlags = !kCluster;
    bool ok_y = true, ok_e = true;
    {
        double ap[3], cp[3];  // previous cell's A and C
#pragma unroll
        for (int j = 0; j < C; ++j) {
            const bool vor = th.fl[j] & TILE_F_VOR;
            double Qn[9];
#pragma unroll
            for (int i = 0; i < 9; ++i)
                Qn[i] = (j + 1 < C) ? Q[(j + 1) % C][i]
                        : BR2_TILE_SHUFFLE ? tile_next(Q[0][i], &sm[L::QF + i * R + th.tn], next_via_row) : sm[L::QF + i * R + th.tn];
            const double lenn = (j + 1 < C) ? len[(j + 1) % C]
                                : BR2_TILE_SHUFFLE ? tile_next(len[0], &sm[L::LF + th.tn], next_via_row) : sm[L::LF + th.tn];
            const double *q = Q[j];
            // antisymmetric part and trace of Qn Q^T
            const double w0 = fma(Qn[6], q[3], fma(Qn[7], q[4], fma(Qn[8], q[5], fma(-Qn[3], q[6], fma(-Qn[4], q[7], -Qn[5] * q[8])))));
            const double w1 = fma(Qn[0], q[6], fma(Qn[1], q[7], fma(Qn[2], q[8], fma(-Qn[6], q[0], fma(-Qn[7], q[1], -Qn[8] * q[2])))));
            const double w2 = fma(Qn[3], q[0], fma(Qn[4], q[1], fma(Qn[5], q[2], fma(-Qn[0], q[3], fma(-Qn[1], q[4], -Qn[2] * q[5])))));
            const double tr = fma(Qn[0], q[0], fma(Qn[1], q[1], fma(Qn[2], q[2], fma(Qn[3], q[3], fma(Qn[4], q[4],
                              fma(Qn[5], q[5], fma(Qn[6], q[6], fma(Qn[7], q[7], Qn[8] * q[8]))))))));
            // y = 1 - cos(theta) with the reference's -1e-10 guard; a benign value off the rod
            const double y = vor ? fma(-0.5, tr, kMisc[7]) : kMisc2[3];
            if (!kExact) {
                if (kMergeFlags) ok_y = ok_y && (y <= BR2_Y_MAX && y > 0.0);
                else th.bad |= (y <= BR2_Y_MAX && y > 0.0) ? 0 : 2;
            }
            const double ksr = M::logmap_ratio(y) * KD(4, -0.5 * UNI(BR2_FC_INV_REST_VLEN));
            const double ks = vor ? ksr : 0.0;
            const double vdil = (lenn + len[j]) * KD(3, 0.5 * UNI(BR2_FC_INV_REST_VLEN));
            const double iv = M::rcp(vdil);
            // A = B kappa / E^3 and C = (kappa x B kappa) D / E^3 = (kappa x A) D with kappa = ks w
            const double s3 = ks * (iv * iv * iv), cs = ks * UNI(BR2_FC_REST_VLEN);
            double Ak[3], Ck[3];
            Ak[0] = UNI(BR2_FC_BEND) * (w0 * s3);
            Ak[1] = UNI(BR2_FC_BEND + 1) * (w1 * s3);
            Ak[2] = UNI(BR2_FC_BEND + 2) * (w2 * s3);
            Ck[0] = fma(w1, Ak[2], -w2 * Ak[1]) * cs;
            Ck[1] = fma(w2, Ak[0], -w0 * Ak[2]) * cs;
            Ck[2] = fma(w0, Ak[1], -w1 * Ak[0]) * cs;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                // couple on element k: (A_k - A_{k-1}) + (C_k + C_{k-1}) / 2 + shear and transport terms
                double acc = tint[j][i] + fma(0.5, Ck[i], Ak[i]);
                if (j > 0) acc += fma(0.5, cp[i], -ap[i]);
                tint[j][i] = acc;
                ap[i] = Ak[i];
                cp[i] = Ck[i];
            }
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) sm[L::ML + i * R + t] = fma(0.5, cp[i], -ap[i]);
    }
    if (kSplit) mbar_wait(bars.b2, (uint32_t)gstep & 1u);  // B2

    // =============================== P3b: own joints (Appendix A.7) ==========================
    // my element is rod two; rod one's contribution comes from its published rows
#pragma unroll
    for (int j = 0; j < C; ++j) {
        const bool own = th.fl[j] & TILE_F_OWN;
        const double b0 = JDV(0) * ra2[j], b1 = JDV(1) * ra2[j];
        double rd2[3], cd2[3], dvec[3];
        double proj2 = 0.0;
        // my partner's published row (rod one of the joint I own) and, of my own, the node sums
        double pc1[3], pv1[3], pr1[3], pc_own[3], pv_own[3], ra1;
        if (kAosPub) {
            const double2 *const his = (const double2 *)(sm + L::PC + (j * R + th.t1) * 10);
            const double2 *const mine = (const double2 *)(sm + L::PC + (j * R + t) * 10);
            const double2 h0 = his[0], h1 = his[1], h2 = his[2], h3 = his[3], h4 = his[4];
            pc1[0] = h0.x, pc1[1] = h0.y, pc1[2] = h1.x, pv1[0] = h1.y, pv1[1] = h2.x, pv1[2] = h2.y;
            pr1[0] = h3.x, pr1[1] = h3.y, pr1[2] = h4.x, ra1 = h4.y;
            const double2 m1 = mine[1], m2 = mine[2];
            pc_own[2] = m1.x, pv_own[0] = m1.y, pv_own[1] = m2.x, pv_own[2] = m2.y;
            if (j + 1 == C) {
                const double2 m0 = mine[0];
                pc_own[0] = m0.x, pc_own[1] = m0.y;
            }
        } else {
            ra1 = sm[L::PA + j * R + th.t1];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                pc1[i] = sm[L::PC + (3 * j + i) * R + th.t1];
                pr1[i] = sm[L::PR + (3 * j + i) * R + th.t1];
                pv1[i] = sm[L::PV + (3 * j + i) * R + th.t1];
                pv_own[i] = sm[L::PV + (3 * j + i) * R + t];
                if (j + 1 == C) pc_own[i] = BR2_TILE_LEAN_SQRT ? sm[L::PC + (3 * j + i) * R + t] : sm[L::XF + i * R + th.tn] + x[j][i];
            }
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            rd2[i] = fma(Q[j][i], b0, Q[j][3 + i] * b1);  // Q^T dv2 (r + o)
            // x_k + x_{k+1} of my element: the run's last slot published exactly this sum in P2 (its own row)
            const double xs = (j + 1 < C) ? x[(j + 1) % C][i] + x[j][i] : pc_own[i];
            cd2[i] = xs - pc1[i];       // 2 (c2 - c1)
            const double g = rd2[i] - pr1[i];
            // (c2 + rd2) - (c1 + rd1), rounded to 12 decimals like np.round_(x, 12)
            dvec[i] = rint(fma(0.5, cd2[i], g) * BR2_1E12) * BR2_1EM12;
            const double vr2 = pv_own[i] - pv1[i];  // 2 v_rel
            proj2 = fma(vr2, dvec[i], proj2);
        }
        const double dist2 = fma(dvec[0], dvec[0], fma(dvec[1], dvec[1], dvec[2] * dvec[2]));
        // F = k d - nu (v_rel . d_hat) d_hat = (k - nu (v_rel . d) / |d|^2) d ; no damping below |d| = 1e-12
        const double inv_d2 = (dist2 >= BR2_1EM24) ? M::rcp_lo(dist2) : 0.0;  // damping term ~1e-8 N: seed accuracy is ample
        const double jk = UNI(BR2_FC_JK);
        const double hcoef = fma(KD(5, -0.25 * UNI(BR2_FC_JNU)) * proj2, inv_d2, KD(6, 0.5 * jk));  // half of the bracket
        // Hertz-like repulsion when the surfaces interpenetrate: -k_rep pen^1.5 along the centre line
        const double cn2 = fma(cd2[0], cd2[0], fma(cd2[1], cd2[1], cd2[2] * cd2[2])) + BR2_TINY;  // 4 |cd|^2
        const double rcn = M::rsqrt(cn2);                                 // 1 / (2 |cd|)
#if BR2_TILE_SCALAR_MASKS
        const double pen_raw = sel_gt0(fma(-0.5 * cn2, rcn, ra1 + ra2[j]));  // max(r1 + r2 - |cd|, 0); NaN -> 0 like fmax
#else
        const double pen_raw = fmax(fma(-0.5 * cn2, rcn, ra1 + ra2[j]), 0.0);  // max(r1 + r2 - |cd|, 0)
#endif
        const double pen = rint(pen_raw * BR2_PEN_SCALE) * BR2_PEN_INV_SCALE;  // np.round_(penetration_strain, 12)
        const double hmag = KD(7, -0.5 * UNI(BR2_FC_JKREP)) * (pen * M::sqrt_lo(pen)) * rcn;  // times cd2 = half the repulsion
        double fs[3];
#if BR2_TILE_SCALAR_MASKS
        // a slot without a joint of its own (the rod's last node, padding) is masked through the three scalar coefficients
        // instead of through the six results: every row it read is a row some thread writes each step, so dvec and cd2 are
        // finite and 0 * finite = 0
        const double hcoef_m = own ? hcoef : 0.0, hmag_m = own ? hmag : 0.0, jk_m = own ? jk : 0.0;
#endif
#pragma unroll
        for (int i = 0; i < 3; ++i) {
#if BR2_TILE_SCALAR_MASKS
            const double hf = fma(hcoef_m, dvec[i], hmag_m * cd2[i]);
            fs[i] = jk_m * dvec[i];
#else
            const double h = fma(hcoef, dvec[i], hmag * cd2[i]);
            const double hf = own ? h : 0.0;
            fs[i] = own ? jk * dvec[i] : 0.0;
#endif
            sm[L::HF + (3 * j + i) * R + t] = hf;
            sm[L::FS + (3 * j + i) * R + t] = fs[i];
            // rod two takes -F: half on each node of the element (node k+1 of the run's last element belongs
            // to the next thread, which reads HF)
            v[j][i] = fma(-th.dtmc[j], hf, v[j][i]);
            if (j + 1 < C) v[(j + 1) % C][i] = fma(-th.dtmc[(j + 1) % C], hf, v[(j + 1) % C][i]);
            if (kLast) {
                r_fext[j][i] = (j == 0) ? -hf : r_fext[j][i] - hf;
                if (j + 1 < C) r_fext[(j + 1) % C][i] = -hf;
            }
        }
        // torques use the spring part only; on rod two: -(rd2 x k d), still in the world frame
        tw[j][0] = fma(rd2[2], fs[1], -rd2[1] * fs[2]);
        tw[j][1] = fma(rd2[0], fs[2], -rd2[2] * fs[0]);
        tw[j][2] = fma(rd2[1], fs[0], -rd2[0] * fs[1]);
    }
    tile_arrive<kSplit>(bars.b3);

    // ---- work that needs nobody's rows, while the others finish their joints -----------------------------------
    // one tip-to-base joint COMPONENT per designated thread (Appendix A.8: un-rounded spring between the two element
    // surfaces along fixed material directions, un-projected damping, no torque), from the pushed data of this step
    if (kExtras) {
        if (th.xjoint >= 0) {
            const int xj = th.xjoint >> 2, i = th.xjoint & 3;
            const int2 ja = *(const int2 *)(XJI + xj * 8), jb = *(const int2 *)(XJI + xj * 8 + 2);
            const double *jd = XJD + xj * 8;
            double f;
            if (XJI[xj * 8 + 4] == 1) {
                // point force (custom_activation.PointForces): F min(1, t / ramp_up_time) on one node
                const double fac = fmin(1.0, time_mid / jd[3]);
                f = jd[i] * fac;
            } else {
                if (kCluster && p.tile.xb_count[rank] > 0) mbar_wait(xbar, xparity);  // what other CTAs push has landed (own elements: B2)
                const double *e1 = EXD + ja.x * L::EXW, *e2 = EXD + ja.y * L::EXW;
                const double *q1 = EXD + jb.x * L::EXW, *q2 = EXD + jb.y * L::EXW;
                const double kj = jd[0], nuj = jd[1];
                const double c1 = 0.5 * e1[13 + i], c2 = 0.5 * e2[13 + i];
                const double vrel = 0.5 * e2[16 + i] - 0.5 * e1[16 + i];
                const double a1 = fma(q1[i], jd[2], fma(q1[3 + i], jd[3], q1[6 + i] * jd[4])) * e1[12];
                const double a2 = fma(q2[i], jd[5], fma(q2[3 + i], jd[6], q2[6 + i] * jd[7])) * e2[12];
                const double gap = (c2 + a2) - (c1 + a1);
                f = 0.5 * fma(kj, gap, -nuj * vrel);
            }
            // into the rows of the joint's nodes that live in this CTA: rod one's take +F/2, rod two's -F/2
            const int2 ta = *(const int2 *)(XJT + xj * 4), tb2 = *(const int2 *)(XJT + xj * 4 + 2);
            if (ta.x >= 0) NF[ta.x + i] = f;
            if (ta.y >= 0) NF[ta.y + i] = f;
            if (tb2.x >= 0) NF[tb2.x + i] = -f;
            if (tb2.y >= 0) NF[tb2.y + i] = -f;
        }
        if (kSplit) mbar_arrive_warp(bars.b4);  // (CTA barriers: B3 below orders the joint forces before their readers)
    }
    // damper exponentials (element lengths are mine): before the wait when the barrier is split-phase
    constexpr bool kHoistDamper = kSplit && BR2_TILE_HOIST_DAMPER;
    double dexp[C][3];
    if (kHoistDamper) {
#pragma unroll
        for (int j = 0; j < C; ++j) tile_damper_exponentials<kExact, kPerInst>(p, smu, len[j], th.fl[j], dexp[j], th.bad);
    }
    if (kSplit)
        mbar_wait(bars.b3, (uint32_t)gstep & 1u);  // B3
    else
        __syncthreads();  // B3

    // =============================== P4: assemble, dynamic step ==============================
    {
        const double ramp_raw = time_mid * tb.act_inv_ramp;
        const double ramp = (ramp_raw < 1.0) ? ramp_raw : 1.0;
        // what the previous run's last element contributes to my first node / element (rows: see TileSmem)
        const int tprev = t - 1, top = (th.to > 0) ? th.to - 1 : 0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const double fp = sm[L::HF + (3 * (C - 1) + i) * R + top] - sm[L::HF + (3 * (C - 1) + i) * R + tprev];
            const double sp = sm[L::SKL + i * R + tprev];
            v[0][i] = fma(th.dtmc[0], fp - sp, v[0][i]);
            tint[0][i] += sm[L::ML + i * R + tprev];
            if (kLast) {
                r_fint[0][i] -= sp;
                r_fext[0][i] += fp;
            }
        }
#pragma unroll
        for (int j = 0; j < C; ++j) {
            const int k = wh.k0 + j;
            double text[3];
            {
                double fso[3], rd1[3];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    // rod one takes +F of the joint its partner owns: half on each node of the element
                    const double fo = sm[L::HF + (3 * j + i) * R + th.to];
                    v[j][i] = fma(th.dtmc[j], fo, v[j][i]);
                    if (j + 1 < C) v[(j + 1) % C][i] = fma(th.dtmc[(j + 1) % C], fo, v[(j + 1) % C][i]);
                    if (kLast) {
                        r_fext[j][i] += fo;
                        if (j + 1 < C) r_fext[(j + 1) % C][i] += fo;
                    }
                    fso[i] = sm[L::FS + (3 * j + i) * R + th.to];
                    if (!kAosPub) rd1[i] = sm[L::PR + (3 * j + i) * R + t];
                }
                if (kAosPub) {
                    const double2 *const mine = (const double2 *)(sm + L::PC + (j * R + t) * 10);
                    const double2 m3 = mine[3], m4 = mine[4];
                    rd1[0] = m3.x, rd1[1] = m3.y, rd1[2] = m4.x;
                }
                // rod-one side of the joint owned by my partner: +(rd1 x k d); then both into my material frame
                if (kLeanFma & 1) {  // the cross product accumulated straight onto the own joint's torque: two FMAs per component
                    tw[j][0] = fma(rd1[1], fso[2], fma(-rd1[2], fso[1], tw[j][0]));
                    tw[j][1] = fma(rd1[2], fso[0], fma(-rd1[0], fso[2], tw[j][1]));
                    tw[j][2] = fma(rd1[0], fso[1], fma(-rd1[1], fso[0], tw[j][2]));
                } else {
                    tw[j][0] += fma(rd1[1], fso[2], -rd1[2] * fso[1]);
                    tw[j][1] += fma(rd1[2], fso[0], -rd1[0] * fso[2]);
                    tw[j][2] += fma(rd1[0], fso[1], -rd1[1] * fso[0]);
                }
            }
            // (lean: the internal couple rides at the end of the rotation's FMA chain, so tsum = tint + text costs no addition;
            // the launch's last step, which reports the two separately, keeps them apart)
            constexpr bool kSum = (kLeanFma & 2) && !kLast;
#pragma unroll
            for (int i = 0; i < 3; ++i)
                text[i] = fma(Q[j][3 * i], tw[j][0], fma(Q[j][3 * i + 1], tw[j][1],
                              kSum ? fma(Q[j][3 * i + 2], tw[j][2], tint[j][i]) : Q[j][3 * i + 2] * tw[j][2]));
#pragma unroll
            for (int i = 0; i < 3; ++i) text[i] = fma(TAV(j, i), ramp, text[i]);
            if (kExtras && (th.xinfo[j] >> 8) != 0) {
                if (kLast) {  // half forces of the extra joints on this node (every other step: before the rotations below)
                    const int cnt = (th.xinfo[j] >> 16) & 0xF;
                    const double *nf = NF + (th.xinfo[j] >> 20) * nf_row;
                    if (kSplit && cnt) mbar_wait(bars.b4, (uint32_t)gstep & 1u);
                    for (int q = 0; q < cnt; ++q) {
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            const double f = nf[4 * q + i];
                            v[j][i] = fma(th.dtmc[j], f, v[j][i]);
                            r_fext[j][i] += f;
                        }
                    }
                }
                // a tip-to-base joint overwrites the downstream base element's director and omega during synchronize
                // (after the surface joints above have used the old director); sources are start-of-step values
                const int ow = ((th.xinfo[j] >> 8) & 0xFF) - 1;
                if (ow >= 0) {
                    if (kCluster && p.tile.xb_count[rank] > 0) mbar_wait(xbar, xparity);
                    const double *src = EXD + ow * L::EXW;
#pragma unroll
                    for (int i = 0; i < 9; ++i) Q[j][i] = src[i];
#pragma unroll
                    for (int i = 0; i < 3; ++i) w[j][i] = src[9 + i];
                }
            }
            const double dil = len[j] * UNI(BR2_FC_INV_REST_LEN);
            double damp[3];
            if (kHoistDamper) {
                if (p.tile.round_section) {
                    const double ex = dexp[j][0];
                    damp[2] = UNI(BR2_FC_CR + 2) * ex;
                    damp[0] = damp[1] = UNI(BR2_FC_CR) * (ex * ex);
                } else {
#pragma unroll
                    for (int i = 0; i < 3; ++i) damp[i] = UNI(BR2_FC_CR + i) * dexp[j][i];
                }
            } else {
                const double strain = (th.fl[j] & TILE_F_ELEM) ? dil - 1.0 : 0.0;
                if (p.tile.round_section) {
                    const double e2 = strain * UNI(BR2_FC_LN_CR + 2);
                    if (!kExact) {
                        if (kMergeFlags) ok_e = ok_e && (e2 * e2 <= BR2_E2_MAX);
                        else th.bad |= (e2 * e2 <= BR2_E2_MAX) ? 0 : 4;
                    }
                    const double ex = M::exp(e2);
#if BR2_TILE_SCALAR_MASKS
                    // an element whose angular velocity is held at zero (and the node-only slot): the mask goes onto the one
                    // exponential instead of onto the three results (the bracket below is finite, so 0 * it = 0)
                    const double ex_m = (th.fl[j] & TILE_F_ZERO_W) ? 0.0 : ex;
                    damp[2] = UNI(BR2_FC_CR + 2) * ex_m;
                    damp[0] = damp[1] = UNI(BR2_FC_CR) * (ex_m * ex);
#else
                    damp[2] = UNI(BR2_FC_CR + 2) * ex;
                    damp[0] = damp[1] = UNI(BR2_FC_CR) * (ex * ex);
#endif
                } else {
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const double e = strain * UNI(BR2_FC_LN_CR + i);
                        if (!kExact) {
                            if (kMergeFlags) ok_e = ok_e && (e * e <= BR2_E2_MAX);
                            else th.bad |= (e * e <= BR2_E2_MAX) ? 0 : 4;
                        }
                        damp[i] = UNI(BR2_FC_CR + i) * M::exp(e);
                        if (BR2_TILE_SCALAR_MASKS) damp[i] = (th.fl[j] & TILE_F_ZERO_W) ? 0.0 : damp[i];
                    }
                }
            }
            double wnew[3];
#pragma unroll
            for (int i = 0; i < 3; ++i)
                wnew[i] = fma(KD(8 + i, dt * UNI(BR2_FC_JINV + i)) * dil, kSum ? text[i] : tint[j][i] + text[i], w[j][i]) * damp[i];
            if (kLast && (th.fl[j] & TILE_F_VALID)) {
                // quantities the reference's callbacks report (free_simulator.py:56-72), as of this step
                const double mass = tb.slot_c[(long long)BR2_SC_MASS * tb.n_slots + wh.s0 + j];
                double a[3], alpha[3], fext[3];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    fext[i] = fma(UNI(BR2_FC_GRAVITY + i), mass, r_fext[j][i]);
                    a[i] = (r_fint[j][i] + fext[i]) / mass;
                    alpha[i] = UNI(BR2_FC_JINV + i) * (tint[j][i] + text[i]) * dil;
                }
                if (th.fl[j] & TILE_F_SOFTFIX) {
                    // FreeBaseEndSoftFixed's spring acts on node 0 in both constrain_values calls of the step;
                    // it never changes the motion (node 0's rates are zeroed) and is only visible here.
                    const double *cd = tb.cons_d + tb.fast_i[(wh.s0 + j) * BR2_FAST_NI + BR2_FAST_SOFTFIX_ROW] * BR2_C_ND;
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const double spring = cd[12] * (cd[i] - x[j][i]);
                        fext[i] = (fext[i] + spring) + spring;
                    }
                }
                if (th.fl[j] & TILE_F_REPORT0) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) a[i] = alpha[i] = 0.0;
                }
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    inst[6 * np1 + 12 * n + i * np1 + k] = a[i];
                    inst[9 * np1 + 15 * n + i * np1 + k] = r_fint[j][i];
                    inst[12 * np1 + 18 * n + i * np1 + k] = fext[i];
                }
                if (th.fl[j] & TILE_F_ELEM) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        inst[9 * np1 + 12 * n + i * n + k] = alpha[i];
                        inst[12 * np1 + 15 * n + i * n + k] = tint[j][i];
                        inst[15 * np1 + 18 * n + i * n + k] = text[i];
                    }
                }
            }
            // rate constraints
#pragma unroll
            for (int i = 0; i < 3; ++i)  // (scalar masks: already zero through damp[])
                w[j][i] = (BR2_TILE_SCALAR_MASKS && !kHoistDamper) ? wnew[i] : ((th.fl[j] & TILE_F_ZERO_W) ? 0.0 : wnew[i]);
        }
        // node rates are complete only now (the element after a node lives in the same thread, later in the loop):
        // constraints, then the trailing kinematic half step of this step fused with the leading half step of the
        // next one (unless this is the last step of the launch)
        if (kExtras && !kLast) {
            // half forces of the tip-to-base joints on my nodes: after the dynamic step, so that their evaluation (by other
            // warps, between B3's arrive and wait) has long finished, and BEFORE the rotations, whose long dependent chains
            // cover the latency of these few loads (at the very end of the step it was exposed: +1 ... 2 % measured; fetching
            // them at the start of P4 in the CTA-barrier kernels is no better)
            bool any = false;
#pragma unroll
            for (int j = 0; j < C; ++j) any = any || ((th.xinfo[j] >> 16) & 0xF) != 0;
#if BR2_TILE_GATHER_MERGED
            // One branch per WARP (most warps of a large CTA hold no rod end), and inside it both slots' rows without a branch
            // between them: the 18 loads are in flight together and the additions of a slot without items are masked (s = 0),
            // instead of two divergent blocks with a load-to-use bubble each.  Same order of summation per node.
            if (__any_sync(0xffffffffu, any)) {
                if (kSplit && any) mbar_wait(bars.b4, (uint32_t)gstep & 1u);
                int cnt[C];
                double f[C][3][3];
#pragma unroll
                for (int j = 0; j < C; ++j) {
                    cnt[j] = (th.xinfo[j] >> 16) & 0xF;
                    const double *nf = NF + (th.xinfo[j] >> 20) * nf_row;  // (row 0 for a slot without items: finite, masked)
#pragma unroll
                    for (int q = 0; q < 3; ++q)
#pragma unroll
                        for (int i = 0; i < 3; ++i) f[j][q][i] = nf[4 * min(q, max(cnt[j] - 1, 0)) + i];
                }
#pragma unroll
                for (int q = 0; q < 3; ++q)
#pragma unroll
                    for (int j = 0; j < C; ++j) {
                        const double s = (q < cnt[j]) ? th.dtmc[j] : 0.0;
#pragma unroll
                        for (int i = 0; i < 3; ++i) v[j][i] = fma(s, f[j][q][i], v[j][i]);
                    }
#pragma unroll
                for (int j = 0; j < C; ++j) {
                    const double *nf = NF + (th.xinfo[j] >> 20) * nf_row;
                    for (int q = 3; q < cnt[j]; ++q)
#pragma unroll
                        for (int i = 0; i < 3; ++i) v[j][i] = fma(th.dtmc[j], nf[4 * q + i], v[j][i]);
                }
            }
#else
            if (kSplit && any) mbar_wait(bars.b4, (uint32_t)gstep & 1u);
#pragma unroll
            for (int j = 0; j < C; ++j) {
                const int cnt = (th.xinfo[j] >> 16) & 0xF;
                if (cnt) {
                    // the node's row: item q at 4 q (a joined element's node has three items, one per rod of the other segment);
                    // the first three without a branch between them, same order of summation as the loop
                    const double *nf = NF + (th.xinfo[j] >> 20) * nf_row;
                    double f[3][3];
#pragma unroll
                    for (int q = 0; q < 3; ++q)
#pragma unroll
                        for (int i = 0; i < 3; ++i) f[q][i] = nf[4 * min(q, cnt - 1) + i];
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        const double s = (q < cnt) ? th.dtmc[j] : 0.0;
#pragma unroll
                        for (int i = 0; i < 3; ++i) v[j][i] = fma(s, f[q][i], v[j][i]);
                    }
                    for (int q = 3; q < cnt; ++q)
#pragma unroll
                        for (int i = 0; i < 3; ++i) v[j][i] = fma(th.dtmc[j], nf[4 * q + i], v[j][i]);
                }
            }
#endif
        }
        const double hmult = kEnd ? hdt : dt;
        bool bad_r = false;
#pragma unroll
        for (int j = 0; j < C; ++j) {
            const bool out_of_range = M::rotate(Q[j], w[j], hdt, hmult);
            if (kMergeFlags) bad_r = out_of_range || bad_r;
            else th.bad |= out_of_range ? 1 : 0;
        }
        if (kMergeFlags) th.bad |= (bad_r ? 1 : 0) | (ok_y ? 0 : 2) | (ok_e ? 0 : 4);
#pragma unroll
        for (int j = 0; j < C; ++j) {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                v[j][i] = (th.fl[j] & TILE_F_ZERO_V) ? 0.0 : v[j][i];
                x[j][i] = fma(hmult, v[j][i], x[j][i]);
            }
        }
    }
#undef UNI
#undef KD
#undef JDV
#undef TAV
}

template <int C, int NT, int kMinBlocks, bool kExtras, bool kCluster, bool kExact, bool kSplit, bool kPerInst = false>
__global__ void __launch_bounds__(NT, kMinBlocks) br2_step_tile_kernel(const __grid_constant__ StepParams p) {
    using L = TileSmem<C, NT>;
    constexpr int R = L::R;
    extern __shared__ __align__(16) double sm[];
    __shared__ long long s_item;
    __shared__ int s_chunk;
    __shared__ int s_why;
    __shared__ unsigned long long s_bar[6];  // B1, B2, B3, B4 (joint forces ready), XB[2] (pushed element data, by step parity)
    // kPerInst: this instance's constants (uni, then kd), one block per class of rods
    __shared__ double s_uni[kPerInst ? BR2_TILE_MAX_CLS * (BR2_FC_COUNT + BR2_KD_COUNT) : 1];

    const DevTables &tb = p.t;
    const TilePlan &tp = p.tile;
    // Which logical warp a hardware warp executes (TilePlan::warp_map; identity unless the plan says otherwise): the warp
    // scheduler a warp runs on is (hardware warp index % 4), and a 10-warp CTA alone on an SM loads the four schedulers
    // 3, 3, 2, 2 -- per-warp clock64 timing shows a warp ~40 % slower on a 3-warp scheduler.  The plan puts the warps that
    // carry the rods' ends (the tip-to-base protocol) on the 2-warp schedulers, where the slack absorbs it.
    const int t = kCluster ? ((tp.warp_map[threadIdx.x >> 5] << 5) | (threadIdx.x & 31)) : (int)threadIdx.x;
    // kCluster: the CTAs of a thread-block cluster integrate one instance together, one group of ring-connected
    // rods (a segment) each; they exchange the tip-to-base joints' data through distributed shared memory (asynchronous
    // pushes that count themselves off the consumer's mbarrier, see tile_step) and meet at a cluster barrier only between work items.  The cluster, not the CTA,
    // pulls work items.
    const int rank = kCluster ? (int)cooperative_groups::this_cluster().block_rank() : 0;
    const long long total_items = p.batch * (long long)p.n_chunks;
    // exact-path pass: nothing to do unless the optimistic pass just before it flagged an instance
    if (kExact && *(volatile unsigned long long *)p.fail_counter == 0ULL) return;
    TileBars bars;
    bars.b1 = smem_addr(&s_bar[0]);
    bars.b2 = bars.b1 + 8;
    bars.b3 = bars.b1 + 16;
    bars.b4 = bars.b1 + 24;
    bars.xb = bars.b1 + 32;
    if (kSplit || kExtras) {
        if (t == 0) {
            for (int i = 0; i < 4; ++i) mbar_init(bars.b1 + 8 * i, NT / 32);
            mbar_init(bars.xb, 1);  // one arming arrive per step (expect_tx); the pushes count bytes, not arrivals
            mbar_init(bars.xb + 8, 1);
            if (kCluster) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        // (visible to the CTA, and to the cluster, at the first tile_sync of the item loop)
    }
    int gstep = 0;  // steps executed so far: the barriers' phase parity

    // Persistent CTAs pull (chunk, instance) items off one queue: the launch is cut into n_chunks runs of chunk_steps
    // steps so that the last wave of instances does not leave most SMs idle (4096 instances on 888 resident CTAs would
    // otherwise cost 5 full waves for 4.6 waves of work).  Order: instances in groups of p.queue_group (about two waves of
    // CTAs), chunk-major inside a group, group after group -- so the state a chunk writes back is read again `queue_group`
    // items later, while it still sits in the L2 (a group's state is ~32 MB), instead of a whole batch later.  Chunk c of an
    // instance starts when chunk c-1 of the same instance has been written back (progress[instance] == c); that was two
    // waves ago, so nobody actually waits unless the batch is tiny.
    for (;;) {
        if (rank == 0 && t == 0) {
            const long long next = (long long)atomicAdd(p.work_counter, 1ULL);
            long long packed = -1;
            int chunk_of = 0;
            if (next < total_items) {
                const long long per_group = p.queue_group * (long long)p.n_chunks;
                const long long g = next / per_group, r = next - g * per_group;
                const long long first = g * p.queue_group;
                const long long size = (p.batch - first < p.queue_group) ? p.batch - first : p.queue_group;  // the last group may be short
                const long long c = r / size, bb = first + (r - c * size);
                while (*(volatile int *)(p.progress + bb) < (int)c) __nanosleep(100);
                __threadfence();
                // (how the pair travels is a measured choice: the single-segment kernel is 1.2 % faster with two words, the
                // multi-segment kernels 1.3 % faster with one packed word -- the code around the step loop shifts ptxas' schedule)
                packed = kExtras ? (bb | (c << 48)) : bb;
                chunk_of = (int)c;
            }
            for (int r = 0; r < (kCluster ? tp.n_ctas : 1); ++r) {
                *tile_peer<kCluster>(&s_item, r) = packed;
                if (!kExtras) *tile_peer<kCluster>(&s_chunk, r) = chunk_of;
            }
            s_why = 0;
        }
        tile_sync<kCluster>();
        const long long item = s_item;
        tile_sync<kCluster>();
        if (item < 0) break;
        const long long b = kExtras ? (item & ((1LL << 48) - 1)) : item;
        const int chunk = kExtras ? (int)(item >> 48) : s_chunk;
        {
            // optimistic pass: an instance waiting for the exact-path replay skips the rest of the launch;
            // exact pass: only flagged instances, from the chunk in which they were flagged (status = 1 + chunk | why << 8)
            const int st = *(volatile int *)(p.status + b);
            const bool skip = kExact ? (st == 0 || chunk < (st & 0xFF) - 1) : (st != 0);
            if (skip) {
                if (rank == 0 && t == 0) *(volatile int *)(p.progress + b) = chunk + 1;
                continue;
            }
        }
        const int first_step = chunk * p.chunk_steps;
        const int n_steps = min(p.chunk_steps, (int)p.n_steps - first_step);
        const bool final_chunk = (chunk == p.n_chunks - 1);

        TileThread<C> th;
        th.t = t + 1;
        th.bad = 0;
        // ---- which run is mine -------------------------------------------------------------------
        const bool active = t < tp.cta_threads[rank];
        const int S = tb.n_slots;
        {
            const TileWhere wh = tile_where<C>(p, t, b, rank);
            const int rod = wh.rod, n = wh.n, k0 = wh.k0;

            // rows I read from: next / previous run of my rod, my partner as rod one of the joints I own, the owner
            // of the joints in which I am rod one (and the owner of my predecessor's); the zero row when there is none
            th.tn = (active && k0 + C <= n) ? t + 2 : t + 1;
            const int r1 = tp.partner[rod], ro = tp.owner[rod];
            th.t1 = (active && r1 >= 0) ? 1 + tp.rod_t0[r1] + (t - tp.rod_t0[rod]) : 0;
            th.to = (active && ro >= 0) ? 1 + tp.rod_t0[ro] + (t - tp.rod_t0[rod]) : 0;

            const double dt = p.dt;
            const int np1 = n + 1;
            const double *const inst = wh.inst;  // read through L2 (__ldcg): another SM wrote the previous chunk
#pragma unroll
            for (int j = 0; j < C; ++j) {
                const int k = k0 + j;
                const bool valid = k <= n, elem = k < n, vor = k < n - 1;
                const int sc = valid ? wh.s0 + j : S - 1;
                const int *fi = tb.fast_i + sc * BR2_FAST_NI;
                const int pf = valid ? fi[BR2_FAST_FLAGS] : 0;
                int f = (valid ? TILE_F_VALID : 0) | (elem ? TILE_F_ELEM : 0) | (vor ? TILE_F_VOR : 0);
                if (!valid || (pf & BR2_FAST_ZERO_NODE_RATE)) f |= TILE_F_ZERO_V;
                if (!elem || (pf & BR2_FAST_ZERO_ELEM_RATE)) f |= TILE_F_ZERO_W;
                if (elem && fi[BR2_FAST_OWN_JOINT] >= 0) f |= TILE_F_OWN;
                if (pf & BR2_FAST_ZERO_REPORTED) f |= TILE_F_REPORT0;
                if (valid && fi[BR2_FAST_SOFTFIX_ROW] >= 0) f |= TILE_F_SOFTFIX;
                th.fl[j] = f;
                th.xinfo[j] = (kExtras && tp.n_xj > 0 && active) ? tp.xi[(rank * tp.xi_stride + t) * (C + 1) + j] : 0;
                const double c_t = kPerInst ? p.inst_uni[(b * p.inst_stride + tp.rod_cls[rod]) * BR2_FC_COUNT + BR2_FC_CT] : p.uni[BR2_FC_CT];
                th.dtmc[j] = dt * c_t / tb.slot_c[(long long)BR2_SC_MASS * S + sc];
                double ta[3] = {0.0, 0.0, 0.0};
                if (valid) {
                    for (int r = fi[BR2_FAST_ACT_BEGIN]; r < fi[BR2_FAST_ACT_END]; ++r) {
                        const double P = p.pressures[(long long)tb.act_i[r] * p.batch + b];
#pragma unroll
                        for (int i = 0; i < 3; ++i)
                            ta[i] = fma(P, (kPerInst && p.inst_act_d) ? p.inst_act_d[(b * p.n_act + r) * 3 + i] : tb.act_d[r * 3 + i], ta[i]);
                    }
                }
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    sm[L::TA + (3 * j + i) * R + t + 1] = ta[i];
                    th.x[j][i] = valid ? __ldcg(inst + i * np1 + k) : 0.0;
                    th.v[j][i] = valid ? __ldcg(inst + 3 * np1 + i * np1 + k) : 0.0;
                    th.w[j][i] = elem ? __ldcg(inst + 6 * np1 + 9 * n + i * n + k) : 0.0;
                }
#pragma unroll
                for (int i = 0; i < 9; ++i) th.Q[j][i] = elem ? __ldcg(inst + 6 * np1 + i * n + k) : ((i % 4 == 0) ? 1.0 : 0.0);
                // pinned values are constants of the launch: set once; their rates are zero from here on
                const double *FCg = tb.fast_c + sc;
                if (pf & BR2_FAST_PIN_POSITION) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        th.x[j][i] = FCg[(long long)(BR2_FC_PIN + i) * S];
                        th.v[j][i] = 0.0;
                    }
                }
                if (pf & BR2_FAST_PIN_DIRECTOR) {
#pragma unroll
                    for (int i = 0; i < 9; ++i) th.Q[j][i] = FCg[(long long)(BR2_FC_PIN + 3 + i) * S];
#pragma unroll
                    for (int i = 0; i < 3; ++i) th.w[j][i] = 0.0;
                }
            }
            // (a cluster's idle lanes may evaluate joint components too: the plan fills the last warp)
            const bool in_last_warp = kCluster && t < 32 * ((tp.cta_threads[rank] + 31) / 32);
            th.xjoint = (kExtras && tp.n_xj > 0 && (active || in_last_warp)) ? tp.xi[(rank * tp.xi_stride + t) * (C + 1) + C] : -1;
            if (kExtras) {  // this CTA's copy of the extra joints' tables (layout: TileSmem)
                double *const XJD = sm + L::EX + 2 * L::EXW * tp.n_xe + 4 * tp.nf_items * tp.n_nr;
                int *const XJI = (int *)(XJD + 8 * tp.n_xj), *const XJT = XJI + 8 * tp.n_xj, *const XPM = XJT + 4 * tp.n_xj;
                for (int i = t; i < 8 * tp.n_xj; i += NT) {
                    // (a tip-to-base joint's stiffness is proportional to Young's modulus: free_simulator.py:220-228)
                    const bool stiffness = kPerInst && p.inst_joint_scale && (i & 7) == 0 && tp.xj_i[12 * (i >> 3) + 10] == 0;
                    XJD[i] = stiffness ? tp.xj_d[i] * p.inst_joint_scale[b] : tp.xj_d[i];
                }
                for (int i = t; i < tp.n_xj; i += NT) {
                    const int *row = tp.xj_i + 12 * i;
                    XJI[8 * i + 0] = row[8];   // published element of rod one's element (node sums, radius)
                    XJI[8 * i + 1] = row[9];   // ... of rod two's element
                    XJI[8 * i + 2] = row[6];   // ... of director one
                    XJI[8 * i + 3] = row[7];   // ... of director two
                    XJI[8 * i + 4] = row[10];  // type: 0 tip-to-base joint, 1 point force
                }
                for (int i = t; i < tp.n_xe; i += NT) XPM[i] = tp.xe_mask[i];
                for (int i = t; i < 4 * tp.n_xj; i += NT) XJT[i] = tp.xjt[(size_t)rank * 4 * tp.n_xj + i];
            }
            th.rod = rod;  // indexes the per-rod joint descriptors in the constant bank (TilePlan::jd)
            th.uoff = kPerInst ? tp.rod_cls[rod] * (BR2_FC_COUNT + BR2_KD_COUNT) : 0;
        }
        // zero the exchange arrays (the zero row stays zero for the whole chunk)
        for (int i = t; i < L::ZERO_END; i += NT) sm[i] = 0.0;
        if (kPerInst) {
            // this instance's constants, and the products br2_step forms on the host for the shared ones (same expressions)
            // (one block per class of rods with equal constants; inst_stride = 0: every instance reads the same table)
            for (int i = t; i < tp.n_cls * BR2_FC_COUNT; i += NT) {
                const int c = i / BR2_FC_COUNT, k = i - c * BR2_FC_COUNT;
                s_uni[c * (BR2_FC_COUNT + BR2_KD_COUNT) + k] = p.inst_uni[(b * p.inst_stride + c) * BR2_FC_COUNT + k];
            }
            for (int tt = t; tt < BR2_KD_COUNT * tp.n_cls; tt += NT) {
                const int c = tt / BR2_KD_COUNT, t = tt - BR2_KD_COUNT * c;
                const double *u = p.inst_uni + (b * p.inst_stride + c) * BR2_FC_COUNT;
                const double dt = p.dt;
                double kd = 0.5 * dt;                                                   // [11]
                if (t < 3) kd = dt * u[BR2_FC_CT] * u[BR2_FC_GRAVITY + t];
                else if (t == 3) kd = 0.5 * u[BR2_FC_INV_REST_VLEN];
                else if (t == 4) kd = -0.5 * u[BR2_FC_INV_REST_VLEN];
                else if (t == 5) kd = -0.25 * u[BR2_FC_JNU];
                else if (t == 6) kd = 0.5 * u[BR2_FC_JK];
                else if (t == 7) kd = -0.5 * u[BR2_FC_JKREP];
                else if (t < 11) kd = dt * u[BR2_FC_JINV + (t - 8)];
                else if (t >= 12) kd = u[BR2_FC_SHEAR + (t - 12)] * u[BR2_FC_INV_REST_LEN];
                s_uni[c * (BR2_FC_COUNT + BR2_KD_COUNT) + BR2_FC_COUNT + t] = kd;
            }
        }
        tile_sync<kCluster>();

        double time = p.chunk_t0[chunk];
        const double hdt = 0.5 * p.dt;

        // leading kinematic half step of the first step; every later one is fused with the trailing half step
        // of the step before it (same velocities, same omega)
#pragma unroll
        for (int j = 0; j < C; ++j) {
#pragma unroll
            for (int i = 0; i < 3; ++i) th.x[j][i] = fma(hdt, th.v[j][i], th.x[j][i]);
            th.bad |= TileMath<kExact>::rotate(th.Q[j], th.w[j], hdt, hdt) ? 1 : 0;
        }
        for (int step = 0; step < n_steps - 1; ++step) {
            tile_step<C, NT, false, false, kExtras, kCluster, kExact, kSplit, kPerInst>(p, sm, s_uni + th.uoff, th, time + hdt, b, rank, gstep, bars);
            time = (time + hdt) + hdt;
            ++gstep;
        }
        if (final_chunk)
            tile_step<C, NT, true, true, kExtras, kCluster, kExact, kSplit, kPerInst>(p, sm, s_uni + th.uoff, th, time + hdt, b, rank, gstep, bars);
        else
            tile_step<C, NT, true, false, kExtras, kCluster, kExact, kSplit, kPerInst>(p, sm, s_uni + th.uoff, th, time + hdt, b, rank, gstep, bars);
        ++gstep;

        // a CTA that left the series' validity range keeps the chunk-start state and asks for the exact path, which
        // resumes this instance from the start of this chunk (status = 1 + chunk, reason bits from bit 8 up)
        int why = (__syncthreads_or(th.bad & 1) ? 1 : 0) | (__syncthreads_or(th.bad & 2) ? 2 : 0) | (__syncthreads_or(th.bad & 4) ? 4 : 0);
        if (kCluster) {  // the instance fails as a whole
            if (t == 0 && why) atomicOr(tile_peer<kCluster>(&s_why, 0), why);
            tile_sync<kCluster>();
            why = *tile_peer<kCluster>(&s_why, 0);
            tile_sync<kCluster>();  // everyone has read it before rank 0 clears it for the next item
        }
        if (why) {
            if (rank == 0 && t == 0) {
                atomicAdd(p.fail_counter, 1ULL);
                *(volatile int *)(p.status + b) = (1 + chunk) | (why << 8);
                __threadfence();
                *(volatile int *)(p.progress + b) = chunk + 1;
            }
            continue;
        }
        {
            const TileWhere wh = tile_where<C>(p, t, b, rank);
            const int n = wh.n, np1 = wh.n + 1;
            double *const inst = wh.inst;
#pragma unroll
            for (int j = 0; j < C; ++j) {
                const int k = wh.k0 + j;
                if (th.fl[j] & TILE_F_VALID) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        inst[i * np1 + k] = th.x[j][i];
                        inst[3 * np1 + i * np1 + k] = th.v[j][i];
                    }
                }
                if (th.fl[j] & TILE_F_ELEM) {
#pragma unroll
                    for (int i = 0; i < 9; ++i) inst[6 * np1 + i * n + k] = th.Q[j][i];
#pragma unroll
                    for (int i = 0; i < 3; ++i) inst[6 * np1 + 9 * n + i * n + k] = th.w[j][i];
                }
            }
        }
        __threadfence();
        tile_sync<kCluster>();
        if (rank == 0 && t == 0) {
            if (kExact && final_chunk) {  // replay complete: count it (and why it was needed), clear the flag
                const int st = *(volatile int *)(p.status + b) >> 8;
                atomicAdd(p.replays, 1ULL);
                for (int bit = 0; bit < 3; ++bit)
                    if (st & (1 << bit)) atomicAdd(p.replays + 1 + bit, 1ULL);
                *(volatile int *)(p.status + b) = 0;
            }
            *(volatile int *)(p.progress + b) = chunk + 1;
        }
    }
}

}  // namespace br2
