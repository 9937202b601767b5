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
// Semantics are those of br2_fast.cuh / br2_generic.cuh; the optimistic-series
// protocol is the same: a CTA that leaves a series' validity range flags its instance and does not write back;
// the exact-path pass launched right behind on the stream redoes the instance from the chunk that failed (the generic
// kernel, or this kernel's kExact instantiation for assemblies too large for it).
//
// Work distribution: persistent CTAs (or clusters) pull (chunk, instance) items off one queue -- see the kernel.
// Template flags: kExtras = tip-to-base joints between segments (director overwrites included); kCluster = one CTA per
// group of ring-connected rods, a thread-block cluster per instance (assemblies beyond one SM's registers);
// kExact = IEEE / libm arithmetic without validity ranges.
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
// split-phase cluster barrier: arrive (release) ... independent work ... wait (acquire)
template <bool kCluster>
__device__ __forceinline__ void tile_cluster_arrive() {
    if (kCluster) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
template <bool kCluster>
__device__ __forceinline__ void tile_cluster_wait() {
    if (kCluster) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <bool kCluster, typename T>
__device__ __forceinline__ T *tile_peer(T *local, int rank) {
    if (kCluster) return cooperative_groups::this_cluster().map_shared_rank(local, rank);
    return local;
}


// Tensor memory as per-thread scratch (kPark kernels).  This kernel has no use for the tensor cores, so the SM's 256 KB of
// tensor memory is idle; tcgen05.st / tcgen05.ld (32x32b shape: thread i of a warp <-> lane 32 * (warp % 4) + i, one
// 32-bit column per register) move registers there and back without touching the shared-memory crossbar or the L1
// (profiles/probes/tmem_scratch_probe.cu: 39-47 cycle round trip).  Values that only cross a phase are parked there so
// that the phase in between has their registers for instruction-level parallelism.
__device__ __forceinline__ void tmem_park(uint32_t addr, double v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(__double2loint(v)), "r"(__double2hiint(v)) : "memory");
}
__device__ __forceinline__ double tmem_fetch(uint32_t addr) {
    int lo, hi;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(addr) : "memory");
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
constexpr int kTileParkCols = 64;  // 32-bit columns per CTA (power of two >= 32): 32 doubles per thread
// kPark == 2 (C = 2): which double lives where.  x[0] and Q[0] are not parked: their published rows (XF, QF) are the copy.
constexpr int kPkX1 = 0, kPkV = 3, kPkW = 9, kPkQ1 = 15, kPkTA = 24;  // x[1] 3 | v 6 | w 6 | Q[1] 9 | actuation torque 6 = 30
#define TM_AT(idx) (tm + 2 * (idx))

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
    static __device__ __forceinline__ double exp(double d) { return kExact ? ::exp(d) : exp_tiny(d); }
    // theta / sin(theta + 1e-14) from y = 1 - cos(theta)
    static __device__ __forceinline__ double logmap_ratio(double y) {
        if (!kExact) return logmap_ratio_tiny(y);
        const double theta = acos(1.0 - y);
        return theta / sin(theta + 1e-14);
    }
    // Q <- R Q for one (hmult == hdt) or two fused (hmult == 2 hdt) kinematic half steps; true = left the series' range
    static __device__ __forceinline__ bool rotate(double (&Q)[9], const double (&w)[3], double hdt, double hmult) {
        if (!kExact) return rotate_directors_tiny(Q, w, hdt, hmult);
        const double u0 = hdt * w[0], u1 = hdt * w[1], u2 = hdt * w[2];
        const double theta = sqrt(u0 * u0 + u1 * u1 + u2 * u2);
        const double inv = 1.0 / (theta + 1e-14);
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

template <int C, int NT, bool kLean = false>
struct TileSmem {
    // rows of every exchange array: thread t owns row t + 1; row 0 is all zeros and never written.  A run without a
    // predecessor reads row (own - 1): row 0, or the previous rod's last run, whose "last element" entries are zero because
    // its last slot is the node-only slot.  A run without a successor reads its own row (the result is masked).  Both
    // keep a half-warp's addresses contiguous, i.e. free of bank conflicts.
    // kLean (the 128-register kernel, eight CTAs per SM): NT rows -- the CTA's idle threads (at least one: plan requirement)
    // take row 0 as their own, where their all-zero state publishes zeros in every array whose row 0 is read as "nothing
    // there" -- and no actuation rows (tensor memory holds them).
    static constexpr int R = kLean ? NT : NT + 1;
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
    static constexpr int ROWS_END = kLean ? TA : TA + 3 * C * R;
    // after the rows (host: TileSmem<C, 0>::ROWS_END * (threads + 1) words), the extra-joint area
    static constexpr int EX = ROWS_END;           // extra joints (rank 0's copy is the live one in a cluster):
                                                  // EXQ [2][9][n_xe], EXW [2][3][n_xe] (by step parity), EXR [n_xe], JF [3][n_xj]
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
    // extra joints (kExtras kernels): per slot published xe + 1 | overwrite source xe + 1 << 8 | node items << 16 |
    // first item << 20; the extra joint component this thread evaluates (joint << 2 | component) or -1
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

// One PositionVerlet step of a thread's run.  kEnd: last step of a chunk (ends with a single kinematic half step
// instead of the fused pair, so the state is a whole-step state); kLast: last step of the launch (also reports
// the callback quantities).
template <int C, int NT, bool kEnd, bool kLast, bool kExtras, bool kCluster, bool kExact, int kPark>
__device__ __forceinline__ void tile_step(const StepParams &p, double *sm, TileThread<C> &th, double time_mid, long long b,
                                          int rank, int parity, uint32_t tm) {
    using L = TileSmem<C, NT, kPark == 2>;
    constexpr int R = L::R;
    static_assert(kExtras || !kCluster, "the cluster kernels carry the extra-joint code (their barrier phases live in it)");
    using M = TileMath<kExact>;
    const DevTables &tb = p.t;
    const int t = th.t;
    const double dt = p.dt, hdt = p.kd[11];
    // loop-invariant products of the uniform constants and dt come from the constant bank (StepParams::kd, filled by br2_step
    // with the very expressions quoted here) instead of being hoisted into registers that then spill: +2-3 % on the
    // multi-segment kernels, neutral on the single-segment one, and what lets the 128-register build run without spills
#define KD(idx, expr) p.kd[(idx)]
    TileWhere wh{};
    if (kLast) wh = tile_where<C>(p, (int)threadIdx.x, b, rank);
    const int n = wh.n, np1 = wh.n + 1;
    double *const inst = wh.inst;
#define UNI(idx) p.uni[(idx)]
#define JDV(i) p.tile.jd[th.rod][(i)]
#define TAV(j, i) sm[L::TA + (3 * (j) + (i)) * (kPark == 2 ? 0 : R) + t]  // (kPark == 2: unused, tensor memory holds it)
    double (&x)[C][3] = th.x;
    double (&v)[C][3] = th.v;
    double (&Q)[C][9] = th.Q;
    double (&w)[C][3] = th.w;

    // reported quantities (kLast only)
    double r_fint[C][3], r_fext[C][3];

    // =============================== P1: publish the run's first node / element ==============
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        sm[L::XF + i * R + t] = x[0][i];
        sm[L::VF + i * R + t] = v[0][i];
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) sm[L::QF + i * R + t] = Q[0][i];
    // extra joints (tip-to-base joints between segments; kExtras kernels only): a handful of threads involved
    const int n_xe = kExtras ? p.tile.n_xe : 0, n_xj = kExtras ? p.tile.n_xj : 0;
    // director / omega rows are double-buffered by step parity: the overwrite in P4 reads them while a faster
    // thread may already be publishing the next step's values
    double *const EX0 = tile_peer<kCluster>(sm + L::EX, 0);
    double *const EXQ = EX0 + parity * 9 * n_xe, *const EXW = EX0 + 18 * n_xe + parity * 3 * n_xe;
    double *const EXR = EX0 + 24 * n_xe, *const JF = EXR + n_xe;
    // read-only copies of the joint tables (every CTA has its own): xj_d [n_xj][8], xj_i [n_xj][12], node items
    const double *const XJD = sm + L::EX + 25 * n_xe + 3 * n_xj;
    const int *const XJI = (const int *)(XJD + 8 * n_xj), *const XIT = XJI + 12 * n_xj;
    if (kExtras) {
#pragma unroll
        for (int j = 0; j < C; ++j) {
            const int pub = (th.xinfo[j] & 0xFF) - 1;
            if (pub >= 0) {
#pragma unroll
                for (int i = 0; i < 9; ++i) EXQ[i * n_xe + pub] = Q[j][i];
#pragma unroll
                for (int i = 0; i < 3; ++i) EXW[i * n_xe + pub] = w[j][i];
            }
        }
    }
    __syncthreads();  // B1 (the remote director rows become visible with the cluster barrier after P2)

    // =============================== P2: element quantities ==================================
    double len[C], ra2[C];
    double tint[C][3];   // internal couple on element k (complete except for the previous run's last cell)
    {
        const double dtg[3] = {KD(0, dt * UNI(BR2_FC_CT) * UNI(BR2_FC_GRAVITY)), KD(1, dt * UNI(BR2_FC_CT) * UNI(BR2_FC_GRAVITY + 1)),
                               KD(2, dt * UNI(BR2_FC_CT) * UNI(BR2_FC_GRAVITY + 2))};
        double skp[3];  // the previous element's Q^T n_L / e
#pragma unroll
        for (int j = 0; j < C; ++j) {
            const bool elem = th.fl[j] & TILE_F_ELEM;
            double d[3], dv[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const double xn = (j + 1 < C) ? x[(j + 1) % C][i] : sm[L::XF + i * R + th.tn];
                const double vn = (j + 1 < C) ? v[(j + 1) % C][i] : sm[L::VF + i * R + th.tn];
                d[i] = xn - x[j][i];
                dv[i] = vn - v[j][i];
                sm[L::PC + (3 * j + i) * R + t] = xn + x[j][i];
                sm[L::PV + (3 * j + i) * R + t] = vn + v[j][i];
            }
            double d2 = fma(d[0], d[0], fma(d[1], d[1], d[2] * d[2]));
            d2 = elem ? d2 : 1.0;  // keeps the node-only slot finite
            const double rs = M::rsqrt(d2);
            len[j] = fma(d2, rs, BR2_EPS14);
            const double inv_len = rs * fma(-BR2_EPS14, rs, 1.0);  // 1/(|d| + 1e-14), second order 1e-23
            const double rest_len = UNI(BR2_FC_REST_LEN), inv_rest_len = UNI(BR2_FC_INV_REST_LEN);
            const double r2 = UNI(BR2_FC_VOL_OVER_PI) * inv_len;
            const double rad = r2 * M::rsqrt(r2);
            const double dil = len[j] * inv_rest_len;
            const double inv_dil = rest_len * inv_len;
            // only thins the ~1e-8 m rest gap of a joint: 2^-20 relative is ample.  (1 / sqrt(e) is also radius x a constant;
            // folding it into a per-rod arm factor saves the seed and two instructions per slot and was measured SLOWER,
            // 3.12e10 against 3.17e10 el-steps/s: ptxas' schedule, not the instruction count, sets this kernel's pace.)
            const double rse = M::rsqrt_lo(dil);
            ra2[j] = fma(JDV(3), rse, rad);
            if (kExtras) {
                const int pub = (th.xinfo[j] & 0xFF) - 1;
                if (pub >= 0) EXR[pub] = rad;
            }
            if (kLast && elem) {
                inst[15 * np1 + 21 * n + wh.k0 + j] = len[j];
                inst[15 * np1 + 22 * n + wh.k0 + j] = dil;
                inst[15 * np1 + 23 * n + wh.k0 + j] = rad;
            }
            // sigma = e Q t - d3 = Q d / l_hat - d3 (the 1/|d| of the tangent cancels against the dilatation);
            // nL holds n_L / e, which is what both the nodal force Q^T n_L / e and the couple (Q t x n_L) l_hat need
            double qd[3], nL[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                qd[i] = fma(Q[j][3 * i], d[0], fma(Q[j][3 * i + 1], d[1], Q[j][3 * i + 2] * d[2]));
                const double sigma = fma(inv_rest_len, qd[i], (i == 2) ? -1.0 : 0.0);
                nL[i] = (UNI(BR2_FC_SHEAR + i) * inv_dil) * sigma;
            }
            // damping and gravity first (the update is linear), then the shear forces of elements k and k-1
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const double raw = fma(Q[j][i], nL[0], fma(Q[j][3 + i], nL[1], Q[j][6 + i] * nL[2]));
                const double sk = elem ? raw : 0.0;
                const double fi = (j == 0) ? sk : sk - skp[i];
                skp[i] = sk;
                v[j][i] = fma(th.dtmc[j], fi, fma(v[j][i], UNI(BR2_FC_CT), dtg[i]));
                if (kLast) r_fint[j][i] = fi;
            }
            {
                // d(e)/dt = (x_{k+1} - x_k).(v_{k+1} - v_k) / l / l_hat  (expanded form in the reference)
                const double edot = fma(d[0], dv[0], fma(d[1], dv[1], d[2] * dv[2])) * inv_len * inv_rest_len;
                double jw[3];
#pragma unroll
                for (int i = 0; i < 3; ++i) jw[i] = UNI(BR2_FC_J + i) * w[j][i] * inv_dil;
                const double ue = edot * inv_dil;
                tint[j][0] = fma(jw[1], w[j][2], fma(-jw[2], w[j][1], fma(jw[0], ue, fma(qd[1], nL[2], -qd[2] * nL[1]))));
                tint[j][1] = fma(jw[2], w[j][0], fma(-jw[0], w[j][2], fma(jw[1], ue, fma(qd[2], nL[0], -qd[0] * nL[2]))));
                tint[j][2] = fma(jw[0], w[j][1], fma(-jw[1], w[j][0], fma(jw[2], ue, fma(qd[0], nL[1], -qd[1] * nL[0]))));
            }
            // what the joint in which this element is rod one needs from it
            const double ra1 = fma(JDV(7), rse, rad);
            const double a0 = JDV(4) * ra1, a1 = JDV(5) * ra1;  // planar arms: no d3 component (tile plan requirement)
            sm[L::PA + j * R + t] = ra1;
#pragma unroll
            for (int i = 0; i < 3; ++i)
                sm[L::PR + (3 * j + i) * R + t] = fma(Q[j][i], a0, Q[j][3 + i] * a1);  // Q^T dv1 (r + o)
            if (kPark == 2) {
                // done with this slot until P4: its velocity and angular velocity go to tensor memory; slot 0 makes room for
                // slot 1's director (parked since the end of the previous step) -- Q[0] comes back from its published row in P3
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    tmem_park(TM_AT(kPkV + 3 * j + i), v[j][i]);
                    tmem_park(TM_AT(kPkW + 3 * j + i), w[j][i]);
                }
                if (j == 0) {
                    tmem_wait_st();
#pragma unroll
                    for (int i = 0; i < 9; ++i) Q[1][i] = tmem_fetch(TM_AT(kPkQ1 + i));
                    tmem_wait_ld();
                } else {
#pragma unroll
                    for (int i = 0; i < 3; ++i) tmem_park(TM_AT(kPkX1 + i), x[1][i]);
                }
            }
        }
        sm[L::LF + t] = len[0];
#pragma unroll
        for (int i = 0; i < 3; ++i) sm[L::SKL + i * R + t] = skp[i];
    }
    if (kPark == 1) {  // positions and angular velocities are not needed before P4: their registers serve P3
#pragma unroll
        for (int j = 0; j < C; ++j)
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                tmem_park(tm + 2 * (3 * j + i), x[j][i]);
                tmem_park(tm + 2 * (3 * C + 3 * j + i), w[j][i]);
            }
    }
    __syncthreads();  // B2
    // cluster: the other CTAs' rows are needed only by the tip-to-base joints at the end of P3 -- arrive now, wait there
    tile_cluster_arrive<kCluster>();

    // =============================== P3: Voronoi couples and own joints ======================
    double tw[C][3];  // world-frame joint torque on my element (own joint now, rod-one joint in P4)
    if (kPark == 2) {
#pragma unroll
        for (int i = 0; i < 9; ++i) Q[0][i] = sm[L::QF + i * R + t];  // my own published row
    }
    {
        double ap[3], cp[3];  // previous cell's A and C
#pragma unroll
        for (int j = 0; j < C; ++j) {
            const bool vor = th.fl[j] & TILE_F_VOR;
            double Qn[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) Qn[i] = (j + 1 < C) ? Q[(j + 1) % C][i] : sm[L::QF + i * R + th.tn];
            const double lenn = (j + 1 < C) ? len[(j + 1) % C] : sm[L::LF + th.tn];
            const double *q = Q[j];
            // antisymmetric part and trace of Qn Q^T
            const double w0 = fma(Qn[6], q[3], fma(Qn[7], q[4], fma(Qn[8], q[5], fma(-Qn[3], q[6], fma(-Qn[4], q[7], -Qn[5] * q[8])))));
            const double w1 = fma(Qn[0], q[6], fma(Qn[1], q[7], fma(Qn[2], q[8], fma(-Qn[6], q[0], fma(-Qn[7], q[1], -Qn[8] * q[2])))));
            const double w2 = fma(Qn[3], q[0], fma(Qn[4], q[1], fma(Qn[5], q[2], fma(-Qn[0], q[3], fma(-Qn[1], q[4], -Qn[2] * q[5])))));
            const double tr = fma(Qn[0], q[0], fma(Qn[1], q[1], fma(Qn[2], q[2], fma(Qn[3], q[3], fma(Qn[4], q[4],
                              fma(Qn[5], q[5], fma(Qn[6], q[6], fma(Qn[7], q[7], Qn[8] * q[8]))))))));
            // y = 1 - cos(theta) with the reference's -1e-10 guard; a benign value off the rod
            const double y = vor ? fma(-0.5, tr, kMisc[7]) : kMisc2[3];
            if (!kExact) th.bad |= (y <= BR2_Y_MAX && y > 0.0) ? 0 : 2;
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
    // own joints (Appendix A.7): my element is rod two; rod one's contribution comes from its published rows
#pragma unroll
    for (int j = 0; j < C; ++j) {
        const bool own = th.fl[j] & TILE_F_OWN;
        const double b0 = JDV(0) * ra2[j], b1 = JDV(1) * ra2[j];
        const double ra1 = sm[L::PA + j * R + th.t1];
        double rd2[3], cd2[3], dvec[3];
        double proj2 = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            rd2[i] = fma(Q[j][i], b0, Q[j][3 + i] * b1);  // Q^T dv2 (r + o)
            double xs;  // x_k + x_{k+1} of my element
            if (kPark) {
                xs = sm[L::PC + (3 * j + i) * R + t];
            } else {
                const double xn = (j + 1 < C) ? x[(j + 1) % C][i] : sm[L::XF + i * R + th.tn];
                xs = xn + x[j][i];
            }
            cd2[i] = xs - sm[L::PC + (3 * j + i) * R + th.t1];       // 2 (c2 - c1)
            const double g = rd2[i] - sm[L::PR + (3 * j + i) * R + th.t1];
            // (c2 + rd2) - (c1 + rd1), rounded to 12 decimals like np.round_(x, 12)
            dvec[i] = rint(fma(0.5, cd2[i], g) * BR2_1E12) * BR2_1EM12;
            const double vr2 = sm[L::PV + (3 * j + i) * R + t] - sm[L::PV + (3 * j + i) * R + th.t1];  // 2 v_rel
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
        const double pen = fmax(fma(-0.5 * cn2, rcn, ra1 + ra2[j]), 0.0);     // max(r1 + r2 - |cd|, 0)
        const double hmag = KD(7, -0.5 * UNI(BR2_FC_JKREP)) * (pen * M::sqrt_lo(pen)) * rcn;  // times cd2 = half the repulsion
        double fs[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const double h = fma(hcoef, dvec[i], hmag * cd2[i]);
            const double hf = own ? h : 0.0;
            fs[i] = own ? jk * dvec[i] : 0.0;
            sm[L::HF + (3 * j + i) * R + t] = hf;
            sm[L::FS + (3 * j + i) * R + t] = fs[i];
            // rod two takes -F: half on each node of the element (node k+1 of the run's last element belongs
            // to the next thread, which reads HF)
            if (kPark != 2) {  // (kPark == 2: the velocities are parked; P4 applies this force from the HF row)
                v[j][i] = fma(-th.dtmc[j], hf, v[j][i]);
                if (j + 1 < C) v[(j + 1) % C][i] = fma(-th.dtmc[(j + 1) % C], hf, v[(j + 1) % C][i]);
            }
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
    // one tip-to-base joint per designated thread (Appendix A.8): un-rounded spring between the two element
    // surfaces along fixed material directions, un-projected damping, no torque
    tile_cluster_wait<kCluster>();
    if (kExtras && th.xjoint >= 0) {
        // one COMPONENT of one joint per thread (three consecutive lanes per joint): the warp that holds them is late at
        // the next barrier by one component's chain of loads, not three
        const int xj = th.xjoint >> 2, i = th.xjoint & 3;
        {
            const int *ji = XJI + xj * 12;
            const double *jd = XJD + xj * 8;
            if (ji[10] == 1) {
                // point force (custom_activation.PointForces): F min(1, t / ramp_up_time) on one node
                const double fac = fmin(1.0, time_mid / jd[3]);
                JF[i * n_xj + xj] = jd[i] * fac;
            } else {
                const double *const sm1 = tile_peer<kCluster>(sm, ji[0]), *const sm2 = tile_peer<kCluster>(sm, ji[3]);
                const int row1 = ji[1] + 1, j1 = ji[2], row2 = ji[4] + 1, j2 = ji[5], q1 = ji[6], q2 = ji[7];
                const double kj = jd[0], nuj = jd[1];
                const double r1 = EXR[ji[8]], r2 = EXR[ji[9]];
                const double c1 = 0.5 * sm1[L::PC + (3 * j1 + i) * R + row1], c2 = 0.5 * sm2[L::PC + (3 * j2 + i) * R + row2];
                const double vrel = 0.5 * sm2[L::PV + (3 * j2 + i) * R + row2] - 0.5 * sm1[L::PV + (3 * j1 + i) * R + row1];
                const double a1 = fma(EXQ[i * n_xe + q1], jd[2], fma(EXQ[(3 + i) * n_xe + q1], jd[3], EXQ[(6 + i) * n_xe + q1] * jd[4])) * r1;
                const double a2 = fma(EXQ[i * n_xe + q2], jd[5], fma(EXQ[(3 + i) * n_xe + q2], jd[6], EXQ[(6 + i) * n_xe + q2] * jd[7])) * r2;
                const double gap = (c2 + a2) - (c1 + a1);
                JF[i * n_xj + xj] = 0.5 * fma(kj, gap, -nuj * vrel);
            }
        }
    }
    // cluster: the joint forces are gathered at the very end of P4 -- arrive now, wait there
    tile_cluster_arrive<kCluster>();
    __syncthreads();  // B3

    // =============================== P4: assemble, dynamic step ==============================
    {
        if (kPark == 1) {
            tmem_wait_st();
#pragma unroll
            for (int j = 0; j < C; ++j)
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    x[j][i] = tmem_fetch(tm + 2 * (3 * j + i));
                    w[j][i] = tmem_fetch(tm + 2 * (3 * C + 3 * j + i));
                }
            tmem_wait_ld();
        }
        if (kPark == 2) {
            tmem_wait_st();
#pragma unroll
            for (int j = 0; j < C; ++j)
#pragma unroll
                for (int i = 0; i < 3; ++i) v[j][i] = tmem_fetch(TM_AT(kPkV + 3 * j + i));
            if (kLast) {  // the reporting step reads positions (soft-fix spring)
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    x[1][i] = tmem_fetch(TM_AT(kPkX1 + i));
                    x[0][i] = sm[L::XF + i * R + t];
                }
            }
            tmem_wait_ld();
            // the force of the joints I own (rod two takes -F, half on each node of the element), deferred from P3
#pragma unroll
            for (int j = 0; j < C; ++j)
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const double hf = sm[L::HF + (3 * j + i) * R + t];
                    v[j][i] = fma(-th.dtmc[j], hf, v[j][i]);
                    if (j + 1 < C) v[(j + 1) % C][i] = fma(-th.dtmc[(j + 1) % C], hf, v[(j + 1) % C][i]);
                }
        }
        if (kLast) tile_cluster_wait<kCluster>();  // the reporting step gathers them early (external forces are reported)
        const double ramp_raw = time_mid * tb.act_inv_ramp;
        const double ramp = (ramp_raw < 1.0) ? ramp_raw : 1.0;
        // what the previous run's last element contributes to my first node / element (rows: see TileSmem)
        const int tprev = (kPark == 2) ? max(t - 1, 0) : t - 1, top = (th.to > 0) ? th.to - 1 : 0;  // (an idle thread of the lean layout sits in row 0)
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
            double ta[3];
            if (kPark == 2) {  // (fetching everything at the start of P4, or earlier than needed anywhere, was measured: the longer
                               // live ranges spill, 2.94e10 against 3.00e10 el-steps/s)
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    w[j][i] = tmem_fetch(TM_AT(kPkW + 3 * j + i));
                    ta[i] = tmem_fetch(TM_AT(kPkTA + 3 * j + i));
                }
            }
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
                    rd1[i] = sm[L::PR + (3 * j + i) * R + t];
                }
                // rod-one side of the joint owned by my partner: +(rd1 x k d); then both into my material frame
                tw[j][0] += fma(rd1[1], fso[2], -rd1[2] * fso[1]);
                tw[j][1] += fma(rd1[2], fso[0], -rd1[0] * fso[2]);
                tw[j][2] += fma(rd1[0], fso[1], -rd1[1] * fso[0]);
            }
#pragma unroll
            for (int i = 0; i < 3; ++i)
                text[i] = fma(Q[j][3 * i], tw[j][0], fma(Q[j][3 * i + 1], tw[j][1], Q[j][3 * i + 2] * tw[j][2]));
            if (kPark == 2) tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 3; ++i) text[i] = fma((kPark == 2) ? ta[i] : TAV(j, i), ramp, text[i]);
            if (kExtras && (th.xinfo[j] >> 8) != 0) {
                if (kLast) {  // half forces of the extra joints on this node (every other step: after the rotation below)
                    const int cnt = (th.xinfo[j] >> 16) & 0xF, first = th.xinfo[j] >> 20;
                    for (int q = 0; q < cnt; ++q) {
                        const int item = XIT[first + q];
                        const double sgn = (item & 1) ? -1.0 : 1.0;
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            const double f = sgn * JF[i * n_xj + (item >> 1)];
                            v[j][i] = fma(th.dtmc[j], f, v[j][i]);
                            r_fext[j][i] += f;
                        }
                    }
                }
                // a tip-to-base joint overwrites the downstream base element's director and omega during synchronize
                // (after the surface joints above have used the old director); sources are start-of-step values
                const int ow = ((th.xinfo[j] >> 8) & 0xFF) - 1;
                if (ow >= 0) {
#pragma unroll
                    for (int i = 0; i < 9; ++i) Q[j][i] = EXQ[i * n_xe + ow];
#pragma unroll
                    for (int i = 0; i < 3; ++i) w[j][i] = EXW[i * n_xe + ow];
                }
            }
            const double dil = len[j] * UNI(BR2_FC_INV_REST_LEN);
            const double strain = (th.fl[j] & TILE_F_ELEM) ? dil - 1.0 : 0.0;
            double damp[3];
            if (p.tile.round_section) {
                const double e2 = strain * UNI(BR2_FC_LN_CR + 2);
                if (!kExact) th.bad |= (e2 * e2 <= BR2_E2_MAX) ? 0 : 4;
                const double ex = M::exp(e2);
                damp[2] = UNI(BR2_FC_CR + 2) * ex;
                damp[0] = damp[1] = UNI(BR2_FC_CR) * (ex * ex);
            } else {
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const double e = strain * UNI(BR2_FC_LN_CR + i);
                    if (!kExact) th.bad |= (e * e <= BR2_E2_MAX) ? 0 : 4;
                    damp[i] = UNI(BR2_FC_CR + i) * M::exp(e);
                }
            }
            double wnew[3];
#pragma unroll
            for (int i = 0; i < 3; ++i)
                wnew[i] = fma(KD(8 + i, dt * UNI(BR2_FC_JINV + i)) * dil, tint[j][i] + text[i], w[j][i]) * damp[i];
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
            for (int i = 0; i < 3; ++i) w[j][i] = (th.fl[j] & TILE_F_ZERO_W) ? 0.0 : wnew[i];
        }
        // node rates are complete only now (the element after a node lives in the same thread, later in the loop):
        // constraints, then the trailing kinematic half step of this step fused with the leading half step of the
        // next one (unless this is the last step of the launch)
        const double hmult = kEnd ? hdt : dt;
#pragma unroll
        for (int j = 0; j < C; ++j) th.bad |= M::rotate(Q[j], w[j], hdt, hmult) ? 1 : 0;
        if (kExtras && !kLast) {
            // half forces of the tip-to-base joints on my nodes -- last, so that a cluster's other CTAs had the whole of
            // P4 to deliver them
            tile_cluster_wait<kCluster>();
#pragma unroll
            for (int j = 0; j < C; ++j) {
                const int cnt = (th.xinfo[j] >> 16) & 0xF, first = th.xinfo[j] >> 20;
                if (cnt) {
                    // the first three items (a joined element's node has three: one per rod of the other segment) without a
                    // branch between them, so that their loads are in flight together; same order of summation as the loop
                    int item[3];
                    double f[3][3];
#pragma unroll
                    for (int q = 0; q < 3; ++q) item[q] = XIT[first + min(q, cnt - 1)];
#pragma unroll
                    for (int q = 0; q < 3; ++q)
#pragma unroll
                        for (int i = 0; i < 3; ++i) f[q][i] = JF[i * n_xj + (item[q] >> 1)];
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        const double sd = (item[q] & 1) ? -th.dtmc[j] : th.dtmc[j];
                        const double s = (q < cnt) ? sd : 0.0;
#pragma unroll
                        for (int i = 0; i < 3; ++i) v[j][i] = fma(s, f[q][i], v[j][i]);
                    }
                    for (int q = 3; q < cnt; ++q) {
                        const int it = XIT[first + q];
                        const double sgn = (it & 1) ? -1.0 : 1.0;
#pragma unroll
                        for (int i = 0; i < 3; ++i) v[j][i] = fma(th.dtmc[j] * sgn, JF[i * n_xj + (it >> 1)], v[j][i]);
                    }
                }
            }
        }
        if (kPark == 2) {
            // the rotated director of slot 1 waits in tensor memory until P2 has finished with slot 0; the positions come
            // back: x[1] from tensor memory, x[0] from its published row (nobody has overwritten it: P1 of the next step is mine)
#pragma unroll
            for (int i = 0; i < 9; ++i) tmem_park(TM_AT(kPkQ1 + i), Q[1][i]);
            if (!kLast) {
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    x[1][i] = tmem_fetch(TM_AT(kPkX1 + i));
                    x[0][i] = sm[L::XF + i * R + t];
                }
                tmem_wait_ld();
            }
        }
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

template <int C, int NT, int kMinBlocks, bool kExtras, bool kCluster, bool kExact, int kPark = 0>
__global__ void __launch_bounds__(NT, kMinBlocks) br2_step_tile_kernel(const __grid_constant__ StepParams p) {
    static_assert(!kPark || (NT <= 128 && !kExtras), "register parking: one lane per thread, no divergent joint code");
    using L = TileSmem<C, NT, kPark == 2>;
    constexpr int R = L::R;
    extern __shared__ double sm[];
    __shared__ long long s_item;
    __shared__ int s_chunk;
    __shared__ int s_why;

    const DevTables &tb = p.t;
    const TilePlan &tp = p.tile;
    const int t = threadIdx.x;
    // kCluster: the CTAs of a thread-block cluster integrate one instance together, one group of ring-connected
    // rods (a segment) each; they meet at the step's barriers and exchange the tip-to-base joints' data through
    // distributed shared memory.  The cluster, not the CTA, pulls work items.
    const int rank = kCluster ? (int)cooperative_groups::this_cluster().block_rank() : 0;
    const long long total_items = p.batch * (long long)p.n_chunks;
    // exact-path pass: nothing to do unless the optimistic pass just before it flagged an instance
    if (kExact && *(volatile unsigned long long *)p.fail_counter == 0ULL) return;
    __shared__ uint32_t s_tmem;
    uint32_t tm = 0;
    if (kPark) {  // this CTA's columns of tensor memory; my warp's lane quadrant
        if (t < 32) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"l"((uint64_t)__cvta_generic_to_shared(&s_tmem)), "n"(kTileParkCols));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;");
        tm = s_tmem + ((uint32_t)(32 * ((t >> 5) & 3)) << 16);
    }

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
        th.t = t + 1;  // (kPark == 2: an idle thread's own row is row 0, set below)
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
            if (kPark == 2 && !active) th.t = th.tn = 0;
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
                th.dtmc[j] = dt * p.uni[BR2_FC_CT] / tb.slot_c[(long long)BR2_SC_MASS * S + sc];
                double ta[3] = {0.0, 0.0, 0.0};
                if (valid) {
                    for (int r = fi[BR2_FAST_ACT_BEGIN]; r < fi[BR2_FAST_ACT_END]; ++r) {
                        const double P = p.pressures[(long long)tb.act_i[r] * p.batch + b];
#pragma unroll
                        for (int i = 0; i < 3; ++i) ta[i] = fma(P, tb.act_d[r * 3 + i], ta[i]);
                    }
                }
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    if (kPark == 2)
                        tmem_park(TM_AT(kPkTA + 3 * j + i), ta[i]);
                    else
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
            th.xjoint = (kExtras && tp.n_xj > 0 && active) ? tp.xi[(rank * tp.xi_stride + t) * (C + 1) + C] : -1;
            if (kExtras) {  // this CTA's copy of the extra joints' tables
                double *const XJD = sm + L::EX + 25 * tp.n_xe + 3 * tp.n_xj;
                int *const XJI = (int *)(XJD + 8 * tp.n_xj), *const XIT = XJI + 12 * tp.n_xj;
                for (int i = t; i < 8 * tp.n_xj; i += NT) XJD[i] = tp.xj_d[i];
                for (int i = t; i < 12 * tp.n_xj; i += NT) XJI[i] = tp.xj_i[i];
                for (int i = t; i < tp.n_xitems; i += NT) XIT[i] = tp.xitems[i];
            }
            th.rod = rod;  // indexes the per-rod joint descriptors in the constant bank (TilePlan::jd)
        }
        // zero the exchange arrays (the zero row stays zero for the whole chunk)
        for (int i = t; i < L::ZERO_END; i += NT) sm[i] = 0.0;
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
        if (kPark == 2) {
#pragma unroll
            for (int i = 0; i < 9; ++i) tmem_park(TM_AT(kPkQ1 + i), th.Q[1][i]);
        }
        for (int step = 0; step < n_steps - 1; ++step) {
            tile_step<C, NT, false, false, kExtras, kCluster, kExact, kPark>(p, sm, th, time + hdt, b, rank, step & 1, tm);
            time = (time + hdt) + hdt;
        }
        if (final_chunk)
            tile_step<C, NT, true, true, kExtras, kCluster, kExact, kPark>(p, sm, th, time + hdt, b, rank, (n_steps - 1) & 1, tm);
        else
            tile_step<C, NT, true, false, kExtras, kCluster, kExact, kPark>(p, sm, th, time + hdt, b, rank, (n_steps - 1) & 1, tm);

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
    if (kPark) {
        __syncthreads();
        if (t < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "n"(kTileParkCols));
    }
}

}  // namespace br2
