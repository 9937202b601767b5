// br2_fast.cuh -- the performance kernel: one fused PositionVerlet step loop for BR2-structured
// assemblies (rods with damper / gravity / fibre actuation / base constraints, ring of surface joints
// between parallel rods, tip-to-base joints between segments).
//
// Same phase structure and semantics as the generic kernel (br2_generic.cuh), but
//   * straight-line code: no table walking and (almost) no branches inside the step loop -- each thread's
//     share of the operator tables is decoded once per launch, masks replace `if`s, so the compiler can
//     interleave the many short dependent fp64 chains of a step;
//   * optimistic math: reciprocals instead of divisions, hardware-seeded rsqrt, short series for the
//     tiny-angle transcendentals (br2_math.cuh).  Every series has a validity range; a thread that leaves
//     it raises the instance's flag in `status`, the CTA then does NOT write its state back, and the
//     generic kernel -- launched right behind this one on the same stream in "flagged instances only"
//     mode -- redoes the launch for that instance with libm.  Results are therefore always exact-path
//     results; the flag costs nothing when it is not raised (never, for physically sensible motion);
//   * shared-memory exchange arrays with compile-time strides; per-element constants from the constant
//     bank when every element of the assembly shares them (kUniform), else from L1-cached global memory;
//   * the trailing kinematic half step of one step and the leading one of the next share omega and are
//     applied as ONE rotation.
#pragma once

#include "br2_math.cuh"
#include "br2_tables.cuh"

namespace br2 {

template <int kThreads, int kMinBlocks, bool kUniform>
__global__ void __launch_bounds__(kThreads, kMinBlocks) br2_step_fast_kernel(const __grid_constant__ StepParams p) {
    constexpr int T = kThreads;
    constexpr int TP = kThreads + 8;   // stride of the per-element joint accumulators; slot T is a trash row
    constexpr int TJ = BR2_FAST_MAX_EXTRA;  // rows for "extra" joints (tip-to-base, or surface joints without an owner)
    extern __shared__ double sm[];
    double *X = sm;               // [3][T] node position
    double *Vs = X + 3 * T;       // [3][T] node velocity
    double *Qs = Vs + 3 * T;      // [9][T] element director
    double *Ls = Qs + 9 * T;      // [T] length
    double *Rs = Ls + T;          // [T] radius
    double *RSE = Rs + T;         // [T] 1/sqrt(dilatation)
    double *Sk = RSE + T;         // [3][T] Q^T n_L / e
    double *Av = Sk + 3 * T;      // [3][T] tau_L / E^3
    double *Cv = Av + 3 * T;      // [3][T] (kappa x tau_L) D / E^3
    double *JD = Cv + 3 * T;      // [7][T] own joint: dv1[3], dv2[3], offset
    double *TA = JD + 7 * T;      // [3][T] actuation torque at full ramp
    double *LT = TA + 3 * T;      // [3][T] thread-private stash: shear couple + transport terms (P2 -> P4)
    double *TO = LT + 3 * T;      // [3][T] thread-private stash: torque of my joint on my element (P3 -> P4)
    double *MS = TO + 3 * T;      // [2][T] node mass and its reciprocal
    double *EFo = MS + 2 * T;     // [3][TP] -F/2 of the joint this element owns (it is rod two); row T = 0
    double *TQp = EFo + 3 * TP;   // [3][TP] torque on this element of the joint in which it is rod one; row T = trash
    int *O1 = (int *)(TQp + 3 * TP);  // [T] slot of the element that owns the joint in which this element is rod one (else T)
    // tail, only allocated (by the host) for assemblies with "extra" joints / director overwrites
    double *Ws = TQp + 3 * TP + T / 2;  // [3][T] element omega
    double *JF = Ws + 3 * T;            // [3][TJ] half force of extra joints
    double *JT = JF + 3 * TJ;           // [2][3][TJ] torques of extra joints (rod one / rod two)
    int *XI = (int *)(JT + 6 * TJ);     // [14][T] per-thread extra-joint bookkeeping

    const DevTables &tb = p.t;
    const int S = tb.n_slots;
    const int s = threadIdx.x;
    const long long b = blockIdx.x;
    if (p.status[b] != 0) return;  // already waiting for the exact-path replay (whole CTA, before any barrier)
    const bool valid = s < S;
    const int sc = valid ? s : S - 1;

    const int rod = tb.slot_rod[sc];
    const int *ri = tb.rod_i + rod * BR2_ROD_NI;
    const int n = ri[BR2_ROD_N];
    const int k = sc - ri[BR2_ROD_SLOT0];
    const bool is_elem = valid && (k < n);
    const bool is_vor = valid && (k < n - 1);
    const bool first = (k == 0);

    const int *fi = tb.fast_i + sc * BR2_FAST_NI;
    const int flags = valid ? fi[BR2_FAST_FLAGS] : 0;
    const int own_joint = valid ? fi[BR2_FAST_OWN_JOINT] : -1;
    // a thread without a joint evaluates a dummy one against a neighbour (zero descriptor, result masked / trashed)
    const bool has_joint = own_joint >= 0;
    const int s1 = has_joint ? tb.joint_i[own_joint * BR2_J_NI + BR2_J_S1] : ((s + 1 < T) ? s : s - 1);
    const int jdst = has_joint ? s1 : T;   // TQp row the rod-one side of my joint goes to (row T = trash)

    // ---- per-slot constants ---------------------------------------------------------------
    const double *FCg = tb.fast_c + sc;
#define FC(idx) (kUniform ? p.uni[(idx)] : __ldg(FCg + (long long)(idx) * S))

    // joint descriptor, actuation torque, masses, accumulators -> shared memory (written once per launch)
    int any_extra_local;
    {
        const double mass = tb.slot_c[(long long)BR2_SC_MASS * S + sc];
        MS[s] = mass;
        MS[T + s] = 1.0 / mass;
        const double *jd = tb.joint_d + (has_joint ? own_joint : 0) * BR2_J_ND;
#pragma unroll
        for (int i = 0; i < 7; ++i) JD[i * T + s] = has_joint ? jd[3 + i] : 0.0;
        double ta[3] = {0.0, 0.0, 0.0};
        if (valid) {
            for (int r = fi[BR2_FAST_ACT_BEGIN]; r < fi[BR2_FAST_ACT_END]; ++r) {
                const double P = p.pressures[(long long)tb.act_i[r] * p.batch + b];
#pragma unroll
                for (int i = 0; i < 3; ++i) ta[i] = fma(P, tb.act_d[r * 3 + i], ta[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) TA[i * T + s] = ta[i];
        for (int j = s; j < 6 * TP; j += T) EFo[j] = 0.0;  // EFo and TQp are contiguous
        O1[s] = T;
        const int extra_joint = valid ? fi[BR2_FAST_EXTRA_JOINT] : -1;
        const int ow = valid ? tb.overwrite_src[sc] : -1;
        int cnt = 0;
        if (valid)
            for (int q = 0; q < BR2_FAST_NODE_ITEMS; ++q) cnt += fi[BR2_FAST_NODE_ITEM0 + q] >= 0;
        any_extra_local = (extra_joint >= 0) | (ow >= 0) | (cnt > 0) | (valid && fi[BR2_FAST_ELEM_ITEM0] >= 0);
    }
    const bool any_extra = __syncthreads_or(any_extra_local) != 0;  // CTA-uniform; also orders the O1 initialisation
    if (has_joint) O1[s1] = s;
    if (any_extra) {  // the tail of the shared-memory layout exists only then (host sizes it)
        XI[0 * T + s] = valid ? fi[BR2_FAST_EXTRA_JOINT] : -1;
        XI[1 * T + s] = fi[BR2_FAST_EXTRA_ID];
        XI[2 * T + s] = valid ? tb.overwrite_src[sc] : -1;
        int cnt = 0;
#pragma unroll
        for (int q = 0; q < BR2_FAST_NODE_ITEMS; ++q) {
            const int item = valid ? fi[BR2_FAST_NODE_ITEM0 + q] : -1;
            XI[(4 + q) * T + s] = item;
            cnt += item >= 0;
        }
        XI[3 * T + s] = cnt;
        XI[12 * T + s] = valid ? fi[BR2_FAST_ELEM_ITEM0] : -1;
        XI[13 * T + s] = valid ? fi[BR2_FAST_ELEM_ITEM0 + 1] : -1;
        for (int j = s; j < 9 * TJ; j += T) JF[j] = 0.0;  // JF and JT are contiguous
    }
    __syncthreads();

    // ---- state in registers -------------------------------------------------------------
    // field offsets inside the rod's block (include/br2cuda.h), derived where needed from n
    double *const inst = p.state + b * tb.words + ri[BR2_ROD_WOFF];
    const int np1 = n + 1;
#define G_X (inst)
#define G_V (inst + 3 * np1)
#define G_Q (inst + 6 * np1)
#define G_W (inst + 6 * np1 + 9 * n)
#define G_A (inst + 6 * np1 + 12 * n)
#define G_ALPHA (inst + 9 * np1 + 12 * n)
#define G_FINT (inst + 9 * np1 + 15 * n)
#define G_TINT (inst + 12 * np1 + 15 * n)
#define G_FEXT (inst + 12 * np1 + 18 * n)
#define G_TEXT (inst + 15 * np1 + 18 * n)
#define G_LEN (inst + 15 * np1 + 21 * n)
#define G_DIL (inst + 15 * np1 + 22 * n)
#define G_RAD (inst + 15 * np1 + 23 * n)

    double x[3], v[3], Q[9], w[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        x[i] = valid ? G_X[i * np1 + k] : 0.0;
        v[i] = valid ? G_V[i * np1 + k] : 0.0;
        w[i] = is_elem ? G_W[i * n + k] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) Q[i] = is_elem ? G_Q[i * n + k] : ((i % 4 == 0) ? 1.0 : 0.0);
    // pinned values are constants of the launch: set once, never updated
    if (flags & BR2_FAST_PIN_POSITION) {
#pragma unroll
        for (int i = 0; i < 3; ++i) x[i] = FCg[(long long)(BR2_FC_PIN + i) * S];
    }
    if (flags & BR2_FAST_PIN_DIRECTOR) {
#pragma unroll
        for (int i = 0; i < 9; ++i) Q[i] = FCg[(long long)(BR2_FC_PIN + 3 + i) * S];
    }
    const double dt = p.dt, hdt = 0.5 * p.dt;
    // masks instead of branches: a pinned node does not translate, a pinned / absent element does not rotate,
    // constrained rates are multiplied by zero
    const double hx = (flags & BR2_FAST_PIN_POSITION) ? 0.0 : hdt;
    const double hq = (is_elem && !(flags & BR2_FAST_PIN_DIRECTOR)) ? hdt : 0.0;
    const double keep_v = (flags & BR2_FAST_ZERO_NODE_RATE) ? 0.0 : 1.0;
    const double keep_w = (flags & BR2_FAST_ZERO_ELEM_RATE) ? 0.0 : 1.0;
    const double elem_mask = is_elem ? 1.0 : 0.0, vor_mask = is_vor ? 1.0 : 0.0, prev_mask = first ? 0.0 : 1.0;
    const int sp = first ? s : s - 1;            // previous slot of the same rod (masked out on the first node)
    const int sn = (s + 1 < T) ? s + 1 : s;      // next slot (the block's last thread reads itself; result masked)

    double time = p.t0;  // time at the start of the current step
    const int n_steps = (int)p.n_steps;  // the host splits longer runs
    bool bad = false;    // some series left its validity range (or produced a NaN)

    // leading kinematic half step of the first step; every later one is fused with the trailing half
    // step of the step before it (same velocities, same omega)
#pragma unroll
    for (int i = 0; i < 3; ++i) x[i] = fma(hx, v[i], x[i]);
    bad |= rotate_directors(Q, hq * w[0], hq * w[1], hq * w[2], 1.0);

    for (int step = 0; step < n_steps; ++step) {
        const bool last = (step == n_steps - 1);
        // =============================== P1: publish the half-stepped state ================
        const double time_mid = time + hdt;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            X[i * T + s] = x[i];
            Vs[i * T + s] = v[i];
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) Qs[i * T + s] = Q[i];
        if (any_extra) {
#pragma unroll
            for (int i = 0; i < 3; ++i) Ws[i * T + s] = w[i];
        }
        __syncthreads();

        // =============================== P2: element quantities (all threads; masked) =======
        {
        double d[3], dv[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            d[i] = X[i * T + sn] - x[i];
            dv[i] = Vs[i * T + sn] - v[i];
        }
        const double d2 = fma(d[0], d[0], fma(d[1], d[1], d[2] * d[2])) + (1.0 - elem_mask);  // keeps non-elements finite
        const double rs = fast_rsqrt(d2);
        const double len = fma(d2, rs, BR2_EPS_LEN);
        const double inv_len = rs * fma(-BR2_EPS_LEN, rs, 1.0);  // 1/(|d| + 1e-14), second order 1e-23
        const double rest_len = FC(BR2_FC_REST_LEN), inv_rest_len = FC(BR2_FC_INV_REST_LEN);
        const double r2 = FC(BR2_FC_VOL_OVER_PI) * inv_len;
        const double rad = r2 * fast_rsqrt(r2);
        const double dil = len * inv_rest_len;
        const double inv_dil = rest_len * inv_len;
        const double rs_e = seed_rsqrt(dil);  // only thins the ~1e-8 m rest gap of a joint: 2^-20 relative is ample
        double qt[3], nL[3], sk[3], local_t[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            qt[i] = fma(Q[3 * i], d[0], fma(Q[3 * i + 1], d[1], Q[3 * i + 2] * d[2])) * inv_len;
            const double sigma = fma(dil, qt[i], (i == 2) ? -1.0 : 0.0);
            nL[i] = FC(BR2_FC_SHEAR + i) * sigma;
        }
        const double sk_scale = inv_dil * elem_mask;
#pragma unroll
        for (int i = 0; i < 3; ++i) sk[i] = fma(Q[i], nL[0], fma(Q[3 + i], nL[1], Q[6 + i] * nL[2])) * sk_scale;
        {
            // d(e)/dt = (x_{k+1} - x_k).(v_{k+1} - v_k) / l / l_hat  (expanded form in the reference)
            const double edot = fma(d[0], dv[0], fma(d[1], dv[1], d[2] * dv[2])) * inv_len * inv_rest_len;
            double jw[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) jw[i] = FC(BR2_FC_J + i) * w[i] * inv_dil;
            const double ue = edot * inv_dil;
            local_t[0] = fma(fma(qt[1], nL[2], -qt[2] * nL[1]), rest_len, fma(jw[1], w[2], fma(-jw[2], w[1], jw[0] * ue)));
            local_t[1] = fma(fma(qt[2], nL[0], -qt[0] * nL[2]), rest_len, fma(jw[2], w[0], fma(-jw[0], w[2], jw[1] * ue)));
            local_t[2] = fma(fma(qt[0], nL[1], -qt[1] * nL[0]), rest_len, fma(jw[0], w[1], fma(-jw[1], w[0], jw[2] * ue)));
        }
        Ls[s] = len;
        Rs[s] = rad;
        RSE[s] = rs_e;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            Sk[i * T + s] = sk[i];
            LT[i * T + s] = local_t[i];
        }
        }
        __syncthreads();

        // =============================== P3: Voronoi couples and joints =====================
        const double len = Ls[s], rad = Rs[s], rs_e = RSE[s];
        {
            double Ak[3], Ck[3];
            double Qn[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) Qn[i] = Qs[i * T + sn];
#define ROWDOT(a_, b_) fma(Qn[3 * (a_)], Q[3 * (b_)], fma(Qn[3 * (a_) + 1], Q[3 * (b_) + 1], Qn[3 * (a_) + 2] * Q[3 * (b_) + 2]))
            const double w0 = ROWDOT(2, 1) - ROWDOT(1, 2);
            const double w1 = ROWDOT(0, 2) - ROWDOT(2, 0);
            const double w2 = ROWDOT(1, 0) - ROWDOT(0, 1);
            const double tr = ROWDOT(0, 0) + ROWDOT(1, 1) + ROWDOT(2, 2);
#undef ROWDOT
            const double rest_vlen = FC(BR2_FC_REST_VLEN), inv_rest_vlen = FC(BR2_FC_INV_REST_VLEN);
            // y = 1 - cos(theta) with the reference's -1e-10 guard; a benign value off the rod
            const double y = is_vor ? fma(-0.5, tr, kMisc[7]) : 0.01;
            bad |= !(y <= 0.125 && y > 0.0);
            const double ks = logmap_scale_small(y) * inv_rest_vlen * vor_mask;
            const double kap[3] = {w0 * ks, w1 * ks, w2 * ks};
            const double tau[3] = {FC(BR2_FC_BEND) * kap[0], FC(BR2_FC_BEND + 1) * kap[1], FC(BR2_FC_BEND + 2) * kap[2]};
            const double vdil = 0.5 * (Ls[sn] + len) * inv_rest_vlen;
            const double iv = fast_rcp(vdil);
            const double inv3 = iv * iv * iv;
            const double dinv3 = rest_vlen * inv3;
#pragma unroll
            for (int i = 0; i < 3; ++i) Ak[i] = tau[i] * inv3;
            Ck[0] = fma(kap[1], tau[2], -kap[2] * tau[1]) * dinv3;
            Ck[1] = fma(kap[2], tau[0], -kap[0] * tau[2]) * dinv3;
            Ck[2] = fma(kap[0], tau[1], -kap[1] * tau[0]) * dinv3;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                Av[i * T + s] = Ak[i];
                Cv[i * T + s] = Ck[i];
            }
        }
        // own joint (Appendix A.7): this element is rod two (own data in registers), rod one's element s1 is in
        // shared memory.  Threads without a joint run it against a neighbour with a zero descriptor.
        {
            const double off = JD[6 * T + s];
            const double ra1 = fma(0.5 * off, RSE[s1], Rs[s1]);   // r1 + o1
            const double ra2 = fma(0.5 * off, rs_e, rad);         // r2 + o2
            double rd1[3], rd2[3];
            {
                const double a0 = JD[0 * T + s] * ra1, a1 = JD[1 * T + s] * ra1, a2 = JD[2 * T + s] * ra1;
                const double b0 = JD[3 * T + s] * ra2, b1 = JD[4 * T + s] * ra2, b2 = JD[5 * T + s] * ra2;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    rd1[i] = fma(Qs[i * T + s1], a0, fma(Qs[(3 + i) * T + s1], a1, Qs[(6 + i) * T + s1] * a2));  // Q1^T dv1 (r1 + o1)
                    rd2[i] = fma(Q[i], b0, fma(Q[3 + i], b1, Q[6 + i] * b2));                                  // Q2^T dv2 (r2 + o2)
                }
            }
            const double jk = FC(BR2_FC_JK);
            double cd[3], dvec[3];
            double proj = 0.0;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const double c1 = 0.5 * (X[i * T + s1] + X[i * T + s1 + 1]);
                const double c2 = 0.5 * (x[i] + X[i * T + sn]);
                cd[i] = c2 - c1;
                // (c2 + rd2) - (c1 + rd1), rounded to 12 decimals like np.round_(x, 12)
                dvec[i] = rint(((c2 + rd2[i]) - (c1 + rd1[i])) * BR2_1E12) * BR2_1EM12;
                const double vr = 0.5 * (v[i] + Vs[i * T + sn]) - 0.5 * (Vs[i * T + s1] + Vs[i * T + s1 + 1]);
                proj = fma(vr, dvec[i], proj);
            }
            const double dist2 = fma(dvec[0], dvec[0], fma(dvec[1], dvec[1], dvec[2] * dvec[2]));
            // F = k d - nu (v_rel . d_hat) d_hat = (k - nu (v_rel . d) / |d|^2) d ; no damping below |d| = 1e-12
            const double inv_d2 = (dist2 >= BR2_1EM24) ? seed_rcp(dist2) : 0.0;  // damping term ~1e-8 N: seed accuracy is ample
            const double coef = fma(-FC(BR2_FC_JNU) * proj, inv_d2, jk);
            // Hertz-like repulsion when the surfaces interpenetrate: -k_rep |pen|^1.5 along the centre line
            const double cn2 = fma(cd[0], cd[0], fma(cd[1], cd[1], cd[2] * cd[2])) + (has_joint ? 0.0 : 1.0);
            const double rcn = fast_rsqrt(cn2);
            const double ap_raw = fmax(fma(-cn2, rcn, ra1 + ra2), 0.0);   // max(-pen, 0)
            const double ap = rint(ap_raw * BR2_PEN_SCALE) * BR2_PEN_INV_SCALE;  // np.round_(penetration_strain, 12)
            const double ap3 = ap * ap * ap;
            const double mag = -FC(BR2_FC_JKREP) * (ap3 * fast_rsqrt(ap3 + BR2_TINY)) * rcn;  // ap^1.5 = ap^3 / sqrt(ap^3)
            // torques use the spring part only: t = rd x (k d); rotate into each material frame
            const double t1[3] = {jk * fma(rd1[1], dvec[2], -rd1[2] * dvec[1]), jk * fma(rd1[2], dvec[0], -rd1[0] * dvec[2]),
                                  jk * fma(rd1[0], dvec[1], -rd1[1] * dvec[0])};
            const double t2[3] = {jk * fma(rd2[2], dvec[1], -rd2[1] * dvec[2]), jk * fma(rd2[0], dvec[2], -rd2[2] * dvec[0]),
                                  jk * fma(rd2[1], dvec[0], -rd2[0] * dvec[1])};
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const double hF = 0.5 * fma(coef, dvec[i], mag * cd[i]);
                EFo[i * TP + s] = has_joint ? -hF : 0.0;   // the rod-one element picks up +hF through O1
                TQp[i * TP + jdst] = fma(Qs[(3 * i) * T + s1], t1[0], fma(Qs[(3 * i + 1) * T + s1], t1[1], Qs[(3 * i + 2) * T + s1] * t1[2]));
                TO[i * T + s] = fma(Q[3 * i], t2[0], fma(Q[3 * i + 1], t2[1], Q[3 * i + 2] * t2[2]));
            }
        }
        // one further joint per thread (tip-to-base joints, or surface joints that did not fit the
        // "own" pattern), both sides read from shared memory, descriptor from the generic tables
        const int extra_joint = any_extra ? XI[0 * T + s] : -1;
        if (extra_joint >= 0) {
            const int j = extra_joint;
            const int extra_row = XI[1 * T + s];
            const int *ji = tb.joint_i + j * BR2_J_NI;
            const double *jd = tb.joint_d + j * BR2_J_ND;
            const int type = ji[BR2_J_TYPE], e1 = ji[BR2_J_S1], e2 = ji[BR2_J_S2], q1 = ji[BR2_J_Q1], q2 = ji[BR2_J_Q2];
            const double kj = jd[0], nuj = jd[1];
            double Q1[9], Q2[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                Q1[i] = Qs[i * T + q1];
                Q2[i] = Qs[i * T + q2];
            }
            double c1[3], c2[3], cd[3], vrel[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                c1[i] = 0.5 * (X[i * T + e1] + X[i * T + e1 + 1]);
                c2[i] = 0.5 * (X[i * T + e2] + X[i * T + e2 + 1]);
                cd[i] = c2[i] - c1[i];
                vrel[i] = 0.5 * (Vs[i * T + e2] + Vs[i * T + e2 + 1]) - 0.5 * (Vs[i * T + e1] + Vs[i * T + e1 + 1]);
            }
            double F[3], T1[3] = {0, 0, 0}, T2[3] = {0, 0, 0};
            if (type == BR2_JOINT_SURFACE) {
                const double off = jd[9];
                const double ra1 = fma(0.5 * off, RSE[e1], Rs[e1]);
                const double ra2 = fma(0.5 * off, RSE[e2], Rs[e2]);
                double rd1[3], rd2[3], dvec[3], Fs[3];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const double g1 = fma(Q1[i], jd[3], fma(Q1[3 + i], jd[4], Q1[6 + i] * jd[5]));
                    const double g2 = fma(Q2[i], jd[6], fma(Q2[3 + i], jd[7], Q2[6 + i] * jd[8]));
                    rd1[i] = g1 * ra1;
                    rd2[i] = g2 * ra2;
                    const double di = (c2[i] + rd2[i]) - (c1[i] + rd1[i]);
                    dvec[i] = rint(di * BR2_1E12) * BR2_1EM12;
                    Fs[i] = kj * dvec[i];
                    F[i] = Fs[i];
                }
                const double dist2 = fma(dvec[0], dvec[0], fma(dvec[1], dvec[1], dvec[2] * dvec[2]));
                if (dist2 >= BR2_1EM24) {
                    const double proj = fma(vrel[0], dvec[0], fma(vrel[1], dvec[1], vrel[2] * dvec[2]));
                    const double coef = -nuj * proj * seed_rcp(dist2);
#pragma unroll
                    for (int i = 0; i < 3; ++i) F[i] = fma(coef, dvec[i], F[i]);
                }
                const double cn2 = fma(cd[0], cd[0], fma(cd[1], cd[1], cd[2] * cd[2]));
                const double rcn = fast_rsqrt(cn2);
                const double pen = rint(fma(cn2, rcn, -(ra1 + ra2)) * BR2_PEN_SCALE) * BR2_PEN_INV_SCALE;  // np.round_(penetration_strain, 12)
                if (pen < 0.0) {
                    const double ap = -pen;
                    const double mag = -jd[2] * ap * (ap * fast_rsqrt(ap)) * rcn;
#pragma unroll
                    for (int i = 0; i < 3; ++i) F[i] = fma(mag, cd[i], F[i]);
                }
                const double t1[3] = {fma(rd1[1], Fs[2], -rd1[2] * Fs[1]), fma(rd1[2], Fs[0], -rd1[0] * Fs[2]),
                                      fma(rd1[0], Fs[1], -rd1[1] * Fs[0])};
                const double t2[3] = {fma(rd2[2], Fs[1], -rd2[1] * Fs[2]), fma(rd2[0], Fs[2], -rd2[2] * Fs[0]),
                                      fma(rd2[1], Fs[0], -rd2[0] * Fs[1])};
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    T1[i] = fma(Q1[3 * i], t1[0], fma(Q1[3 * i + 1], t1[1], Q1[3 * i + 2] * t1[2]));
                    T2[i] = fma(Q2[3 * i], t2[0], fma(Q2[3 * i + 1], t2[1], Q2[3 * i + 2] * t2[2]));
                }
            } else {  // tip-to-base (Appendix A.8): un-rounded spring, un-projected damping, no torque
                const double r1 = Rs[e1], r2 = Rs[e2];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const double a1 = fma(Q1[i], jd[3], fma(Q1[3 + i], jd[4], Q1[6 + i] * jd[5])) * r1;
                    const double a2 = fma(Q2[i], jd[6], fma(Q2[3 + i], jd[7], Q2[6 + i] * jd[8])) * r2;
                    const double gap = (c2[i] + a2) - (c1[i] + a1);
                    F[i] = fma(kj, gap, -nuj * vrel[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                JF[i * TJ + extra_row] = 0.5 * F[i];
                JT[i * TJ + extra_row] = T1[i];
                JT[(3 + i) * TJ + extra_row] = T2[i];
            }
        }
        // tip-to-base joints overwrite the downstream base element's director / omega during
        // synchronize; sources are start-of-synchronize values, still in Qs / Ws
        const int ow_src = any_extra ? XI[2 * T + s] : -1;
        if (ow_src >= 0) {
#pragma unroll
            for (int i = 0; i < 9; ++i) Q[i] = Qs[i * T + ow_src];
#pragma unroll
            for (int i = 0; i < 3; ++i) w[i] = Ws[i * T + ow_src];
        }
        __syncthreads();

        // =============================== P4: assemble, dynamic step =========================
        {
            double f_int[3], t_int[3], f_ext[3], t_ext[3];
            const double ramp_raw = time_mid * tb.act_inv_ramp;
            const double ramp = (ramp_raw < 1.0) ? ramp_raw : 1.0;
            const double mass = MS[s], inv_mass = MS[T + s];
            const int o1s = O1[s], o1p = O1[sp];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                // own values come back from shared memory rather than staying in registers across the barriers
                f_int[i] = fma(-prev_mask, Sk[i * T + sp], Sk[i * T + s]);  // Sk is 0 on the node-only slot k = n
                const double dA = fma(-prev_mask, Av[i * T + sp], Av[i * T + s]);  // Av = Cv = 0 on k >= n-1
                const double sC = fma(prev_mask, Cv[i * T + sp], Cv[i * T + s]);
                t_int[i] = dA + fma(0.5, sC, LT[i * T + s]);
                // joints: node k collects the half forces of elements k and k-1 of its rod
                // element e carries -hF of the joint it owns and +hF of the joint owned by element O1[e]
                const double fe = EFo[i * TP + s] - EFo[i * TP + o1s];
                const double fp = EFo[i * TP + sp] - EFo[i * TP + o1p];
                f_ext[i] = fma(FC(BR2_FC_GRAVITY + i), mass, fma(prev_mask, fp, fe));
                t_ext[i] = fma(TA[i * T + s], ramp, TO[i * T + s] + TQp[i * TP + s]);
            }
            if (any_extra) {  // CTA-uniform: only assemblies with tip-to-base (or otherwise unowned) joints
                const int x_node_count = XI[3 * T + s];
                for (int q = 0; q < x_node_count; ++q) {
                    const int item = XI[(4 + q) * T + s];
                    const double sgn = (item & 1) ? -1.0 : 1.0;
                    const int j = item >> 1;
#pragma unroll
                    for (int i = 0; i < 3; ++i) f_ext[i] = fma(sgn, JF[i * TJ + j], f_ext[i]);
                }
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int item = XI[(12 + e) * T + s];
                    if (item >= 0) {
                        const int base = (item & 1) * 3 * TJ + (item >> 1);
#pragma unroll
                        for (int i = 0; i < 3; ++i) t_ext[i] += JT[base + i * TJ];
                    }
                }
            }
            // dynamic step, damper, rate constraints
            double a[3], alpha[3], damp[3];
            const double c_t = FC(BR2_FC_CT) * keep_v;
            const double len4 = Ls[s];
            const double dil4 = len4 * FC(BR2_FC_INV_REST_LEN);
            const double strain = (dil4 - 1.0) * elem_mask;
            const double e0 = strain * FC(BR2_FC_LN_CR), e1 = strain * FC(BR2_FC_LN_CR + 1), e2 = strain * FC(BR2_FC_LN_CR + 2);
            // (the tile kernel's wider series, |e| <= 0.09: with the first generation's |e| <= 0.02 the strain waves that the
            // tip-to-base springs send through a multi-segment assembly flagged EVERY instance in every launch, and the generic
            // replay redid them all -- measured on double_br2: 8 replays for 8 instances)
            bad |= !(e0 * e0 <= BR2_E2_MAX) | !(e1 * e1 <= BR2_E2_MAX) | !(e2 * e2 <= BR2_E2_MAX);
            damp[0] = FC(BR2_FC_CR) * exp_tiny(e0) * keep_w;
            damp[1] = (FC(BR2_FC_LN_CR + 1) == FC(BR2_FC_LN_CR))   // I1 == I2 (round cross-section)
                          ? damp[0]
                          : FC(BR2_FC_CR + 1) * exp_tiny(e1) * keep_w;
            damp[2] = FC(BR2_FC_CR + 2) * exp_tiny(e2) * keep_w;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                a[i] = (f_int[i] + f_ext[i]) * inv_mass;
                v[i] = fma(dt, a[i], v[i]) * c_t;
                alpha[i] = FC(BR2_FC_JINV + i) * (t_int[i] + t_ext[i]) * dil4;
                w[i] = fma(dt, alpha[i], w[i]) * damp[i];
            }
            if (last && valid) {
                // quantities the reference's callbacks report (free_simulator.py:56-72), as of this step
                const int softfix_row = fi[BR2_FAST_SOFTFIX_ROW];
                if (softfix_row >= 0) {
                    // FreeBaseEndSoftFixed's spring acts on node 0 in both constrain_values calls of the step;
                    // it never changes the motion (node 0's rates are zeroed) and is only visible here.
                    // The velocity of node 0 is zero by construction, so the damping part vanishes.
                    const double *cd = tb.cons_d + softfix_row * BR2_C_ND;
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const double spring = cd[12] * (cd[i] - x[i]);
                        f_ext[i] = (f_ext[i] + spring) + spring;
                    }
                }
                if (flags & BR2_FAST_ZERO_REPORTED) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) a[i] = alpha[i] = 0.0;
                }
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    G_A[i * np1 + k] = a[i];
                    G_FINT[i * np1 + k] = f_int[i];
                    G_FEXT[i * np1 + k] = f_ext[i];
                }
                if (is_elem) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        G_ALPHA[i * n + k] = alpha[i];
                        G_TINT[i * n + k] = t_int[i];
                        G_TEXT[i * n + k] = t_ext[i];
                    }
                    G_LEN[k] = len4;
                    G_DIL[k] = dil4;
                    G_RAD[k] = Rs[s];
                }
            }
        }
        // trailing kinematic half step of this step + (unless this is the last step of the launch)
        // the leading half step of the next one; value constraints = the pins, which never move
        const double hx2 = last ? 0.0 : hx;
#pragma unroll
        for (int i = 0; i < 3; ++i) x[i] = fma(hx2, v[i], fma(hx, v[i], x[i]));
        bad |= rotate_directors(Q, hq * w[0], hq * w[1], hq * w[2], last ? 1.0 : 2.0);
        time = time_mid + hdt;
    }

    // a CTA that left the series' validity range keeps its launch-start state and asks for the exact path
    if (__syncthreads_or(bad)) {
        if (s == 0) p.status[b] = 1;
        return;
    }
    if (valid) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            G_X[i * np1 + k] = x[i];
            G_V[i * np1 + k] = v[i];
        }
        if (is_elem) {
#pragma unroll
            for (int i = 0; i < 9; ++i) G_Q[i * n + k] = Q[i];
#pragma unroll
            for (int i = 0; i < 3; ++i) G_W[i * n + k] = w[i];
        }
    }
#undef FC
}

// evaluation of the math helpers on arrays (tests only)
__global__ void selftest_math_kernel(int which, const double *in, double *out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = in[i];
    double y = 0.0, sinc, cosc;
    switch (which) {
    case 0: y = fast_rcp(x); break;
    case 1: y = fast_rsqrt(x); break;
    case 2:
        sinc_cosc_small(x, sinc, cosc);
        y = sinc;
        break;
    case 3:
        sinc_cosc_small(x, sinc, cosc);
        y = cosc;
        break;
    case 4: y = logmap_scale_small(1.0 - x); break;
    case 5: y = exp_small(x); break;
    // second generation (tile kernel)
    case 6:
        sinc_cosc_tiny(x, sinc, cosc);
        y = sinc;
        break;
    case 7:
        sinc_cosc_tiny(x, sinc, cosc);
        y = cosc;
        break;
    case 8: y = -0.5 * logmap_ratio_tiny(1.0 - x); break;
    case 9: y = exp_tiny(x); break;
    case 10: y = sqrt_mid(x); break;
    case 12: y = fast_sqrt(x); break;
    default: {  // 11: angle of rotate_directors_tiny for omega = (0, 0, x), half step 1e-5: atan2(R10, R00)
        double Q[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        const double w[3] = {0.0, 0.0, x};
        rotate_directors_tiny(Q, w, 1e-5, 1e-5);
        y = atan2(Q[1], Q[0]);
        break;
    }
    }
    out[i] = y;
}

}  // namespace br2
