// br2_generic.cuh -- table-interpreting step kernel: handles every operator combination the tables of
// include/br2cuda.h can express, with IEEE division / libm transcendentals.  The fast kernel
// (br2_fast.cuh) covers the BR2 assembly structure; this one is the fallback for everything else and
// the in-tree cross-check of the fast kernel (tests/test_gpu_parity.py runs both).
#pragma once

#include "br2_tables.cuh"

namespace br2 {

// ---------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mat_rotate(double (&Q)[9], const double (&w)[3], double prefac) {
    // Q <- R(prefac * w) Q with the row-director Rodrigues matrix (Appendix A.4)
    double v0 = prefac * w[0], v1 = prefac * w[1], v2 = prefac * w[2];
    double theta = sqrt(v0 * v0 + v1 * v1 + v2 * v2);
    double inv = 1.0 / (theta + BR2_ROTATION_NORM_GUARD);
    v0 *= inv;
    v1 *= inv;
    v2 *= inv;
    double s, c;
    sincos(theta, &s, &c);
    double q = 1.0 - c;
    double R00 = 1.0 - q * (v1 * v1 + v2 * v2);
    double R11 = 1.0 - q * (v0 * v0 + v2 * v2);
    double R22 = 1.0 - q * (v0 * v0 + v1 * v1);
    double q01 = q * v0 * v1, q02 = q * v0 * v2, q12 = q * v1 * v2;
    double R01 = s * v2 + q01, R10 = -s * v2 + q01;
    double R02 = -s * v1 + q02, R20 = s * v1 + q02;
    double R12 = s * v0 + q12, R21 = -s * v0 + q12;
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        double a = Q[0 + m], b = Q[3 + m], cc = Q[6 + m];
        Q[0 + m] = R00 * a + R01 * b + R02 * cc;
        Q[3 + m] = R10 * a + R11 * b + R12 * cc;
        Q[6 + m] = R20 * a + R21 * b + R22 * cc;
    }
}

__device__ __forceinline__ double ramp_factor(double time, double ramp_up) { return fmin(1.0, time / ramp_up); }

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <int kThreads, int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks) br2_step_generic_kernel(const __grid_constant__ StepParams p) {
    extern __shared__ double sm[];
    const DevTables &T = p.t;
    const int S = T.n_slots, NJ = T.n_joints;
    double *X = sm;             // [3][S] node position
    double *Vs = X + 3 * S;     // [3][S] node velocity
    double *Qs = Vs + 3 * S;    // [9][S] element director
    double *Ws = Qs + 9 * S;    // [3][S] element omega
    double *Ls = Ws + 3 * S;    // [S] length
    double *Rs = Ls + S;        // [S] radius
    double *Es = Rs + S;        // [S] dilatation
    double *Sk = Es + S;        // [3][S] Q^T n_L / e
    double *Av = Sk + 3 * S;    // [3][S] tau_L / E^3
    double *Cv = Av + 3 * S;    // [3][S] (kappa x tau_L) D / E^3
    double *JF = Cv + 3 * S;    // [3][NJ] half of the joint force
    double *JT1 = JF + 3 * NJ;  // [3][NJ] torque on rod one (material frame)
    double *JT2 = JT1 + 3 * NJ; // [3][NJ] torque on rod two
    double *Pr = JT2 + 3 * NJ;  // [n_actions] this instance's pressures

    const int s = threadIdx.x;
    const long long b = blockIdx.x;
    if (p.only_flagged && p.status[b] == 0) return;  // replay pass: nothing to redo for this instance
    const bool valid = s < S;
    const int sc = valid ? s : S - 1;

    const int rod = T.slot_rod[sc];
    const int *ri = T.rod_i + rod * BR2_ROD_NI;
    const int n = ri[BR2_ROD_N];
    const int k = sc - ri[BR2_ROD_SLOT0];
    const bool is_elem = valid && (k < n);
    const bool is_vor = valid && (k < n - 1);
    const int rod_flags = ri[BR2_ROD_FLAGS];
    const int f_begin = ri[BR2_ROD_F_BEGIN], f_end = ri[BR2_ROD_F_END];
    const int c_begin = ri[BR2_ROD_C_BEGIN], c_end = ri[BR2_ROD_C_END];
    const double c_t = T.rod_d[rod * BR2_ROD_ND + BR2_ROD_CT];
    const int ow_src = valid ? T.overwrite_src[sc] : -1;

    if (s < T.n_actions) Pr[s] = p.pressures[(long long)s * p.batch + b];

    // per-slot constants
    const double *C = T.slot_c + sc;
#define SCONST(idx) C[(long long)(idx) * S]
    const double mass = SCONST(BR2_SC_MASS);
    const double rest_len = SCONST(BR2_SC_REST_LEN);
    const double volume = SCONST(BR2_SC_VOLUME);
    const double rest_vlen = SCONST(BR2_SC_REST_VLEN);

    // ---- state in registers -------------------------------------------------------------
    double *inst = p.state + b * T.words + ri[BR2_ROD_WOFF];
    const int np1 = n + 1;
    double *g_x = inst, *g_v = g_x + 3 * np1, *g_Q = g_v + 3 * np1, *g_w = g_Q + 9 * n;
    double *g_a = g_w + 3 * n, *g_alpha = g_a + 3 * np1, *g_fint = g_alpha + 3 * n, *g_tint = g_fint + 3 * np1;
    double *g_fext = g_tint + 3 * n, *g_text = g_fext + 3 * np1, *g_len = g_text + 3 * n, *g_dil = g_len + n,
           *g_rad = g_dil + n;

    double x[3], v[3], Q[9], w[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        x[i] = valid ? g_x[i * np1 + k] : 0.0;
        v[i] = valid ? g_v[i * np1 + k] : 0.0;
        w[i] = is_elem ? g_w[i * n + k] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) Q[i] = is_elem ? g_Q[i * n + k] : ((i % 4 == 0) ? 1.0 : 0.0);

    const double dt = p.dt, hdt = 0.5 * p.dt;
    // exact-path replay: resume at the chunk in which the optimistic kernel left its range (status = 1 + chunk)
    const int resume_chunk = p.only_flagged ? (p.status[b] & 0xFF) - 1 : 0;
    const long long first_step = (long long)resume_chunk * p.chunk_steps;
    double time = p.only_flagged ? p.chunk_t0[resume_chunk] : p.t0;

    // values carried from P2 to P4
    double len = 1.0, rad = 1.0, dil = 1.0;
    double a[3] = {0, 0, 0}, alpha[3] = {0, 0, 0}, f_int[3] = {0, 0, 0}, t_int[3] = {0, 0, 0};
    double f_ext[3] = {0, 0, 0}, t_ext[3] = {0, 0, 0};

    __syncthreads();  // Pr visible

    for (long long step = first_step; step < p.n_steps; ++step) {
        // =============================== P1: kinematic half step ===========================
#pragma unroll
        for (int i = 0; i < 3; ++i) x[i] += hdt * v[i];
        if (is_elem) mat_rotate(Q, w, hdt);
        time += hdt;
        // constrain_values, state part (the spring force is added to f_ext in P4)
        for (int c = c_begin; c < c_end; ++c) {
            const int type = T.cons_i[c * BR2_C_NI + BR2_C_TYPE];
            const double *cd = T.cons_d + c * BR2_C_ND;
            if (type == BR2_CONS_SOFT_FIXED_BASE) {
                if (k == 0) {
#pragma unroll
                    for (int i = 0; i < 9; ++i) Q[i] = cd[3 + i];
                }
            } else {
                if (k == n) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) x[i] = cd[i];
                }
                if (k == n - 1) {
#pragma unroll
                    for (int i = 0; i < 9; ++i) Q[i] = cd[3 + i];
                }
            }
        }
        if (valid) {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                X[i * S + s] = x[i];
                Vs[i * S + s] = v[i];
                Ws[i * S + s] = w[i];
            }
#pragma unroll
            for (int i = 0; i < 9; ++i) Qs[i * S + s] = Q[i];
        }
        __syncthreads();

        // =============================== P2: element quantities =============================
        double sk[3] = {0, 0, 0};     // Q^T n_L / e
        double local_t[3] = {0, 0, 0};  // shear couple + Lagrangian transport + unsteady dilatation
        if (is_elem) {
            double xn[3], vn[3], d[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                xn[i] = X[i * S + s + 1];
                vn[i] = Vs[i * S + s + 1];
                d[i] = xn[i] - x[i];
            }
            len = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) + BR2_LENGTH_GUARD;
            const double tg[3] = {d[0] / len, d[1] / len, d[2] / len};
            rad = sqrt(volume / len / BR2_PI);
            dil = len / rest_len;
            double qt[3], nL[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                qt[i] = Q[3 * i] * tg[0] + Q[3 * i + 1] * tg[1] + Q[3 * i + 2] * tg[2];
                const double sigma = dil * qt[i] - (i == 2 ? 1.0 : 0.0);
                nL[i] = SCONST(BR2_SC_SHEAR + i) * (sigma - SCONST(BR2_SC_REST_SIGMA + i));
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) sk[i] = (Q[i] * nL[0] + Q[3 + i] * nL[1] + Q[6 + i] * nL[2]) / dil;
            const double ssc[3] = {(qt[1] * nL[2] - qt[2] * nL[1]) * rest_len, (qt[2] * nL[0] - qt[0] * nL[2]) * rest_len,
                                   (qt[0] * nL[1] - qt[1] * nL[0]) * rest_len};
            const double rv0 = x[0] * v[0] + x[1] * v[1] + x[2] * v[2];
            const double rv1 = xn[0] * vn[0] + xn[1] * vn[1] + xn[2] * vn[2];
            const double rp1v = xn[0] * v[0] + xn[1] * v[1] + xn[2] * v[2];
            const double rvp1 = x[0] * vn[0] + x[1] * vn[1] + x[2] * vn[2];
            const double edot = (rv0 + rv1 - rvp1 - rp1v) / len / rest_len;
            double jw[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) jw[i] = (SCONST(BR2_SC_J + i) * w[i]) / dil;
            const double lag[3] = {jw[1] * w[2] - jw[2] * w[1], jw[2] * w[0] - jw[0] * w[2], jw[0] * w[1] - jw[1] * w[0]};
#pragma unroll
            for (int i = 0; i < 3; ++i) local_t[i] = ssc[i] + lag[i] + jw[i] * edot / dil;
            Ls[s] = len;
            Rs[s] = rad;
            Es[s] = dil;
#pragma unroll
            for (int i = 0; i < 3; ++i) Sk[i * S + s] = sk[i];
        }
        __syncthreads();

        // =============================== P3: Voronoi couples and joints =====================
        double Ak[3] = {0, 0, 0}, Ck[3] = {0, 0, 0};
        if (is_vor) {
            double Qn[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) Qn[i] = Qs[i * S + s + 1];
            // antisymmetric part and trace of R = Q_{k+1} Q_k^T (rows of Q are directors)
#define ROWDOT(a_, b_) (Qn[3 * (a_)] * Q[3 * (b_)] + Qn[3 * (a_) + 1] * Q[3 * (b_) + 1] + Qn[3 * (a_) + 2] * Q[3 * (b_) + 2])
            double w0 = ROWDOT(2, 1) - ROWDOT(1, 2);
            double w1 = ROWDOT(0, 2) - ROWDOT(2, 0);
            double w2 = ROWDOT(1, 0) - ROWDOT(0, 1);
            const double tr = ROWDOT(0, 0) + ROWDOT(1, 1) + ROWDOT(2, 2);
#undef ROWDOT
            const double theta = acos(0.5 * tr - 0.5 - BR2_ARCCOS_GUARD);
            const double scale = -0.5 * theta / sin(theta + BR2_SIN_GUARD);
            const double kap[3] = {w0 * scale / rest_vlen, w1 * scale / rest_vlen, w2 * scale / rest_vlen};
            double tau[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) tau[i] = SCONST(BR2_SC_BEND + i) * (kap[i] - SCONST(BR2_SC_REST_KAPPA + i));
            const double vdil = (0.5 * (Ls[s + 1] + len)) / rest_vlen;
            const double inv3 = 1.0 / (vdil * vdil * vdil);
#pragma unroll
            for (int i = 0; i < 3; ++i) Ak[i] = tau[i] * inv3;
            Ck[0] = (kap[1] * tau[2] - kap[2] * tau[1]) * rest_vlen * inv3;
            Ck[1] = (kap[2] * tau[0] - kap[0] * tau[2]) * rest_vlen * inv3;
            Ck[2] = (kap[0] * tau[1] - kap[1] * tau[0]) * rest_vlen * inv3;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                Av[i * S + s] = Ak[i];
                Cv[i * S + s] = Ck[i];
            }
        }
        // joints: entry j is evaluated once, by thread j (Appendix A.7 / A.8)
        for (int j = s; j < NJ; j += kThreads) {
            const int *ji = T.joint_i + j * BR2_J_NI;
            const double *jd = T.joint_d + j * BR2_J_ND;
            const int type = ji[BR2_J_TYPE], s1 = ji[BR2_J_S1], s2 = ji[BR2_J_S2], q1 = ji[BR2_J_Q1], q2 = ji[BR2_J_Q2];
            const double kj = jd[0], nuj = jd[1];
            double Q1[9], Q2[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                Q1[i] = Qs[i * S + q1];
                Q2[i] = Qs[i * S + q2];
            }
            double c1[3], c2[3], vrel[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                c1[i] = 0.5 * (X[i * S + s1] + X[i * S + s1 + 1]);
                c2[i] = 0.5 * (X[i * S + s2] + X[i * S + s2 + 1]);
                const double u1 = 0.5 * (Vs[i * S + s1] + Vs[i * S + s1 + 1]);
                const double u2 = 0.5 * (Vs[i * S + s2] + Vs[i * S + s2 + 1]);
                vrel[i] = u2 - u1;
            }
            double F[3], T1[3] = {0, 0, 0}, T2[3] = {0, 0, 0};
            if (type == BR2_JOINT_SURFACE) {
                const double krep = jd[2], off = jd[9];
                const double o1 = 0.5 * off / sqrt(Es[s1]);
                const double o2 = 0.5 * off / sqrt(Es[s2]);
                const double r1 = Rs[s1], r2 = Rs[s2];
                double rd1[3], rd2[3], dvec[3], Fs[3];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const double g1 = Q1[i] * jd[3] + Q1[3 + i] * jd[4] + Q1[6 + i] * jd[5];
                    const double g2 = Q2[i] * jd[6] + Q2[3 + i] * jd[7] + Q2[6 + i] * jd[8];
                    rd1[i] = g1 * (r1 + o1);
                    rd2[i] = g2 * (r2 + o2);
                    const double di = (c2[i] + rd2[i]) - (c1[i] + rd1[i]);
                    dvec[i] = rint(di * BR2_DISTANCE_ROUND_SCALE) / BR2_DISTANCE_ROUND_SCALE;
                    Fs[i] = kj * dvec[i];
                }
                const double dist = sqrt(dvec[0] * dvec[0] + dvec[1] * dvec[1] + dvec[2] * dvec[2]);
                double nd[3] = {0, 0, 0};
                if (dist >= BR2_ZERO_DISTANCE_GUARD) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) nd[i] = dvec[i] / dist;
                }
                const double proj = vrel[0] * nd[0] + vrel[1] * nd[1] + vrel[2] * nd[2];
#pragma unroll
                for (int i = 0; i < 3; ++i) F[i] = Fs[i] + (-nuj * (proj * nd[i]));
                const double cd[3] = {c2[0] - c1[0], c2[1] - c1[1], c2[2] - c1[2]};
                const double cn = sqrt(cd[0] * cd[0] + cd[1] * cd[1] + cd[2] * cd[2]);
                const double pen = rint((cn - (r1 + o1 + r2 + o2)) * BR2_PENETRATION_ROUND_SCALE) / BR2_PENETRATION_ROUND_SCALE;
                if (pen < 0) {
                    const double ap = fabs(pen);
                    const double mag = -krep * (ap * sqrt(ap));
#pragma unroll
                    for (int i = 0; i < 3; ++i) F[i] += mag * (cd[i] / cn);
                }
                const double t1[3] = {rd1[1] * Fs[2] - rd1[2] * Fs[1], rd1[2] * Fs[0] - rd1[0] * Fs[2],
                                      rd1[0] * Fs[1] - rd1[1] * Fs[0]};
                const double t2[3] = {rd2[2] * Fs[1] - rd2[1] * Fs[2], rd2[0] * Fs[2] - rd2[2] * Fs[0],
                                      rd2[1] * Fs[0] - rd2[0] * Fs[1]};
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    T1[i] = Q1[3 * i] * t1[0] + Q1[3 * i + 1] * t1[1] + Q1[3 * i + 2] * t1[2];
                    T2[i] = Q2[3 * i] * t2[0] + Q2[3 * i + 1] * t2[1] + Q2[3 * i + 2] * t2[2];
                }
            } else {  // tip-to-base
                const double r1 = Rs[s1], r2 = Rs[s2];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const double a1 = Q1[i] * (jd[3] * r1) + Q1[3 + i] * (jd[4] * r1) + Q1[6 + i] * (jd[5] * r1);
                    const double a2 = Q2[i] * (jd[6] * r2) + Q2[3 + i] * (jd[7] * r2) + Q2[6 + i] * (jd[8] * r2);
                    const double gap = (c2[i] + a2) - (c1[i] + a1);
                    F[i] = kj * gap + (-nuj * vrel[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                JF[i * NJ + j] = 0.5 * F[i];
                JT1[i * NJ + j] = T1[i];
                JT2[i * NJ + j] = T2[i];
            }
        }
        // tip-to-base joints overwrite the downstream base element's director / omega during
        // synchronize (legacy...:474-478); sources are start-of-synchronize values, still in Qs/Ws.
        if (ow_src >= 0) {
#pragma unroll
            for (int i = 0; i < 9; ++i) Q[i] = Qs[i * S + ow_src];
#pragma unroll
            for (int i = 0; i < 3; ++i) w[i] = Ws[i * S + ow_src];
        }
        __syncthreads();

        // =============================== P4: assemble, dynamic step =========================
        if (valid) {
            // internal force on node k: zero-padded difference of Q^T n_L / e
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const double prev = (k > 0) ? Sk[i * S + s - 1] : 0.0;
                f_int[i] = (k == 0) ? sk[i] : ((k == n) ? -prev : sk[i] - prev);
            }
            // internal torque on element k
            if (is_elem) {
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    double d2 = 0.0, d3 = 0.0;
                    if (n > 1) {
                        const double Ap = (k > 0) ? Av[i * S + s - 1] : 0.0;
                        const double Cp = (k > 0) ? Cv[i * S + s - 1] : 0.0;
                        if (k == 0) {
                            d2 = Ak[i];
                            d3 = 0.5 * Ck[i];
                        } else if (k == n - 1) {
                            d2 = -Ap;
                            d3 = 0.5 * Cp;
                        } else {
                            d2 = Ak[i] - Ap;
                            d3 = 0.5 * (Ck[i] + Cp);
                        }
                    }
                    t_int[i] = d2 + d3 + local_t[i];
                }
            }
            // external loads, in the reference's accumulation order:
            // constrain_values -> forcing (sorted by rod) -> connections (registration order)
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                f_ext[i] = 0.0;
                t_ext[i] = 0.0;
            }
            for (int c = c_begin; c < c_end; ++c) {
                if (T.cons_i[c * BR2_C_NI + BR2_C_TYPE] == BR2_CONS_SOFT_FIXED_BASE && k == 0) {
                    const double *cd = T.cons_d + c * BR2_C_ND;
                    const double e0 = cd[0] - x[0], e1 = cd[1] - x[1], e2 = cd[2] - x[2];
                    const double dist = sqrt(e0 * e0 + e1 * e1 + e2 * e2);
                    double u0 = 0.0, u1 = 0.0, u2 = 0.0;
                    if (!(dist <= BR2_SOFTFIX_ZERO_DISTANCE)) {
                        u0 = e0 / dist;
                        u1 = e1 / dist;
                        u2 = e2 / dist;
                    }
                    const double proj = (-v[0]) * u0 + (-v[1]) * u1 + (-v[2]) * u2;
                    f_ext[0] += cd[12] * e0 + (-cd[13] * (proj * u0));
                    f_ext[1] += cd[12] * e1 + (-cd[13] * (proj * u1));
                    f_ext[2] += cd[12] * e2 + (-cd[13] * (proj * u2));
                }
            }
            for (int f = f_begin; f < f_end; ++f) {
                const int *fi = T.forcing_i + f * BR2_F_NI;
                const double *fd = T.forcing_d + f * BR2_F_ND;
                const int type = fi[BR2_F_TYPE];
                if (type == BR2_FORCE_GRAVITY) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) f_ext[i] += fd[i] * mass;
                } else if (type == BR2_FORCE_TWIST) {
                    if (is_elem) {
                        const double torque = Pr[fi[BR2_F_ACTION]] * fd[0] * ramp_factor(time, fd[1]);
                        t_ext[2] += torque / (double)n;
                    }
                } else if (type == BR2_FORCE_BEND) {
                    if (k == n - 1) {
                        const double mag = Pr[fi[BR2_F_ACTION]] * fd[0] * ramp_factor(time, fd[1]);
                        t_ext[0] += mag * fd[2];
                        t_ext[1] += mag * fd[3];
                    }
                } else {  // BR2_FORCE_POINT
                    if (k == fi[BR2_F_NODE]) {
                        const double fac = ramp_factor(time, fd[3]);
#pragma unroll
                        for (int i = 0; i < 3; ++i) f_ext[i] += fd[i] * fac;
                    }
                }
            }
            for (int it = T.node_ptr[s]; it < T.node_ptr[s + 1]; ++it) {
                const int item = T.node_items[it];
                const int j = item >> 1;
                if (item & 1) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) f_ext[i] -= JF[i * NJ + j];
                } else {
#pragma unroll
                    for (int i = 0; i < 3; ++i) f_ext[i] += JF[i * NJ + j];
                }
            }
            for (int it = T.elem_ptr[s]; it < T.elem_ptr[s + 1]; ++it) {
                const int item = T.elem_items[it];
                const double *src = (item & 1) ? JT2 : JT1;
                const int j = item >> 1;
#pragma unroll
                for (int i = 0; i < 3; ++i) t_ext[i] += src[i * NJ + j];
            }
            // dynamic step
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                a[i] = (f_int[i] + f_ext[i]) / mass;
                v[i] += dt * a[i];
            }
            if (is_elem) {
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    alpha[i] = (SCONST(BR2_SC_JINV + i) * (t_int[i] + t_ext[i])) * dil;
                    w[i] += dt * alpha[i];
                }
            }
            // constrain_rates: dampers, then constraints
            if (rod_flags & BR2_ROD_HAS_DAMPER) {
#pragma unroll
                for (int i = 0; i < 3; ++i) v[i] *= c_t;
                if (is_elem) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) w[i] *= exp(dil * SCONST(BR2_SC_LN_CR + i));
                }
            }
            for (int c = c_begin; c < c_end; ++c) {
                const int type = T.cons_i[c * BR2_C_NI + BR2_C_TYPE];
                if (type == BR2_CONS_SOFT_FIXED_BASE) {
                    if (k == 0) {
#pragma unroll
                        for (int i = 0; i < 3; ++i) v[i] = w[i] = a[i] = alpha[i] = 0.0;
                    }
                } else {
                    if (k == n) {
#pragma unroll
                        for (int i = 0; i < 3; ++i) v[i] = 0.0;
                    }
                    if (k == n - 1) {
#pragma unroll
                        for (int i = 0; i < 3; ++i) w[i] = 0.0;
                    }
                }
            }
        }
        // second kinematic half step + value constraints
#pragma unroll
        for (int i = 0; i < 3; ++i) x[i] += hdt * v[i];
        if (is_elem) mat_rotate(Q, w, hdt);
        time += hdt;
        for (int c = c_begin; c < c_end; ++c) {
            const int type = T.cons_i[c * BR2_C_NI + BR2_C_TYPE];
            const double *cd = T.cons_d + c * BR2_C_ND;
            if (type == BR2_CONS_SOFT_FIXED_BASE) {
                if (k == 0) {
#pragma unroll
                    for (int i = 0; i < 9; ++i) Q[i] = cd[3 + i];
                    if (step == p.n_steps - 1) {
                        // only observable through the reported external_forces of the last step
                        const double e0 = cd[0] - x[0], e1 = cd[1] - x[1], e2 = cd[2] - x[2];
                        const double dist = sqrt(e0 * e0 + e1 * e1 + e2 * e2);
                        double u0 = 0.0, u1 = 0.0, u2 = 0.0;
                        if (!(dist <= BR2_SOFTFIX_ZERO_DISTANCE)) {
                            u0 = e0 / dist;
                            u1 = e1 / dist;
                            u2 = e2 / dist;
                        }
                        const double proj = (-v[0]) * u0 + (-v[1]) * u1 + (-v[2]) * u2;
                        f_ext[0] += cd[12] * e0 + (-cd[13] * (proj * u0));
                        f_ext[1] += cd[12] * e1 + (-cd[13] * (proj * u1));
                        f_ext[2] += cd[12] * e2 + (-cd[13] * (proj * u2));
                    }
                }
            } else {
                if (k == n) {
#pragma unroll
                    for (int i = 0; i < 3; ++i) x[i] = cd[i];
                }
                if (k == n - 1) {
#pragma unroll
                    for (int i = 0; i < 9; ++i) Q[i] = cd[3 + i];
                }
            }
        }
    }

    // ---- state + derived back to HBM ---------------------------------------------------------
    if (valid) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            g_x[i * np1 + k] = x[i];
            g_v[i * np1 + k] = v[i];
            g_a[i * np1 + k] = a[i];
            g_fint[i * np1 + k] = f_int[i];
            g_fext[i * np1 + k] = f_ext[i];
        }
        if (is_elem) {
#pragma unroll
            for (int i = 0; i < 9; ++i) g_Q[i * n + k] = Q[i];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                g_w[i * n + k] = w[i];
                g_alpha[i * n + k] = alpha[i];
                g_tint[i * n + k] = t_int[i];
                g_text[i * n + k] = t_ext[i];
            }
            g_len[k] = len;
            g_dil[k] = dil;
            g_rad[k] = rad;
        }
    }
    if (p.only_flagged && s == 0) {
        const int why = p.status[b] >> 8;  // which series' range was left (bit 0 rotation, 1 curvature, 2 strain); 0 = not recorded
        p.status[b] = 0;
        atomicAdd(p.replays, 1ULL);
        for (int bit = 0; bit < 3; ++bit)
            if (why & (1 << bit)) atomicAdd(p.replays + 1 + bit, 1ULL);
    }
#undef SCONST
}


}  // namespace br2
