/*
 * TEST INFRASTRUCTURE -- plain-C, fp64, fused CPU restatement of one PositionVerlet step of a
 * BR2 assembly.  It exists so that parity tests can check thousands of steps of many instances in
 * seconds.  It is NOT product code: only tests/, __graft_entry__.smoke() and bench.py's CPU legs
 * may load it.
 *
 * It follows, term for term and in the same floating-point association order, the operator
 * sequence of the faithful oracle (oracle/mini_elastica + oracle/br2_port), which in turn
 * follows:
 *   - PyElastica 0.3.1.post1 PositionVerlet / CosseratRod / AnalyticalLinearDamper / GravityForces /
 *     experimental SurfaceJointSideBySide   (third-party, pinned at /root/reference/poetry.lock:1784-1785,
 *     call sites /root/reference/br2/environment.py:221-223,279-285 and free_simulator.py:293-435)
 *   - /root/reference/br2/free_custom_systems.py:49-57 (bend), :96-101 (twist), :124-185 (soft-fixed base)
 *   - /root/reference/br2/legacy_surface_connection_parallel_rod_numba.py:364-478 (tip-to-base joint)
 *   - /root/reference/br2/custom_activation.py:46-81, custom_constraint.py:35-88 (optional ops)
 * Compile with -ffp-contract=off (no FMA contraction), so results are bit-identical to the
 * numba path; tests/test_oracle_fused.py asserts exactly that.
 *
 * Parity status: pinned to the reference's in-repo operators through the golden vectors; the
 * upstream PyElastica arithmetic is "parity unpinned" (see oracle/__init__.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>

/* Arithmetic type of the state and of every operation on it: real (= double) for the oracle proper.  -DBR2O_REAL=float builds
 * the SAME code in single precision (libbr2oracle_f32.so) -- not an oracle but a measuring device: how far an fp32 path
 * would drift from the fp64 one (tests/test_precision_modes.py).  Time stays fp64 in both. */
#ifndef BR2O_REAL
#define BR2O_REAL double
#endif
typedef BR2O_REAL real;

/* Named guards of the restated upstream arithmetic (same names in oracle/mini_elastica and in
 * br2-simulator_b200/csrc/br2_math.cuh): flip here if a check against real PyElastica 0.3.1 disagrees. */
#define LENGTH_GUARD 1e-14
#define ROTATION_NORM_GUARD 1e-14
#define ARCCOS_GUARD 1e-10
#define SIN_GUARD 1e-14
#define DISTANCE_ROUND_SCALE 1e12          /* np.round_(distance_vector, 12) */
#define ZERO_DISTANCE_GUARD 1e-12
#define PENETRATION_ROUND_SCALE 1e12       /* np.round_(penetration_strain, 12) */
#define SOFTFIX_ZERO_DISTANCE 1e-8         /* free_custom_systems.py:170-173 */
#endif

#define PI 3.141592653589793

/* ---- tables (filled by oracle/fused/__init__.py) -------------------------------------------- */
enum { ROD_N = 0, ROD_WOFF = 1, ROD_COFF = 2, ROD_FLAGS = 3, ROD_NI = 4 };
enum { ROD_FLAG_DAMPER = 1 };
enum { F_TYPE = 0, F_ROD = 1, F_ACTION = 2, F_NODE = 3, F_NI = 4, F_ND = 6 };
enum { FORCE_GRAVITY = 0, FORCE_TWIST = 1, FORCE_BEND = 2, FORCE_POINT = 3 };
enum { C_TYPE = 0, C_ROD = 1, C_NI = 2, C_ND = 14 };
enum { CONS_SOFT_FIXED_BASE = 0, CONS_LAST_END_FIXED = 1 };
enum { J_TYPE = 0, J_ROD1 = 1, J_ROD2 = 2, J_IDX1 = 3, J_IDX2 = 4, J_NI = 5, J_ND = 10 };
enum { JOINT_SURFACE = 0, JOINT_TIP_TO_BASE = 1 };

typedef struct {
    int32_t n_rods, n_forcing, n_constraints, n_joints, n_actions;
    int64_t words_per_instance;
    const int64_t *rod_i;      /* [n_rods][ROD_NI] */
    const real *consts;      /* per-rod constant blocks */
    const int64_t *forcing_i;  /* [n_forcing][F_NI]  (already sorted by rod, stable) */
    const real *forcing_d;   /* [n_forcing][F_ND] */
    const int64_t *cons_i;     /* [n_constraints][C_NI] (sorted by rod, stable) */
    const real *cons_d;      /* [n_constraints][C_ND] */
    const int64_t *joint_i;    /* [n_joints][J_NI] (registration order) */
    const real *joint_d;     /* [n_joints][J_ND] */
} br2o_program;

/* workspace layout of one rod inside an instance (doubles), component-major like numpy [3, n] */
typedef struct {
    int n;
    real *x, *v, *Q, *w, *a, *alpha, *f_int, *t_int, *f_ext, *t_ext, *len, *dil, *rad;
    const real *rest_len, *rest_vlen, *volume, *mass, *S, *B, *J, *Jinv, *rest_sigma, *rest_kappa, *c_r;
    real c_t;
    int flags;
} rodv;

int64_t br2o_rod_words(int64_t n) {
    return 3 * (n + 1) * 5 + 9 * n + 3 * n * 4 + 3 * n;
}

int64_t br2o_rod_const_words(int64_t n) {
    return n + (n - 1) + n + (n + 1) + 3 * n + 3 * (n - 1) + 3 * n + 3 * n + 3 * n + 3 * (n - 1) + 3 * n + 1;
}

static void bind_rod(const br2o_program *p, int r, real *W, rodv *o) {
    const int64_t *ri = p->rod_i + (int64_t)r * ROD_NI;
    int n = (int)ri[ROD_N];
    real *b = W + ri[ROD_WOFF];
    const real *c = p->consts + ri[ROD_COFF];
    o->n = n;
    o->flags = (int)ri[ROD_FLAGS];
    o->x = b;           b += 3 * (n + 1);
    o->v = b;           b += 3 * (n + 1);
    o->Q = b;           b += 9 * n;
    o->w = b;           b += 3 * n;
    o->a = b;           b += 3 * (n + 1);
    o->alpha = b;       b += 3 * n;
    o->f_int = b;       b += 3 * (n + 1);
    o->t_int = b;       b += 3 * n;
    o->f_ext = b;       b += 3 * (n + 1);
    o->t_ext = b;       b += 3 * n;
    o->len = b;         b += n;
    o->dil = b;         b += n;
    o->rad = b;
    o->rest_len = c;    c += n;
    o->rest_vlen = c;   c += n - 1;
    o->volume = c;      c += n;
    o->mass = c;        c += n + 1;
    o->S = c;           c += 3 * n;
    o->B = c;           c += 3 * (n - 1);
    o->J = c;           c += 3 * n;
    o->Jinv = c;        c += 3 * n;
    o->rest_sigma = c;  c += 3 * n;
    o->rest_kappa = c;  c += 3 * (n - 1);
    o->c_r = c;         c += 3 * n;
    o->c_t = c[0];
}

#define X(i, k) r->x[(i) * (n + 1) + (k)]
#define V(i, k) r->v[(i) * (n + 1) + (k)]
#define QQ(i, j, k) r->Q[((i) * 3 + (j)) * n + (k)]
#define WW(i, k) r->w[(i) * n + (k)]

/* x += p v ; Q <- R(p w) Q                         (Appendix A.1 step 1/8, A.4) */
static void kinematic_half_step(rodv *r, real prefac) {
    int n = r->n;
    for (int i = 0; i < 3; ++i)
        for (int k = 0; k <= n; ++k) X(i, k) += prefac * V(i, k);
    for (int k = 0; k < n; ++k) {
        real v0 = prefac * WW(0, k), v1 = prefac * WW(1, k), v2 = prefac * WW(2, k);
        real theta = sqrt(v0 * v0 + v1 * v1 + v2 * v2);
        v0 /= theta + ROTATION_NORM_GUARD;
        v1 /= theta + ROTATION_NORM_GUARD;
        v2 /= theta + ROTATION_NORM_GUARD;
        theta *= 1.0;
        real s = sin(theta), q = 1.0 - cos(theta);
        real R[3][3];
        R[0][0] = 1.0 - q * (v1 * v1 + v2 * v2);
        R[1][1] = 1.0 - q * (v0 * v0 + v2 * v2);
        R[2][2] = 1.0 - q * (v0 * v0 + v1 * v1);
        R[0][1] = s * v2 + q * v0 * v1;
        R[1][0] = -s * v2 + q * v0 * v1;
        R[0][2] = -s * v1 + q * v0 * v2;
        R[2][0] = s * v1 + q * v0 * v2;
        R[1][2] = s * v0 + q * v1 * v2;
        R[2][1] = -s * v0 + q * v1 * v2;
        real out[3][3];
        for (int i = 0; i < 3; ++i)
            for (int m = 0; m < 3; ++m) {
                real acc = 0.0;
                for (int j = 0; j < 3; ++j) acc += R[i][j] * QQ(j, m, k);
                out[i][m] = acc;
            }
        for (int i = 0; i < 3; ++i)
            for (int m = 0; m < 3; ++m) QQ(i, m, k) = out[i][m];
    }
}

/* Appendix A.3.  scratch: at least 22 * n doubles. */
static void internal_forces_and_torques(rodv *r, real *scratch) {
    int n = r->n;
    real *tan_ = scratch;            /* [3][n] */
    real *sigma = tan_ + 3 * n;      /* [3][n]  (Q t, then sigma) */
    real *stress = sigma + 3 * n;    /* [3][n] */
    real *cis = stress + 3 * n;      /* [3][n]  Q^T n / e */
    real *vdil = cis + 3 * n;        /* [n-1] */
    real *kappa = vdil + n;          /* [3][n-1] */
    real *couple = kappa + 3 * n;    /* [3][n-1] */
    real *t2d = couple + 3 * n;      /* [3][n-1] tau / E^3 */
    real *t3d = t2d + 3 * n;         /* [3][n-1] (kappa x tau) D / E^3 */
    real *qt = t3d + 3 * n;          /* [3][n] Q t */
    int nv = n - 1;

    /* geometry, dilatations */
    for (int k = 0; k < n; ++k) {
        real d0 = X(0, k + 1) - X(0, k), d1 = X(1, k + 1) - X(1, k), d2 = X(2, k + 1) - X(2, k);
        real l = sqrt(d0 * d0 + d1 * d1 + d2 * d2) + LENGTH_GUARD;
        r->len[k] = l;
        tan_[0 * n + k] = d0 / l;
        tan_[1 * n + k] = d1 / l;
        tan_[2 * n + k] = d2 / l;
        r->rad[k] = sqrt(r->volume[k] / l / PI);
        r->dil[k] = l / r->rest_len[k];
    }
    for (int k = 0; k < nv; ++k) vdil[k] = (0.5 * (r->len[k + 1] + r->len[k])) / r->rest_vlen[k];
    /* sigma = e (Q t) - z ; n_L = S (sigma - sigma_rest) */
    for (int k = 0; k < n; ++k) {
        for (int i = 0; i < 3; ++i) {
            real acc = 0.0;
            for (int j = 0; j < 3; ++j) acc += QQ(i, j, k) * tan_[j * n + k];
            qt[i * n + k] = acc;
            sigma[i * n + k] = r->dil[k] * acc - (i == 2 ? 1.0 : 0.0);
        }
        for (int i = 0; i < 3; ++i)
            stress[i * n + k] = r->S[i * n + k] * (sigma[i * n + k] - r->rest_sigma[i * n + k]);
    }
    /* Q^T n_L / e, then zero-padded difference onto nodes */
    for (int k = 0; k < n; ++k)
        for (int i = 0; i < 3; ++i) {
            real acc = 0.0;
            for (int j = 0; j < 3; ++j) acc += QQ(j, i, k) * stress[j * n + k];
            cis[i * n + k] = acc / r->dil[k];
        }
    for (int i = 0; i < 3; ++i) {
        r->f_int[i * (n + 1) + 0] = cis[i * n + 0];
        r->f_int[i * (n + 1) + n] = -cis[i * n + n - 1];
        for (int k = 1; k < n; ++k) r->f_int[i * (n + 1) + k] = cis[i * n + k] - cis[i * n + k - 1];
    }
    /* kappa from the log map of Q_{k+1} Q_k^T */
    for (int k = 0; k < nv; ++k) {
#define A(i, j) QQ(i, j, k + 1)
#define Bq(i, j) QQ(i, j, k)
        real w0 = (A(2, 0) * Bq(1, 0) + A(2, 1) * Bq(1, 1) + A(2, 2) * Bq(1, 2)) -
                    (A(1, 0) * Bq(2, 0) + A(1, 1) * Bq(2, 1) + A(1, 2) * Bq(2, 2));
        real w1 = (A(0, 0) * Bq(2, 0) + A(0, 1) * Bq(2, 1) + A(0, 2) * Bq(2, 2)) -
                    (A(2, 0) * Bq(0, 0) + A(2, 1) * Bq(0, 1) + A(2, 2) * Bq(0, 2));
        real w2 = (A(1, 0) * Bq(0, 0) + A(1, 1) * Bq(0, 1) + A(1, 2) * Bq(0, 2)) -
                    (A(0, 0) * Bq(1, 0) + A(0, 1) * Bq(1, 1) + A(0, 2) * Bq(1, 2));
        real tr = (A(0, 0) * Bq(0, 0) + A(0, 1) * Bq(0, 1) + A(0, 2) * Bq(0, 2)) +
                    (A(1, 0) * Bq(1, 0) + A(1, 1) * Bq(1, 1) + A(1, 2) * Bq(1, 2)) +
                    (A(2, 0) * Bq(2, 0) + A(2, 1) * Bq(2, 1) + A(2, 2) * Bq(2, 2));
#undef A
#undef Bq
        real theta = acos(0.5 * tr - 0.5 - ARCCOS_GUARD);
        real scale = -0.5 * theta / sin(theta + SIN_GUARD);
        w0 *= scale;
        w1 *= scale;
        w2 *= scale;
        kappa[0 * nv + k] = w0 / r->rest_vlen[k];
        kappa[1 * nv + k] = w1 / r->rest_vlen[k];
        kappa[2 * nv + k] = w2 / r->rest_vlen[k];
        for (int i = 0; i < 3; ++i)
            couple[i * nv + k] = r->B[i * nv + k] * (kappa[i * nv + k] - r->rest_kappa[i * nv + k]);
    }
    /* bend/twist couples */
    for (int k = 0; k < nv; ++k) {
        /* numba lowers `array ** 3` to x*x*x (verified here against pow(x, 3.0)) */
        real inv3 = 1.0 / (vdil[k] * vdil[k] * vdil[k]);
        real k0 = kappa[0 * nv + k], k1 = kappa[1 * nv + k], k2 = kappa[2 * nv + k];
        real c0 = couple[0 * nv + k], c1 = couple[1 * nv + k], c2 = couple[2 * nv + k];
        t2d[0 * nv + k] = c0 * inv3;
        t2d[1 * nv + k] = c1 * inv3;
        t2d[2 * nv + k] = c2 * inv3;
        t3d[0 * nv + k] = (k1 * c2 - k2 * c1) * r->rest_vlen[k] * inv3;
        t3d[1 * nv + k] = (k2 * c0 - k0 * c2) * r->rest_vlen[k] * inv3;
        t3d[2 * nv + k] = (k0 * c1 - k1 * c0) * r->rest_vlen[k] * inv3;
    }
    for (int k = 0; k < n; ++k) {
        /* dilatation rate */
        real rv0 = X(0, k) * V(0, k) + X(1, k) * V(1, k) + X(2, k) * V(2, k);
        real rv1 = X(0, k + 1) * V(0, k + 1) + X(1, k + 1) * V(1, k + 1) + X(2, k + 1) * V(2, k + 1);
        real rp1v = X(0, k + 1) * V(0, k) + X(1, k + 1) * V(1, k) + X(2, k + 1) * V(2, k);
        real rvp1 = X(0, k) * V(0, k + 1) + X(1, k) * V(1, k + 1) + X(2, k) * V(2, k + 1);
        real edot = (rv0 + rv1 - rvp1 - rp1v) / r->len[k] / r->rest_len[k];
        real e = r->dil[k];
        real q0 = qt[0 * n + k], q1 = qt[1 * n + k], q2 = qt[2 * n + k];
        real s0 = stress[0 * n + k], s1 = stress[1 * n + k], s2 = stress[2 * n + k];
        real ssc[3] = {(q1 * s2 - q2 * s1) * r->rest_len[k], (q2 * s0 - q0 * s2) * r->rest_len[k],
                         (q0 * s1 - q1 * s0) * r->rest_len[k]};
        real jw[3];
        for (int i = 0; i < 3; ++i) jw[i] = (r->J[i * n + k] * WW(i, k)) / e;
        real lag[3] = {jw[1] * WW(2, k) - jw[2] * WW(1, k), jw[2] * WW(0, k) - jw[0] * WW(2, k),
                         jw[0] * WW(1, k) - jw[1] * WW(0, k)};
        for (int i = 0; i < 3; ++i) {
            real d2, d3;
            if (nv == 0) {
                d2 = 0.0;
                d3 = 0.0;
            } else if (k == 0) {
                d2 = t2d[i * nv + 0];
                d3 = 0.5 * t3d[i * nv + 0];
            } else if (k == n - 1) {
                d2 = -t2d[i * nv + nv - 1];
                d3 = 0.5 * t3d[i * nv + nv - 1];
            } else {
                d2 = t2d[i * nv + k] - t2d[i * nv + k - 1];
                d3 = 0.5 * (t3d[i * nv + k] + t3d[i * nv + k - 1]);
            }
            real unsteady = jw[i] * edot / e;
            r->t_int[i * n + k] = d2 + d3 + ssc[i] + lag[i] + unsteady;
        }
    }
}

static void soft_fixed_base_values(rodv *r, const real *d) {
    /* free_custom_systems.py:124-135, :151-185 ; d = {x_fix[3], Q_fix[9], k, nu} */
    int n = r->n;
    real e0 = d[0] - X(0, 0), e1 = d[1] - X(1, 0), e2 = d[2] - X(2, 0);
    real dist = sqrt(e0 * e0 + e1 * e1 + e2 * e2);
    real u0 = 0.0, u1 = 0.0, u2 = 0.0;
    if (!(dist <= SOFTFIX_ZERO_DISTANCE)) {
        u0 = e0 / dist;
        u1 = e1 / dist;
        u2 = e2 / dist;
    }
    real k = d[12], nu = d[13];
    real rel0 = -V(0, 0), rel1 = -V(1, 0), rel2 = -V(2, 0);
    real proj = rel0 * u0 + rel1 * u1 + rel2 * u2;
    r->f_ext[0 * (n + 1)] += k * e0 + (-nu * (proj * u0));
    r->f_ext[1 * (n + 1)] += k * e1 + (-nu * (proj * u1));
    r->f_ext[2 * (n + 1)] += k * e2 + (-nu * (proj * u2));
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) QQ(i, j, 0) = d[3 + i * 3 + j];
}

static void constrain_values(const br2o_program *p, rodv *rods) {
    for (int c = 0; c < p->n_constraints; ++c) {
        const int64_t *ci = p->cons_i + (int64_t)c * C_NI;
        const real *cd = p->cons_d + (int64_t)c * C_ND;
        rodv *r = &rods[ci[C_ROD]];
        int n = r->n;
        if (ci[C_TYPE] == CONS_SOFT_FIXED_BASE) {
            soft_fixed_base_values(r, cd);
        } else { /* custom_constraint.py:35-44, :70-71 */
            for (int i = 0; i < 3; ++i) X(i, n) = cd[i];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) QQ(i, j, n - 1) = cd[3 + i * 3 + j];
        }
    }
}

static void constrain_rates(const br2o_program *p, rodv *rods) {
    /* dampers first (Appendix A.5), then the constraints' rates (Appendix A.1 step 7) */
    for (int ri = 0; ri < p->n_rods; ++ri) {
        rodv *r = &rods[ri];
        int n = r->n;
        if (!(r->flags & ROD_FLAG_DAMPER)) continue;
        for (int i = 0; i < 3; ++i)
            for (int k = 0; k <= n; ++k) V(i, k) = V(i, k) * r->c_t;
        for (int i = 0; i < 3; ++i)
            for (int k = 0; k < n; ++k) WW(i, k) = WW(i, k) * pow(r->c_r[i * n + k], r->dil[k]);
    }
    for (int c = 0; c < p->n_constraints; ++c) {
        const int64_t *ci = p->cons_i + (int64_t)c * C_NI;
        rodv *r = &rods[ci[C_ROD]];
        int n = r->n;
        if (ci[C_TYPE] == CONS_SOFT_FIXED_BASE) { /* free_custom_systems.py:144-149 */
            for (int i = 0; i < 3; ++i) {
                V(i, 0) = 0.0;
                WW(i, 0) = 0.0;
                r->a[i * (n + 1)] = 0.0;
                r->alpha[i * n] = 0.0;
            }
        } else { /* custom_constraint.py:46-49, :87-88 */
            for (int i = 0; i < 3; ++i) {
                V(i, n) = 0.0;
                WW(i, n - 1) = 0.0;
            }
        }
    }
}

static void apply_forcing(const br2o_program *p, rodv *rods, const real *pressure, double time) {
    for (int f = 0; f < p->n_forcing; ++f) {
        const int64_t *fi = p->forcing_i + (int64_t)f * F_NI;
        const real *fd = p->forcing_d + (int64_t)f * F_ND;
        rodv *r = &rods[fi[F_ROD]];
        int n = r->n;
        switch (fi[F_TYPE]) {
        case FORCE_GRAVITY:
            for (int i = 0; i < 3; ++i)
                for (int k = 0; k <= n; ++k) r->f_ext[i * (n + 1) + k] += fd[i] * r->mass[k];
            break;
        case FORCE_TWIST: { /* free_custom_systems.py:96-101 ; fd = {scale, ramp_up} */
            real factor = fmin(1.0, time / fd[1]);
            real P = pressure[fi[F_ACTION]];
            real dir[3] = {0.0, 0.0, 1.0};
            for (int i = 0; i < 3; ++i) {
                real torque = P * fd[0] * dir[i] * factor;
                for (int k = 0; k < n; ++k) r->t_ext[i * n + k] += (torque * 1.0) / (real)n;
            }
            break;
        }
        case FORCE_BEND: { /* free_custom_systems.py:49-57 ; fd = {scale, ramp_up, z_angle} */
            real factor = fmin(1.0, time / fd[1]);
            real mag = pressure[fi[F_ACTION]] * fd[0] * factor;
            real u[3] = {cos(fd[2]), sin(fd[2]), 0.0};
            for (int i = 0; i < 3; ++i) r->t_ext[i * n + n - 1] += mag * u[i];
            break;
        }
        case FORCE_POINT: { /* custom_activation.py:46-81 ; fd = {F[3], ramp_up} */
            real factor = fmin(1.0, time / fd[3]);
            for (int i = 0; i < 3; ++i) r->f_ext[i * (n + 1) + fi[F_NODE]] += fd[i] * factor;
            break;
        }
        }
    }
}

static void surface_joint(rodv *r1, int i1, rodv *r2, int i2, const real *jd) {
    /* Appendix A.7 ; jd = {k, nu, k_rep, dv1[3], dv2[3], off} */
    real k = jd[0], nu = jd[1], krep = jd[2], off = jd[9];
    const real *dv1 = jd + 3, *dv2 = jd + 6;
    int n1 = r1->n, n2 = r2->n;
    real g1[3] = {0, 0, 0}, g2[3] = {0, 0, 0};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            g1[i] += r1->Q[(j * 3 + i) * n1 + i1] * dv1[j];
            g2[i] += r2->Q[(j * 3 + i) * n2 + i2] * dv2[j];
        }
    real c1[3], c2[3];
    for (int i = 0; i < 3; ++i) {
        c1[i] = 0.5 * (r1->x[i * (n1 + 1) + i1] + r1->x[i * (n1 + 1) + i1 + 1]);
        c2[i] = 0.5 * (r2->x[i * (n2 + 1) + i2] + r2->x[i * (n2 + 1) + i2 + 1]);
    }
    real o1 = 0.5 * off / sqrt(r1->dil[i1]);
    real o2 = 0.5 * off / sqrt(r2->dil[i2]);
    real rd1[3], rd2[3], d[3], Fs[3], F[3];
    for (int i = 0; i < 3; ++i) {
        rd1[i] = g1[i] * (r1->rad[i1] + o1);
        rd2[i] = g2[i] * (r2->rad[i2] + o2);
    }
    for (int i = 0; i < 3; ++i) {
        d[i] = (c2[i] + rd2[i]) - (c1[i] + rd1[i]);
        d[i] = rint(d[i] * DISTANCE_ROUND_SCALE) / DISTANCE_ROUND_SCALE;
        Fs[i] = k * d[i];
    }
    real vrel[3];
    for (int i = 0; i < 3; ++i) {
        real u1 = 0.5 * (r1->v[i * (n1 + 1) + i1] + r1->v[i * (n1 + 1) + i1 + 1]);
        real u2 = 0.5 * (r2->v[i * (n2 + 1) + i2] + r2->v[i * (n2 + 1) + i2 + 1]);
        vrel[i] = u2 - u1;
    }
    real dist = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    real nd[3] = {0, 0, 0};
    if (dist >= ZERO_DISTANCE_GUARD)
        for (int i = 0; i < 3; ++i) nd[i] = d[i] / dist;
    real proj = vrel[0] * nd[0] + vrel[1] * nd[1] + vrel[2] * nd[2];
    for (int i = 0; i < 3; ++i) F[i] = Fs[i] + (-nu * (proj * nd[i]));
    real cd[3] = {c2[0] - c1[0], c2[1] - c1[1], c2[2] - c1[2]};
    real cn = sqrt(cd[0] * cd[0] + cd[1] * cd[1] + cd[2] * cd[2]);
    real pen = cn - (r1->rad[i1] + o1 + r2->rad[i2] + o2);
    pen = rint(pen * PENETRATION_ROUND_SCALE) / PENETRATION_ROUND_SCALE; /* np.round_(penetration_strain, 12) */
    if (pen < 0) {
        real mag = -krep * pow(fabs(pen), 1.5);
        for (int i = 0; i < 3; ++i) F[i] += mag * (cd[i] / cn);
    }
    for (int i = 0; i < 3; ++i) {
        r1->f_ext[i * (n1 + 1) + i1] += 0.5 * F[i];
        r1->f_ext[i * (n1 + 1) + i1 + 1] += 0.5 * F[i];
        r2->f_ext[i * (n2 + 1) + i2] -= 0.5 * F[i];
        r2->f_ext[i * (n2 + 1) + i2 + 1] -= 0.5 * F[i];
    }
    /* torques: r x F_spring, rotated to each material frame */
    real t1[3] = {rd1[1] * Fs[2] - rd1[2] * Fs[1], rd1[2] * Fs[0] - rd1[0] * Fs[2],
                    rd1[0] * Fs[1] - rd1[1] * Fs[0]};
    real t2[3] = {rd2[1] * (-Fs[2]) - rd2[2] * (-Fs[1]), rd2[2] * (-Fs[0]) - rd2[0] * (-Fs[2]),
                    rd2[0] * (-Fs[1]) - rd2[1] * (-Fs[0])};
    for (int i = 0; i < 3; ++i) {
        real a1 = 0.0, a2 = 0.0;
        for (int j = 0; j < 3; ++j) {
            a1 += r1->Q[(i * 3 + j) * n1 + i1] * t1[j];
            a2 += r2->Q[(i * 3 + j) * n2 + i2] * t2[j];
        }
        r1->t_ext[i * n1 + i1] += a1;
        r2->t_ext[i * n2 + i2] += a2;
    }
}

static void tip_to_base_joint(rodv *r1, rodv *r2, const real *jd) {
    /* Appendix A.8 ; jd = {k, nu, -, l1[3], l2[3]} */
    real k = jd[0], nu = jd[1];
    const real *l1 = jd + 3, *l2 = jd + 6;
    int n1 = r1->n, n2 = r2->n;
    real F[3];
    real a1[3], a2[3];
    /* Q^T (l r): numba's `@` on these small arrays accumulates j = 0,1,2 */
    for (int i = 0; i < 3; ++i) {
        real s1 = 0.0, s2 = 0.0;
        for (int j = 0; j < 3; ++j) {
            s1 += r1->Q[(j * 3 + i) * n1 + n1 - 1] * (l1[j] * r1->rad[n1 - 1]);
            s2 += r2->Q[(j * 3 + i) * n2 + 0] * (l2[j] * r2->rad[0]);
        }
        a1[i] = s1;
        a2[i] = s2;
    }
    for (int i = 0; i < 3; ++i) {
        real c1 = 0.5 * (r1->x[i * (n1 + 1) + n1] + r1->x[i * (n1 + 1) + n1 - 1]);
        real c2 = 0.5 * (r2->x[i * (n2 + 1) + 0] + r2->x[i * (n2 + 1) + 1]);
        real gap = (c2 + a2[i]) - (c1 + a1[i]);
        real u1 = 0.5 * (r1->v[i * (n1 + 1) + n1] + r1->v[i * (n1 + 1) + n1 - 1]);
        real u2 = 0.5 * (r2->v[i * (n2 + 1) + 0] + r2->v[i * (n2 + 1) + 1]);
        F[i] = k * gap + (-nu * (u2 - u1));
    }
    for (int i = 0; i < 3; ++i) {
        r1->f_ext[i * (n1 + 1) + n1] += 0.5 * F[i];
        r1->f_ext[i * (n1 + 1) + n1 - 1] += 0.5 * F[i];
        r2->f_ext[i * (n2 + 1) + 0] -= 0.5 * F[i];
        r2->f_ext[i * (n2 + 1) + 1] -= 0.5 * F[i];
    }
    /* the reference's DEBUG branch: copy director / omega / alpha of rod-1's tip into rod-2's base */
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) r2->Q[(i * 3 + j) * n2 + 0] = r1->Q[(i * 3 + j) * n1 + n1 - 1];
        r2->w[i * n2 + 0] = r1->w[i * n1 + n1 - 1];
        r2->alpha[i * n2 + 0] = r1->alpha[i * n1 + n1 - 1];
    }
}

static void apply_joints(const br2o_program *p, rodv *rods) {
    for (int j = 0; j < p->n_joints; ++j) {
        const int64_t *ji = p->joint_i + (int64_t)j * J_NI;
        const real *jd = p->joint_d + (int64_t)j * J_ND;
        if (ji[J_TYPE] == JOINT_SURFACE)
            surface_joint(&rods[ji[J_ROD1]], (int)ji[J_IDX1], &rods[ji[J_ROD2]], (int)ji[J_IDX2], jd);
        else
            tip_to_base_joint(&rods[ji[J_ROD1]], &rods[ji[J_ROD2]], jd);
    }
}

static void dynamic_step(rodv *r, real dt) {
    int n = r->n;
    for (int i = 0; i < 3; ++i)
        for (int k = 0; k <= n; ++k)
            r->a[i * (n + 1) + k] = (r->f_int[i * (n + 1) + k] + r->f_ext[i * (n + 1) + k]) / r->mass[k];
    for (int i = 0; i < 3; ++i)
        for (int k = 0; k < n; ++k)
            r->alpha[i * n + k] = (r->Jinv[i * n + k] * (r->t_int[i * n + k] + r->t_ext[i * n + k])) * r->dil[k];
    for (int i = 0; i < 3; ++i)
        for (int k = 0; k <= n; ++k) V(i, k) += dt * r->a[i * (n + 1) + k];
    for (int i = 0; i < 3; ++i)
        for (int k = 0; k < n; ++k) WW(i, k) += dt * r->alpha[i * n + k];
}

/* One instance, n_steps steps.  External forces/torques are zeroed at the START of a step, so on
 * return they hold what the reference's callbacks see (they run before the end-of-step zeroing). */
static double run_instance(const br2o_program *p, real *W, const real *pressure, int64_t n_steps,
                           double time, double dt, rodv *rods, real *scratch) {
    for (int r = 0; r < p->n_rods; ++r) bind_rod(p, r, W, &rods[r]);
    double prefac = 0.5 * dt;
    for (int64_t s = 0; s < n_steps; ++s) {
        for (int r = 0; r < p->n_rods; ++r) {
            int n = rods[r].n;
            memset(rods[r].f_ext, 0, sizeof(real) * 3 * (n + 1));
            memset(rods[r].t_ext, 0, sizeof(real) * 3 * n);
        }
        for (int r = 0; r < p->n_rods; ++r) kinematic_half_step(&rods[r], prefac);
        time += prefac;
        constrain_values(p, rods);
        for (int r = 0; r < p->n_rods; ++r) internal_forces_and_torques(&rods[r], scratch);
        apply_forcing(p, rods, pressure, time);
        apply_joints(p, rods);
        for (int r = 0; r < p->n_rods; ++r) dynamic_step(&rods[r], dt);
        constrain_rates(p, rods);
        for (int r = 0; r < p->n_rods; ++r) kinematic_half_step(&rods[r], prefac);
        time += prefac;
        constrain_values(p, rods);
    }
    return time;
}

/* pressures: [n_actions][batch] ; W: [batch][words_per_instance].  Returns the final time. */
double br2o_run(const br2o_program *p, real *W, const real *pressures, int64_t batch,
                int64_t n_steps, double t0, double dt, int32_t n_threads) {
    int max_n = 0;
    for (int r = 0; r < p->n_rods; ++r)
        if (p->rod_i[(int64_t)r * ROD_NI + ROD_N] > max_n) max_n = (int)p->rod_i[(int64_t)r * ROD_NI + ROD_N];
    double t_final = t0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel
#endif
    {
        rodv *rods = (rodv *)malloc(sizeof(rodv) * (size_t)p->n_rods);
        real *scratch = (real *)malloc(sizeof(real) * 32 * (size_t)(max_n + 1));
        real *press = (real *)malloc(sizeof(real) * (size_t)(p->n_actions > 0 ? p->n_actions : 1));
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
        for (int64_t b = 0; b < batch; ++b) {
            for (int a = 0; a < p->n_actions; ++a) press[a] = pressures[(int64_t)a * batch + b];
            double t = run_instance(p, W + b * p->words_per_instance, press, n_steps, t0, dt, rods, scratch);
            if (b == 0) t_final = t;
        }
        free(rods);
        free(scratch);
        free(press);
    }
    return t_final;
}
