// br2_math.cuh -- fp64 building blocks of the fast step kernel.
//
// The reference's arithmetic spends 40 divisions, 8 square roots and 9
// transcendentals per rod-element-step.  IEEE-exact division / sqrt / libm calls cost 15-150 fp64
// instructions each on the GPU and dominate the step.  Everything here is accurate to <= 2 ulp
// (relative 3e-16), far inside the parity budget (1e-9 after 1e4 steps; the dynamics themselves sit
// on a ~5e-13 rounding-noise floor), and validated against libm in tests/test_gpu_math.py through the
// br2_selftest_math export.
#pragma once

#include <cuda_runtime.h>
#include <math.h>

namespace br2 {

// Literal fp64 constants live in the constant bank: a DFMA can take c[bank][offset] as an operand
// directly, whereas a 64-bit immediate costs two extra instructions to materialise in registers.
__constant__ double kSincC[6] = {-1.0 / 39916800.0, 1.0 / 362880.0, -1.0 / 5040.0, 1.0 / 120.0, -1.0 / 6.0, 1.0};
__constant__ double kCoscC[6] = {-1.0 / 479001600.0, 1.0 / 3628800.0, -1.0 / 40320.0, 1.0 / 720.0, -1.0 / 24.0, 0.5};
// theta / sin(theta) = sum_n kLogC[n] z^n with z = sin^2(theta/2):  4^n (n!)^2 / (2n+1)!
__constant__ double kLogC[16] = {
    1.0, 2.0 / 3.0, 8.0 / 15.0, 16.0 / 35.0, 128.0 / 315.0, 256.0 / 693.0, 1024.0 / 3003.0, 2048.0 / 6435.0,
    32768.0 / 109395.0, 65536.0 / 230945.0, 262144.0 / 969969.0, 524288.0 / 2028117.0, 4194304.0 / 16900975.0,
    8388608.0 / 35102025.0, 33554432.0 / 145422675.0, 67108864.0 / 300540195.0};
__constant__ double kExpC[7] = {1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0};
// Named guards of the restated upstream (PyElastica 0.3.1) arithmetic -- SURVEY.md Appendix E's confidence ledger: if a
// check against the real package disagrees with one of them, THIS is the place to flip it (oracle/mini_elastica and
// oracle/fused/br2_oracle.c carry the same names).
#define BR2_LENGTH_GUARD 1e-14             // lengths = |dx| + guard                        (rod geometry)
#define BR2_ROTATION_NORM_GUARD 1e-14      // axis = u / (|u| + guard)                       (Rodrigues update)
#define BR2_ARCCOS_GUARD 1e-10             // theta = arccos(0.5 tr - 0.5 - guard)           (curvature log map)
#define BR2_SIN_GUARD 1e-14                // scale = -0.5 theta / sin(theta + guard)
#define BR2_DISTANCE_ROUND_SCALE 1e12      // np.round_(distance_vector, 12)                 (surface joint)
#define BR2_ZERO_DISTANCE_GUARD 1e-12      // no damping direction below this |distance|
#define BR2_PENETRATION_ROUND_SCALE 1e12   // np.round_(penetration_strain, 12) before the `< 0` test
#define BR2_SOFTFIX_ZERO_DISTANCE 1e-8     // FreeBaseEndSoftFixed: unit vector 0 below (free_custom_systems.py:170-173)
__constant__ double kMisc[8] = {BR2_LENGTH_GUARD, 1e-300, BR2_DISTANCE_ROUND_SCALE, 1.0 / BR2_DISTANCE_ROUND_SCALE,
                                BR2_ZERO_DISTANCE_GUARD * BR2_ZERO_DISTANCE_GUARD, 0.375, -0.5 - BR2_ARCCOS_GUARD, 1.5 + BR2_ARCCOS_GUARD};
__constant__ double kGuard[4] = {BR2_ROTATION_NORM_GUARD, BR2_SIN_GUARD, BR2_PENETRATION_ROUND_SCALE, 1.0 / BR2_PENETRATION_ROUND_SCALE};
#define BR2_EPS_LEN kMisc[0]
#define BR2_TINY kMisc[1]
#define BR2_1E12 kMisc[2]
#define BR2_1EM12 kMisc[3]
#define BR2_1EM24 kMisc[4]
#define BR2_EPS_ROT kGuard[0]
#define BR2_EPS_SIN kGuard[1]
#define BR2_PEN_SCALE kGuard[2]
#define BR2_PEN_INV_SCALE kGuard[3]

// Raw hardware seeds (MUFU.RCP64H / MUFU.RSQ64H): ~2^-20 relative accuracy, for quantities that
// multiply something tiny (damping of a joint, thinning of a 1e-8 m gap).
__device__ __forceinline__ double seed_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}
__device__ __forceinline__ double seed_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}

// 1/x: seed + one cubic step (err -> err^3 <= 2^-60).  tests/test_gpu_math.py holds it to 4 ulp.
__device__ __forceinline__ double fast_rcp(double x) {
    const double y = seed_rcp(x);
    const double e = fma(-x, y, 1.0);
    const double c = fma(e, e, e);            // e + e^2
    return fma(y, c, y);                      // y (1 + e + e^2)
}

// 1/sqrt(x) for x > 0: seed + one cubic step.
__device__ __forceinline__ double fast_rsqrt(double x) {
    const double y = seed_rsqrt(x);
    const double t = x * y;
    const double e = fma(-t, y, 1.0);         // 1 - x y^2
    const double c = fma(kMisc[5], e, 0.5);   // 1/2 + 3/8 e
    return fma(y * e, c, y);                  // y (1 + e/2 + 3 e^2/8)
}

// sqrt(x), x > 0: the same cubic step applied to s = x y (one multiplication less than x * fast_rsqrt(x)).
__device__ __forceinline__ double fast_sqrt(double x) {
    const double y = seed_rsqrt(x);
    const double s = x * y;
    const double e = fma(-s, y, 1.0);         // 1 - x y^2
    const double c = fma(kMisc[5], e, 0.5);   // 1/2 + 3/8 e
    return fma(s * e, c, s);                  // s (1 + e/2 + 3 e^2/8)
}

// sinc(theta) = sin(theta)/theta and cosc(theta) = (1 - cos(theta))/theta^2 as functions of
// t = theta^2, for t <= 0.01 (|theta| <= 0.1 rad per step, i.e. |omega| <= 5e3 rad/s at dt=2e-5).
// Truncation error < 3e-18.
__device__ __forceinline__ void sinc_cosc_small(double t, double &sinc, double &cosc) {
    // Estrin: (c5 + c4 t) + t^2 (c3 + c2 t) + t^4 (c1 + c0 t), coefficients stored highest power first
    const double t2 = t * t, t4 = t2 * t2;
    const double s01 = fma(kSincC[4], t, kSincC[5]), s23 = fma(kSincC[2], t, kSincC[3]), s45 = fma(kSincC[0], t, kSincC[1]);
    const double c01 = fma(kCoscC[4], t, kCoscC[5]), c23 = fma(kCoscC[2], t, kCoscC[3]), c45 = fma(kCoscC[0], t, kCoscC[1]);
    sinc = fma(s45, t4, fma(s23, t2, s01));
    cosc = fma(c45, t4, fma(c23, t2, c01));
}

// Row-director Rodrigues update for `mult` consecutive kinematic half steps with the same rotation
// vector u = (dt/2) omega (mult = 1 or 2).  One half step of the reference uses
// the unit axis u / (|u| + 1e-14), i.e. rotates by g |u| with g = |u| / (|u| + 1e-14); omega does not
// change between the trailing half step of one step and the leading half step of the next, so the two
// rotations share their axis and compose to one rotation by phi = 2 g |u| (applied as ONE matrix).
//   R = I + (sin phi / |u|) K(u) + ((1 - cos phi) / |u|^2) (u u^T - |u|^2 I)
//   sin phi / |u| = sinc(phi) m,   (1 - cos phi) / |u|^2 = cosc(phi) m^2,   m = mult * g.
// Returns true when phi^2 left the series' range (|phi| > 0.1 rad in one step) or is NaN.
__device__ __forceinline__ bool rotate_directors(double (&Q)[9], double u0, double u1, double u2, double mult) {
    const double t = fma(u0, u0, fma(u1, u1, u2 * u2));
    const double tt = t + BR2_TINY;                 // keeps rsqrt finite for a resting element
    const double rs = fast_rsqrt(tt);
    const double theta = tt * rs;
    const double m = mult * theta * fast_rcp(theta + BR2_EPS_ROT);
    const double m2 = m * m;
    const double tphi = m2 * t;                   // phi^2
    double sinc, cosc;
    sinc_cosc_small(tphi, sinc, cosc);
    const double a = sinc * m;
    const double b = cosc * m2;
    const double a0 = a * u0, a1 = a * u1, a2 = a * u2;
    const double b0 = b * u0, b1 = b * u1, b2 = b * u2;
    const double R00 = fma(-b1, u1, fma(-b2, u2, 1.0));
    const double R11 = fma(-b0, u0, fma(-b2, u2, 1.0));
    const double R22 = fma(-b0, u0, fma(-b1, u1, 1.0));
    const double R01 = fma(b0, u1, a2), R10 = fma(b0, u1, -a2);
    const double R02 = fma(b0, u2, -a1), R20 = fma(b0, u2, a1);
    const double R12 = fma(b1, u2, a0), R21 = fma(b1, u2, -a0);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double q0 = Q[c], q1 = Q[3 + c], q2 = Q[6 + c];
        Q[c] = fma(R00, q0, fma(R01, q1, R02 * q2));
        Q[3 + c] = fma(R10, q0, fma(R11, q1, R12 * q2));
        Q[6 + c] = fma(R20, q0, fma(R21, q1, R22 * q2));
    }
    return !(tphi <= 0.01);
}

// Log-map scale of the rod curvature:
//     theta = arccos(c),  scale = -0.5 * theta / sin(theta + 1e-14),   c = 0.5 tr - 0.5 - 1e-10.
// With y = 1 - c = 2 sin^2(theta/2) and z = y/2:  theta / sin(theta) = sum_n kLogC[n] z^n (Estrin scheme,
// degree 15, truncation < 2e-18 for z <= 1/16, i.e. neighbouring elements rotated by <= 0.505 rad), and the
// +1e-14 guard enters to first order as the factor 1 - 1e-14 cos(theta)/sin(theta); 1/sin(theta) =
// 1/(2 sqrt(z (1 - z))) only needs the hardware seed's accuracy there (the factor differs from 1 by < 1e-9).
__device__ __forceinline__ double logmap_scale_small(double y) {
    const double z = 0.5 * y;
    const double z2 = z * z, z4 = z2 * z2, z8 = z4 * z4;
    const double p01 = fma(kLogC[1], z, kLogC[0]), p23 = fma(kLogC[3], z, kLogC[2]);
    const double p45 = fma(kLogC[5], z, kLogC[4]), p67 = fma(kLogC[7], z, kLogC[6]);
    const double p89 = fma(kLogC[9], z, kLogC[8]), pab = fma(kLogC[11], z, kLogC[10]);
    const double pcd = fma(kLogC[13], z, kLogC[12]), pef = fma(kLogC[15], z, kLogC[14]);
    const double q0 = fma(p23, z2, p01), q1 = fma(p67, z2, p45), q2 = fma(pab, z2, p89), q3 = fma(pef, z2, pcd);
    const double r0 = fma(q1, z4, q0), r1 = fma(q3, z4, q2);
    const double f = fma(r1, z8, r0);                          // theta / sin(theta)
    const double inv_sin = 0.5 * seed_rsqrt(z * (1.0 - z));    // 1 / sin(theta), ~2^-20 relative
    const double guard = fma(-BR2_EPS_SIN * (1.0 - y), inv_sin, 1.0);
    return -0.5 * f * guard;
}

// exp(d) for |d| <= 0.02 (damper: exp(e ln c) = c * exp((e - 1) ln c)); degree-7 Taylor, truncation < 1e-18.
__device__ __forceinline__ double exp_small(double d) {
    double p = kExpC[0];
#pragma unroll
    for (int i = 1; i < 7; ++i) p = fma(p, d, kExpC[i]);
    return fma(p, d, 1.0);
}

// ---------------------------------------------------------------------------------------------
// Second generation of the helpers (tile kernel, br2_tile.cuh): shorter series with tighter validity
// ranges -- still far outside anything a BR2 assembly does in one 2e-5 s step -- and guards evaluated at
// hardware-seed accuracy where they only multiply 1e-14.
// ---------------------------------------------------------------------------------------------
__constant__ double kSinc4[5] = {1.0 / 362880.0, -1.0 / 5040.0, 1.0 / 120.0, -1.0 / 6.0, 1.0};
__constant__ double kCosc4[5] = {1.0 / 3628800.0, -1.0 / 40320.0, 1.0 / 720.0, -1.0 / 24.0, 0.5};
__constant__ double kMisc2[6] = {2e-3, 0.03, 8.1e-3, 0.01, 0.25, 0.3125};
__constant__ double kExp10[10] = {1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0, 1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0};
#define BR2_TPHI_MAX kMisc2[0]
#define BR2_Y_MAX kMisc2[1]
#define BR2_E2_MAX kMisc2[2]

// sinc / cosc of theta as functions of t = theta^2 <= 2e-3 (|theta| <= 0.0447 rad in one step); degree 4,
// truncation < t^5 / 11! = 8e-22.
__device__ __forceinline__ void sinc_cosc_tiny(double t, double &sinc, double &cosc) {
    double s = kSinc4[0], c = kCosc4[0];
#pragma unroll
    for (int i = 1; i < 5; ++i) {
        s = fma(s, t, kSinc4[i]);
        c = fma(c, t, kCosc4[i]);
    }
    sinc = s;
    cosc = c;
}

// Q <- R Q for `mult` (1 or 2) consecutive kinematic half steps with rotation vector u = hdt * w each.
// Same rotation as rotate_directors(); the reference's guard g = |u| / (|u| + 1e-14) = 1 - 1e-14 / (|u| + 1e-14)
// only needs |u| to seed accuracy.  hmult = mult * hdt.  Returns true when the angle left the series' range.
__device__ __forceinline__ bool rotate_directors_tiny(double (&Q)[9], const double (&w)[3], double hdt, double hmult) {
    const double tw = fma(w[0], w[0], fma(w[1], w[1], w[2] * w[2]));
    const double tt = tw + BR2_TINY;
    // (the guard from ONE seed, 1 - (1e-14 / hdt) / |w| clamped to [0, 1], is one MUFU and one dependent step shorter and
    // was measured 3.5 % SLOWER on the tile kernel -- ptxas' schedule decides, not the instruction count)
    const double absu = hdt * (tt * seed_rsqrt(tt));
    const double g = fma(-BR2_EPS_ROT, seed_rcp(absu + BR2_EPS_ROT), 1.0);
    const double hm = hmult * g;
    const double u0 = hm * w[0], u1 = hm * w[1], u2 = hm * w[2];
    const double tphi = (hm * hm) * tw;  // phi^2
    double sinc, cosc;
    sinc_cosc_tiny(tphi, sinc, cosc);
    const double a0 = sinc * u0, a1 = sinc * u1, a2 = sinc * u2;
    const double b0 = cosc * u0, b1 = cosc * u1, b2 = cosc * u2;
    const double R00 = fma(-b1, u1, fma(-b2, u2, 1.0));
    const double R11 = fma(-b0, u0, fma(-b2, u2, 1.0));
    const double R22 = fma(-b0, u0, fma(-b1, u1, 1.0));
    const double R01 = fma(b0, u1, a2), R10 = fma(b0, u1, -a2);
    const double R02 = fma(b0, u2, -a1), R20 = fma(b0, u2, a1);
    const double R12 = fma(b1, u2, a0), R21 = fma(b1, u2, -a0);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double q0 = Q[c], q1 = Q[3 + c], q2 = Q[6 + c];
        Q[c] = fma(R00, q0, fma(R01, q1, R02 * q2));
        Q[3 + c] = fma(R10, q0, fma(R11, q1, R12 * q2));
        Q[6 + c] = fma(R20, q0, fma(R21, q1, R22 * q2));
    }
    return !(tphi <= BR2_TPHI_MAX);
}

// theta / sin(theta + 1e-14) from y = 1 - cos(theta) <= 0.03 (neighbouring elements rotated by <= 0.245 rad):
// degree-9 series in z = y/2 (two Horner chains in z^2; truncation 0.27 z^10 < 2e-19) times the guard
// 1 - 1e-14 cos/sin with 1/sin = rsqrt(y (2 - y)) from the hardware seed.
__device__ __forceinline__ double logmap_ratio_tiny(double y) {
    const double z = 0.5 * y, z2 = z * z;
    double pe = kLogC[8], po = kLogC[9];
    pe = fma(pe, z2, kLogC[6]);
    po = fma(po, z2, kLogC[7]);
    pe = fma(pe, z2, kLogC[4]);
    po = fma(po, z2, kLogC[5]);
    pe = fma(pe, z2, kLogC[2]);
    po = fma(po, z2, kLogC[3]);
    pe = fma(pe, z2, kLogC[0]);
    po = fma(po, z2, kLogC[1]);
    const double f = fma(po, z, pe);
    const double inv_sin = seed_rsqrt(fma(-y, y, y + y));
    const double guard = fma(fma(BR2_EPS_SIN, y, -BR2_EPS_SIN), inv_sin, 1.0);
    return f * guard;
}

// exp(d) for |d| <= 0.09 (BR2_E2_MAX = 0.09^2); degree 10, truncation d^11 / 11! < 8e-20.  The range is wide on
// purpose: the tip-to-base springs of a multi-segment assembly send strain waves of several per cent through the
// rods (3.5 % in double_br2_v1), and with ln c_r3 = -0.35 this admits |strain| <= 25 %.
__device__ __forceinline__ double exp_tiny(double d) {
    // even and odd part as two Horner chains in d^2 (same instruction count as one chain, 7 instead of 11 deep)
    const double d2 = d * d;
    double pe = kExp10[0], po = kExp10[1];
#pragma unroll
    for (int i = 2; i < 10; i += 2) {
        pe = fma(pe, d2, kExp10[i]);
        po = fma(po, d2, kExp10[i + 1]);
    }
    pe = fma(pe, d2, 1.0);
    return fma(po, d, pe);
}

// sqrt(x) for x >= 0 to ~2^-40 relative (seed + one Newton step); 0 for x = 0.
__device__ __forceinline__ double sqrt_mid(double x) {
    const double y = seed_rsqrt(x + BR2_TINY);
    const double s = x * y;
    return fma(0.5 * y, fma(-s, s, x), s);
}

}  // namespace br2
