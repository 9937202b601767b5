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

// 1/x: hardware seed (MUFU.RCP64H, ~2^-20 relative) + one cubic step: err -> err^3.
__device__ __forceinline__ double fast_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    const double c = fma(e, e, e);            // e + e^2
    y = fma(y, c, y);                         // y (1 + e + e^2)
    // one more linear step for full double accuracy (seed accuracy is not architecturally guaranteed)
    const double e2 = fma(-x, y, 1.0);
    return fma(y, e2, y);
}

// 1/sqrt(x) for x > 0: hardware seed (MUFU.RSQ64H) + cubic step + linear step.
__device__ __forceinline__ double fast_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double t = x * y;
    double e = fma(-t, y, 1.0);               // 1 - x y^2
    const double c = fma(0.375, e, 0.5);      // 1/2 + 3/8 e
    t = y * e;
    y = fma(t, c, y);                         // y (1 + e/2 + 3 e^2/8)
    t = x * y;
    e = fma(-t, y, 1.0);
    return fma(0.5 * y, e, y);
}

// sinc(theta) = sin(theta)/theta and cosc(theta) = (1 - cos(theta))/theta^2 as functions of
// t = theta^2, for t <= 0.01 (|theta| <= 0.1 rad per half step, i.e. |omega| <= 1e4 rad/s at dt=2e-5).
// Truncation error < 3e-18.
__device__ __forceinline__ void sinc_cosc_small(double t, double &sinc, double &cosc) {
    double s = -1.0 / 39916800.0;
    s = fma(s, t, 1.0 / 362880.0);
    s = fma(s, t, -1.0 / 5040.0);
    s = fma(s, t, 1.0 / 120.0);
    s = fma(s, t, -1.0 / 6.0);
    sinc = fma(s, t, 1.0);
    double c = -1.0 / 479001600.0;
    c = fma(c, t, 1.0 / 3628800.0);
    c = fma(c, t, -1.0 / 40320.0);
    c = fma(c, t, 1.0 / 720.0);
    c = fma(c, t, -1.0 / 24.0);
    cosc = fma(c, t, 0.5);
}

__device__ __noinline__ void sinc_cosc_large(double t, double &sinc, double &cosc) {
    const double theta = sqrt(t);
    double s, c;
    sincos(theta, &s, &c);
    sinc = s / theta;
    cosc = (1.0 - c) / t;
}

// Row-director Rodrigues update Q <- R(u) Q with the reference's guard: the unit axis is
// u / (|u| + 1e-14), i.e. the effective rotation vector is g u with
// g = |u| / (|u| + 1e-14).  R = I + a K + b K^2-like terms with a = sinc * g, b = cosc * g^2:
//   R_ii = 1 - b (u_j^2 + u_k^2),  R_ij = +-a u_k + b u_i u_j.
__device__ __forceinline__ void rotate_directors(double (&Q)[9], double u0, double u1, double u2) {
    const double t = fma(u0, u0, fma(u1, u1, u2 * u2));
    const double tt = t + 1e-300;                 // keeps rsqrt finite for a resting element
    const double rs = fast_rsqrt(tt);
    const double theta = tt * rs;
    const double g = theta * fast_rcp(theta + 1e-14);
    double sinc, cosc;
    if (t <= 0.01) {
        sinc_cosc_small(t, sinc, cosc);
    } else {
        sinc_cosc_large(t, sinc, cosc);
    }
    const double a = sinc * g;
    const double b = cosc * g * g;
    const double a0 = a * u0, a1 = a * u1, a2 = a * u2;
    const double b0 = b * u0, b1 = b * u1, b2 = b * u2;
    const double R00 = fma(-b1, u1, fma(-b2, u2, 1.0));
    const double R11 = fma(-b0, u0, fma(-b2, u2, 1.0));
    const double R22 = fma(-b0, u0, fma(-b1, u1, 1.0));
    const double R01 = fma(b0, u1, a2), R10 = fma(b0, u1, -a2);
    const double R02 = fma(b0, u2, -a1), R20 = fma(b0, u2, a1);
    const double R12 = fma(b1, u2, a0), R21 = fma(b1, u2, -a0);
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        const double q0 = Q[m], q1 = Q[3 + m], q2 = Q[6 + m];
        Q[m] = fma(R00, q0, fma(R01, q1, R02 * q2));
        Q[3 + m] = fma(R10, q0, fma(R11, q1, R12 * q2));
        Q[6 + m] = fma(R20, q0, fma(R21, q1, R22 * q2));
    }
}

// Log-map scale of the rod curvature:
//     theta = arccos(c),  scale = -0.5 * theta / sin(theta + 1e-14),   c = 0.5 tr - 0.5 - 1e-10.
// With y = 1 - c = 2 sin^2(theta/2) and z = y/2:
//     theta / sin(theta) = P(z) / sqrt(1 - z),   P(z) = asin(sqrt z)/sqrt z = sum_n c_n z^n,
// and the +1e-14 guard enters to first order as 1 / (1 + 1e-14 cos/sin).  Valid for z <= 1/16
// (relative rotation of neighbouring elements <= 0.505 rad); truncation error < 3e-17.
__device__ __forceinline__ double logmap_scale_small(double y) {
    const double z = 0.5 * y;
    double p = 0.0073125258735988454;
    p = fma(p, z, 0.008390335809616815);
    p = fma(p, z, 0.009761609529194078);
    p = fma(p, z, 0.011551800896139705);
    p = fma(p, z, 0.01396484375);
    p = fma(p, z, 0.017352764423076924);
    p = fma(p, z, 0.022372159090909092);
    p = fma(p, z, 0.030381944444444444);
    p = fma(p, z, 0.044642857142857144);
    p = fma(p, z, 0.075);
    p = fma(p, z, 1.0 / 6.0);
    p = fma(p, z, 1.0);
    const double omz = 1.0 - z;
    const double r1 = fast_rsqrt(omz);             // 1/sqrt(1-z) = 1/cos(theta/2)
    const double rz = fast_rsqrt(z);               // 1/sin(theta/2)
    const double inv_sin = 0.5 * r1 * rz;          // 1/sin(theta)
    const double cosv = 1.0 - y;
    const double guard = fma(-1e-14 * cosv, inv_sin, 1.0);   // 1/(1 + eps cot) to first order
    return -0.5 * p * r1 * guard;
}

__device__ __noinline__ double logmap_scale_large(double c) {
    const double theta = acos(c);
    return -0.5 * theta / sin(theta + 1e-14);
}

__device__ __forceinline__ double logmap_scale(double c) {
    const double y = 1.0 - c;
    if (y <= 0.125 && y > 0.0) return logmap_scale_small(y);
    return logmap_scale_large(c);
}

// exp(d) for |d| <= 0.02 (damper: exp(e ln c) = c * exp((e - 1) ln c)); truncation error < 3e-16 * 1e-0.
__device__ __forceinline__ double exp_small(double d) {
    double p = 1.0 / 5040.0;
    p = fma(p, d, 1.0 / 720.0);
    p = fma(p, d, 1.0 / 120.0);
    p = fma(p, d, 1.0 / 24.0);
    p = fma(p, d, 1.0 / 6.0);
    p = fma(p, d, 0.5);
    p = fma(p, d, 1.0);
    return fma(p, d, 1.0);
}

__device__ __noinline__ double exp_large(double d) { return exp(d); }

__device__ __forceinline__ double exp_near_zero(double d) {
    if (fabs(d) <= 0.02) return exp_small(d);
    return exp_large(d);
}

}  // namespace br2
