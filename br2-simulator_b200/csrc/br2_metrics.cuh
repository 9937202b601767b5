// br2_metrics.cuh -- per-instance terminal metrics of Environment.run, one pass over the batch state.
//
// Replaces the host-side checks of /root/reference/br2/environment.py:289-349 (numpy over concatenated copies of every
// rod array) with ONE kernel: a CTA per instance walks the instance's state block once and reduces, with warp
// shuffles and one shared-memory hop, to BR2_METRICS_N doubles per instance:
//   [0] NaN mask: bit 0 position, bit 1 velocity, bit 2 director, bit 3 omega has a NaN           (:343-348)
//   [1] max over nodes of |x|   -- what the reference's "velocity" steady-state test measures      (:292-298)
//   [2] max over nodes of |v|
//   [3..8] steady-state mode 2: nanmax over nodes / elements of |previous - current| for position, velocity,
//          acceleration, director (Frobenius norm per element), omega, alpha                        (:300-338)
//          (-inf when `previous` is NULL or every column is NaN)
// Bound by HBM: the state block is read once (twice with `previous`).
#pragma once

#include "br2_tables.cuh"

namespace br2 {

constexpr int kMetricsThreads = 128;
#define BR2_METRICS_MAX_RODS 256

struct MetricsParams {
    DevTables t;
    const double *state;
    const double *previous;  // state block before the run, or nullptr
    double *out;             // [batch][BR2_METRICS_N]
    long long batch;
};

__device__ __forceinline__ double metrics_warp_max(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}

__global__ void __launch_bounds__(kMetricsThreads) br2_metrics_kernel(const MetricsParams p) {
    __shared__ double s_max[kMetricsThreads / 32][9];
    __shared__ int s_mask[kMetricsThreads / 32];
    const DevTables &tb = p.t;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    // element counts and block offsets of the rods, and the totals (field offsets inside a rod's block: include/br2cuda.h)
    __shared__ int s_n[BR2_METRICS_MAX_RODS], s_off[BR2_METRICS_MAX_RODS], s_nodes, s_elems;
    for (int r = t; r < tb.n_rods; r += kMetricsThreads) {
        s_n[r] = tb.rod_i[r * BR2_ROD_NI + BR2_ROD_N];
        s_off[r] = tb.rod_i[r * BR2_ROD_NI + BR2_ROD_WOFF];
    }
    __syncthreads();
    if (t == 0) {
        int nodes = 0, elems = 0;
        for (int r = 0; r < tb.n_rods; ++r) nodes += s_n[r] + 1, elems += s_n[r];
        s_nodes = nodes, s_elems = elems;
    }
    __syncthreads();
    for (long long b = blockIdx.x; b < p.batch; b += gridDim.x) {
        const double *inst = p.state + b * tb.words;
        const double *prev = p.previous ? p.previous + b * tb.words : nullptr;
        const double ninf = -__longlong_as_double(0x7ff0000000000000LL);
        double mx[9];  // 0 |x|, 1 |v|, 2..7 deltas, 8 NaN seen in |x| (max must then be NaN, like numpy's)
#pragma unroll
        for (int i = 0; i < 9; ++i) mx[i] = (i == 8) ? 0.0 : ninf;
        int mask = 0;
        // Nodes and elements of ALL rods are dealt to the threads as one index space (a rod of a BR2 arm has ~40 elements: rod by
        // rod, two thirds of the CTA would idle and the loads in flight would not cover the HBM latency)
        for (int idx = t; idx < s_nodes; idx += kMetricsThreads) {  // nodes: position, velocity, acceleration
            int r = 0, k = idx;
            while (k > s_n[r]) k -= s_n[r] + 1, ++r;   // (k <= n: node k of rod r)
            const int n = s_n[r], np1 = n + 1;
            const double *rod = inst + s_off[r];
            const double *rodp = prev ? prev + s_off[r] : nullptr;
            const int oX = 0, oV = 3 * np1, oA = 6 * np1 + 12 * n;
            double sx = 0.0, sv = 0.0, dx = 0.0, dv = 0.0, da = 0.0;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const double x = rod[oX + i * np1 + k], v = rod[oV + i * np1 + k];
                sx = fma(x, x, sx);
                sv = fma(v, v, sv);
                if (rodp) {
                    const double ex = rodp[oX + i * np1 + k] - x, ev = rodp[oV + i * np1 + k] - v;
                    const double ea = rodp[oA + i * np1 + k] - rod[oA + i * np1 + k];
                    dx = fma(ex, ex, dx);
                    dv = fma(ev, ev, dv);
                    da = fma(ea, ea, da);
                }
            }
            if (sx != sx) mask |= 1, mx[8] = 1.0;
            if (sv != sv) mask |= 2;
            mx[0] = fmax(mx[0], sqrt(sx));  // fmax drops NaN operands: nanmax; the NaN case of [1] is restored below
            mx[1] = fmax(mx[1], sqrt(sv));
            if (rodp) {
                mx[2] = fmax(mx[2], sqrt(dx));
                mx[3] = fmax(mx[3], sqrt(dv));
                mx[4] = fmax(mx[4], sqrt(da));
            }
        }
        for (int idx = t; idx < s_elems; idx += kMetricsThreads) {  // elements: director, omega, alpha
            int r = 0, k = idx;
            while (k >= s_n[r]) k -= s_n[r], ++r;
            const int n = s_n[r], np1 = n + 1;
            const double *rod = inst + s_off[r];
            const double *rodp = prev ? prev + s_off[r] : nullptr;
            const int oQ = 6 * np1, oW = 6 * np1 + 9 * n, oAl = 9 * np1 + 12 * n;
            double sq = 0.0, sw = 0.0, dq = 0.0, dw = 0.0, dal = 0.0;
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                const double q = rod[oQ + i * n + k];
                sq += q;  // only its NaN-ness matters
                if (rodp) {
                    const double e = rodp[oQ + i * n + k] - q;
                    dq = fma(e, e, dq);
                }
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const double w = rod[oW + i * n + k];
                sw += w;
                if (rodp) {
                    const double ew = rodp[oW + i * n + k] - w, eal = rodp[oAl + i * n + k] - rod[oAl + i * n + k];
                    dw = fma(ew, ew, dw);
                    dal = fma(eal, eal, dal);
                }
            }
            if (sq != sq) mask |= 4;
            if (sw != sw) mask |= 8;
            if (rodp) {
                mx[5] = fmax(mx[5], sqrt(dq));
                mx[6] = fmax(mx[6], sqrt(dw));
                mx[7] = fmax(mx[7], sqrt(dal));
            }
        }
        // warp shuffles, then one hop through shared memory
#pragma unroll
        for (int i = 0; i < 9; ++i) mx[i] = metrics_warp_max(mx[i]);
        mask = (int)__reduce_or_sync(0xffffffffu, (unsigned)mask);
        __syncthreads();  // (the previous instance's results have been read)
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 9; ++i) s_max[warp][i] = mx[i];
            s_mask[warp] = mask;
        }
        __syncthreads();
        if (t == 0) {
            double *out = p.out + b * BR2_METRICS_N;
            int m = 0;
            double res[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) res[i] = s_max[0][i];
            for (int wv = 0; wv < kMetricsThreads / 32; ++wv) {
                m |= s_mask[wv];
#pragma unroll
                for (int i = 0; i < 9; ++i) res[i] = fmax(res[i], s_max[wv][i]);
            }
            out[0] = (double)m;
            out[1] = (res[8] > 0.0) ? __longlong_as_double(0x7ff8000000000000LL) : res[0];  // numpy's max propagates NaN
            out[2] = res[1];
#pragma unroll
            for (int i = 0; i < 6; ++i) out[3 + i] = res[2 + i];
            for (int i = 9; i < BR2_METRICS_N; ++i) out[i] = 0.0;
        }
    }
}

}  // namespace br2
