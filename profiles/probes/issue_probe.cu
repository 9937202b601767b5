// Does an fp64 instruction block the SMSP's issue port for both of its pipe cycles on B200?
// 8 independent DFMA chains per thread, interleaved with K independent 32-bit IMADs per DFMA.
// If issue is not blocked, K = 1 is free (DFMA: 1 issue / 2 cycles); if it is, time grows with K from K = 0.
#include <cstdio>
#include <cuda_runtime.h>

template <int K>
__global__ void __launch_bounds__(256) probe(double *out, int *iout, int iters, double seed, int iseed) {
    double a[8];
    int c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = seed + i * 1e-3 + threadIdx.x * 1e-9;
        c[i] = iseed + i + threadIdx.x;
    }
    const double m = 0.999999, d = 1e-7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                a[i] = fma(a[i], m, d);
#pragma unroll
                for (int k = 0; k < K; ++k) c[(i + k) & 7] = c[(i + k) & 7] * 1664525 + 1013904223;
            }
        }
    }
    double s = 0;
    int t = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        s += a[i];
        t ^= c[i];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    iout[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int K>
float run(double *out, int *iout, int blocks) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0);
        probe<K><<<blocks, 256>>>(out, iout, 4000, 1.0 + r, r);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
    }
    return best;
}

int main() {
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 8;
    double *out;
    int *iout;
    cudaMalloc(&out, sizeof(double) * blocks * 256);
    cudaMalloc(&iout, sizeof(int) * blocks * 256);
    const double dfma = 64.0 * 4000 * blocks * 256.0;
    float t0 = run<0>(out, iout, blocks), t1 = run<1>(out, iout, blocks), t2 = run<2>(out, iout, blocks), t3 = run<3>(out, iout, blocks);
    printf("K=0 %.3f ms (%.2f TFLOP/s)  K=1 %.3f ms  K=2 %.3f ms  K=3 %.3f ms   ratios %.3f %.3f %.3f\n", t0,
           2 * dfma / (t0 * 1e-3) / 1e12, t1, t2, t3, t1 / t0, t2 / t0, t3 / t0);
    return 0;
}
