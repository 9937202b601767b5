// How much fp64 instruction-level parallelism does one SM sub-partition of B200 need?
// C independent DFMA chains per thread, W warps per SMSP (one CTA of 4*W warps per SM).
// Prints cycles per DFMA per SMSP (2.0 = the pipe's peak: 16 lanes per clock).
#include <cstdio>
#include <cuda_runtime.h>

template <int C>
__global__ void probe(double *out, long long *cyc, int iters, double seed) {
    double a[C];
#pragma unroll
    for (int i = 0; i < C; ++i) a[i] = seed + i * 1e-3 + threadIdx.x * 1e-9;
    const double m = 0.999999, d = 1e-7;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
#pragma unroll
            for (int i = 0; i < C; ++i) a[i] = fma(a[i], m, d);
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < C; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int C>
void run(double *out, long long *cyc, int sms) {
    for (int W = 1; W <= 8; ++W) {
        const int threads = 128 * W, iters = 2000;
        if (threads > 1024) break;
        probe<C><<<sms, threads>>>(out, cyc, iters, 1.0);
        probe<C><<<sms, threads>>>(out, cyc, iters, 2.0);
        cudaDeviceSynchronize();
        long long c;
        cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
        const double per = (double)c / ((double)iters * 16 * C * W);
        printf("chains %d  warps/SMSP %d : %.2f cycles per DFMA per SMSP (%.0f%% of peak); dependent-issue interval %.1f\n", C, W,
               per, 200.0 / per, (double)c / ((double)iters * 16));
    }
}

int main() {
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *out;
    long long *cyc;
    cudaMalloc(&out, sizeof(double) * sms * 1024);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    run<1>(out, cyc, sms);
    run<2>(out, cyc, sms);
    run<3>(out, cyc, sms);
    run<4>(out, cyc, sms);
    run<6>(out, cyc, sms);
    run<8>(out, cyc, sms);
    return 0;
}
