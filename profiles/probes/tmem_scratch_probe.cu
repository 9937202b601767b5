// Tensor memory as per-thread scratch (round-2 candidate): how expensive is it to park registers in
// TMEM with tcgen05.st and fetch them back with tcgen05.ld, for CTAs of two warps (the tile kernel's shape)?
// Each thread owns one TMEM lane and `kCols` 32-bit columns of it (32x32b shape: lane i of the warp <-> TMEM lane
// 32 * (warp % 4) + i).  Measures (a) the dependent round trip st -> wait -> ld -> wait, (b) throughput of independent
// 8-byte loads with one wait per batch, as a function of CTAs per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kCols = 64;  // columns allocated per CTA (power of two >= 32): 64 x 4 B = 32 doubles per thread

__device__ __forceinline__ void tmem_st2(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void tmem_ld2(uint32_t addr, uint32_t &a, uint32_t &b) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(64) probe(long long *cycles, double *out, int iters) {
    __shared__ uint32_t base_slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"l"((uint64_t)__cvta_generic_to_shared(&base_slot)), "n"(kCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = base_slot + ((uint32_t)(32 * (warp & 3)) << 16);  // my warp's lane quadrant

    // fill my columns
    for (int c = 0; c < kCols; c += 2) tmem_st2(base + c, threadIdx.x * 1000u + c, c);
    tmem_wait_st();

    // (a) dependent round trip: the loaded value is stored back, loaded again, ...
    uint32_t a = threadIdx.x, b = 1;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        tmem_st2(base + 2 * (i & 15), a + 1, b);
        tmem_wait_st();
        tmem_ld2(base + 2 * (i & 15), a, b);
        tmem_wait_ld();
    }
    long long t1 = clock64();
    // (b) 16 independent 8-byte loads per batch (= 16 doubles per thread), one wait per batch
    uint32_t r[32];
    double acc = 0.0;
    __syncthreads();
    long long t2 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < 16; ++c) tmem_ld2(base + 2 * c, r[2 * c], r[2 * c + 1]);
        tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < 16; ++c) acc += __hiloint2double((int)r[2 * c + 1], (int)r[2 * c]) * 1e-300;
    }
    long long t3 = clock64();
    // (c) 16 independent 8-byte stores per batch, one wait per batch
    __syncthreads();
    long long t4 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < 16; ++c) tmem_st2(base + 2 * c, a + i, c);
        tmem_wait_st();
    }
    long long t5 = clock64();
    if (threadIdx.x == 0) {
        cycles[3 * blockIdx.x] = t1 - t0;
        cycles[3 * blockIdx.x + 1] = t3 - t2;
        cycles[3 * blockIdx.x + 2] = t5 - t4;
    }
    out[blockIdx.x * 64 + threadIdx.x] = acc + a + b;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base_slot), "n"(kCols));
}

int main() {
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    long long *cyc;
    double *out;
    cudaMalloc(&cyc, sizeof(long long) * 3 * sms * 8);
    cudaMalloc(&out, sizeof(double) * 64 * sms * 8);
    const int iters = 2000;
    for (int per_sm = 1; per_sm <= 8; ++per_sm) {
        probe<<<sms * per_sm, 64>>>(cyc, out, iters);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("CTAs/SM %d: %s\n", per_sm, cudaGetErrorString(e));
            return 1;
        }
        long long h[3];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("CTAs/SM %d (2 warps each): st+ld round trip %.1f cycles | 16 doubles/thread loaded in %.1f cycles "
               "(%.1f B/clk/SM) | 16 doubles/thread stored in %.1f cycles (%.1f B/clk/SM)\n",
               per_sm, (double)h[0] / iters, (double)h[1] / iters, 64.0 * per_sm * 128.0 / ((double)h[1] / iters),
               (double)h[2] / iters, 64.0 * per_sm * 128.0 / ((double)h[2] / iters));
    }
    return 0;
}
