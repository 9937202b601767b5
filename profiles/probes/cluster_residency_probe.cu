// How many CTAs of a cluster launch are REALLY co-resident (one CTA per SM by shared memory)?  Every CTA counts itself in,
// then stays for 20 ms (globaltimer), records the highest count it saw and counts itself out.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o cluster_residency_probe cluster_residency_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ unsigned long long now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#ifndef LIVE
#define LIVE 0  // doubles kept live across the wait: LIVE=72 makes the kernel a 168-register one like the tile kernel
#endif
__global__ void __launch_bounds__(320, 1) probe_kernel(unsigned int *counter, unsigned int *peak, unsigned int *smid_seen) {
    extern __shared__ double sm[];
    double live[LIVE + 1];
#pragma unroll
    for (int i = 0; i <= LIVE; ++i) live[i] = sm[(threadIdx.x + i) % 64] + i;
    if (threadIdx.x == 0) {
        const unsigned long long t0 = now_ns();
        atomicAdd(counter, 1u);
        unsigned int sid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sid));
        unsigned int best = 0;
        while (now_ns() - t0 < 20000000ull) {
            const unsigned int c = *(volatile unsigned int *)counter;
            best = c > best ? c : best;
        }
        atomicMax(peak, best);
        atomicSub(counter, 1u);  // (a CTA of a later wave must not count the ones that have left)
        if (best == *(volatile unsigned int *)peak) atomicOr(&smid_seen[sid / 32], 1u << (sid % 32));
        sm[0] = best;
    }
    __syncthreads();
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i <= LIVE; ++i) acc = fma(acc, live[i], sm[i % 64]);
    if (acc == 12345.678) *counter = 0;
}
int main() {
    const size_t smem = 161136;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    unsigned int *d;
    cudaMalloc(&d, 64);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, probe_kernel);
    printf("probe kernel: %d registers, %zu B static shared\n", fa.numRegs, fa.sharedSizeBytes);
    for (int size = 1; size <= 4; ++size) {
        for (int clusters : {45, 46, 47, 48, 49, 74, 148}) {
            if (clusters * size > 148 + 8) continue;
            cudaMemset(d, 0, 64);
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(size * clusters);
            cfg.blockDim = dim3(320);
            cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = size;
            attr[0].val.clusterDim.y = attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            void *args[] = {(void *)&d, nullptr, nullptr};
            unsigned int *peak = d + 1, *seen = d + 2;
            args[1] = &peak;
            args[2] = &seen;
            cudaError_t e = cudaLaunchKernelExC(&cfg, (const void *)probe_kernel, args);
            cudaError_t e2 = cudaDeviceSynchronize();
            unsigned int h[16];
            cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
            printf("cluster size %d, %d clusters (%d CTAs) launched: %u CTAs co-resident (%s / %s)\n", size, clusters, size * clusters, h[1],
                   cudaGetErrorString(e), cudaGetErrorString(e2));
        }
    }
    return 0;
}
