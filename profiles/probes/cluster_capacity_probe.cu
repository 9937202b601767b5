// How many thread-block clusters of a given size are resident on the device at one CTA per SM?
// (cudaOccupancyMaxActiveClusters; a cluster must sit inside one GPC, so sizes that do not divide a GPC's SM count strand SMs.)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o cluster_capacity_probe cluster_capacity_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(320, 1) probe_kernel(double *out) {
    extern __shared__ double sm[];
    sm[threadIdx.x] = threadIdx.x;
    __syncthreads();
    if (out) out[blockIdx.x * blockDim.x + threadIdx.x] = sm[(threadIdx.x + 1) % blockDim.x];
}
int main() {
    const size_t smem = 161136;  // the cluster tile kernel's dynamic shared memory for triple @ 200 elements
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int size = 1; size <= 8; ++size) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(size * 64);
        cfg.blockDim = dim3(320);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = size;
        attr[0].val.clusterDim.y = attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int clusters = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&clusters, probe_kernel, &cfg);
        printf("cluster size %d: %d clusters resident = %d of %d SMs (%s)\n", size, clusters, clusters * size, sms, cudaGetErrorString(e));
    }
    return 0;
}
