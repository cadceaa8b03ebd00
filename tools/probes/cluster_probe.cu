// How many thread-block clusters of size S (one 768-thread CTA with ~214 KB of dynamic shared memory per SM)
// can be co-resident on this GPU: decides whether a 3-CTA cluster (one CTA per sampling mode) wastes SMs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o cluster_probe cluster_probe.cu && ./cluster_probe
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(768, 1) probe_kernel(int *out)
{
    extern __shared__ unsigned char smem[];
    if (threadIdx.x == 0 && out) out[blockIdx.x] = smem[0];
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s: %d SMs\n", p.name, p.multiProcessorCount);
    const int smem = 218 * 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int s = 1; s <= 8; ++s) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(s * 64);
        cfg.blockDim = dim3(768);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = s; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = -1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, probe_kernel, &cfg);
        printf("cluster size %d: max active clusters %d (%d SMs used)%s\n", s, n, n * s, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    return 0;
}
