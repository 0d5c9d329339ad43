// How many clusters of a 1-CTA-per-SM kernel (200 KiB of dynamic shared memory, 640 threads) can be co-resident on this GPU?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dummy(int* p) { extern __shared__ char sm[]; if (p) p[0] = sm[0]; }
int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    printf("%s SMs %d\n", pr.name, pr.multiProcessorCount);
    const int smem = 200 * 1024;
    cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int c : {1, 2, 4, 8, 16}) {
        cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(c * 64); cfg.blockDim = dim3(640); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = c; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy, &cfg);
        printf("cluster %2d: max active clusters %d (%d CTAs)  %s\n", c, n, n * c, cudaGetErrorString(e));
    }
    return 0;
}
