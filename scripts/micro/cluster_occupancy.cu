// How many thread-block clusters of size S are co-resident on this GPU when every CTA takes a whole SM
// (~190 KB dynamic shared memory)?   nvcc -arch=sm_100a -o /tmp/cluster_occupancy cluster_occupancy.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dummy(float *p) {
    extern __shared__ float s[];
    s[threadIdx.x] = 1.f;
    if (p) p[0] = s[0];
}

int main() {
    cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int S = 1; S <= 16; ++S) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(S * 64);
        cfg.blockDim = dim3(256);
        cfg.dynamicSmemBytes = 190 * 1024;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = S; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int nc = -1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&nc, dummy, &cfg);
        printf("cluster size %2d: %3d clusters resident = %3d SMs busy  (%s)\n", S, nc, nc * S, cudaGetErrorString(e));
    }
    return 0;
}
