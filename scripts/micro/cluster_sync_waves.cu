// Is compute-sanitizer synccheck's "Divergent thread(s) in warp" on second-wave thread-block clusters a property of our kernels?
// A kernel that does nothing but cluster.sync() + a loop of __syncthreads(), one CTA per SM (200 KB of shared memory), launched with
// more clusters than fit at once.  nvcc -gencode arch=compute_100a,code=sm_100a -o cluster_sync_waves cluster_sync_waves.cu
#include <cooperative_groups.h>
#include <cstdio>
namespace cg = cooperative_groups;

__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(256, 1) waves_kernel(int iters, int *out) {
    extern __shared__ int sm[];
    cg::cluster_group cluster = cg::this_cluster();
    sm[threadIdx.x] = (int)threadIdx.x;
    cluster.sync();
    int acc = 0;
    for (int i = 0; i < iters; ++i) {
        acc += sm[(threadIdx.x + i) & 255];
        __syncthreads();
        sm[threadIdx.x] = acc;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = acc;
    cluster.sync();
}

int main(int argc, char **argv) {
    const int clusters = argc > 1 ? atoi(argv[1]) : 40;      // 40 clusters of 4 = 160 CTAs on 148 SMs: two waves; 30: one wave
    int *out;
    cudaMalloc(&out, clusters * 4 * sizeof(int));
    cudaFuncSetAttribute(waves_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    waves_kernel<<<clusters * 4, 256, 200 * 1024>>>(2000, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("clusters %d: %s\n", clusters, cudaGetErrorString(e));
    return e != cudaSuccess;
}
