// FFMA2 / FADD2 (packed fp32x2, sm_100) vs scalar FFMA: dependent-chain latency and per-SMSP throughput.
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
template <int CHAINS, bool PACKED>
__global__ void k(float *out, long long *cyc, float a) {
    float2 acc[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc[c] = make_float2(threadIdx.x * 0.001f + c, 1.f + c);
    float2 t = make_float2(a, a * 0.5f);
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N / 32; ++i) {
#pragma unroll
      for (int rep = 0; rep < 32; ++rep)
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (PACKED) acc[c] = __ffma2_rn(t, t, acc[c]);
            else { acc[c].x = __fmaf_rn(t.x, t.x, acc[c].x); acc[c].y = __fmaf_rn(t.y, t.y, acc[c].y); }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += acc[c].x + acc[c].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    float *o; long long *c; cudaMalloc(&o, 1 << 20); cudaMalloc(&c, 64); long long h;
#define RUN(CH, PK, TH) for (int r = 0; r < 2; ++r) { k<CH, PK><<<1, TH>>>(o, c, 1.0001f); cudaDeviceSynchronize(); } cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); \
    printf("chains=%d packed=%d threads=%4d : %6.2f cycles per loop iter, %5.2f cycles per fp32x2-op per warp-slot\n", CH, PK, TH, (double)h / N, (double)h / N / CH);
    RUN(1, true, 32) RUN(1, false, 32) RUN(2, true, 32) RUN(2, false, 32) RUN(1, true, 256) RUN(2, true, 256) RUN(1, false, 512) RUN(8, true, 32) RUN(8, false, 32)
    RUN(8, true, 128) RUN(8, false, 128) RUN(8, true, 256) RUN(8, false, 256) RUN(8, true, 512) RUN(8, false, 512)
    RUN(1, true, 256) RUN(1, false, 256)
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
