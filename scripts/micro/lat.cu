// Latency micro-benchmarks for the FPS inner loop primitives on sm_100a (dependent chains timed with clock64).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N 512
__global__ void k_redux(unsigned *out, long long *cyc, unsigned seed) {
    unsigned v = seed + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) v = __reduce_max_sync(0xffffffffu, v ^ i) + threadIdx.x;
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = v; cyc[0] = (t1 - t0); }
}
__global__ void k_redux2(unsigned *out, long long *cyc, unsigned seed) {  // max then dependent min (warp_argmax)
    unsigned v = seed + threadIdx.x, p = threadIdx.x * 7u;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
        unsigned vm = __reduce_max_sync(0xffffffffu, v ^ i);
        unsigned q = (v ^ i) == vm ? p : 0xffffffffu;
        p = __reduce_min_sync(0xffffffffu, q) + threadIdx.x;
        v = vm + threadIdx.x;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = v + p; cyc[0] = (t1 - t0); }
}
__global__ void k_shfl_argmax(unsigned *out, long long *cyc, unsigned seed) {  // 5-level butterfly on (v,p)
    unsigned v = seed + threadIdx.x, p = threadIdx.x * 7u;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
        unsigned a = v ^ i, b = p;
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            unsigned a2 = __shfl_xor_sync(0xffffffffu, a, o), b2 = __shfl_xor_sync(0xffffffffu, b, o);
            bool take = a2 > a || (a2 == a && b2 < b);
            a = take ? a2 : a; b = take ? b2 : b;
        }
        v = a + threadIdx.x; p = b + threadIdx.x;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = v + p; cyc[0] = (t1 - t0); }
}
__global__ void k_shfl64(unsigned *out, long long *cyc, unsigned seed) {  // 5-level butterfly on one packed 64-bit key
    unsigned long long k = ((unsigned long long)(seed + threadIdx.x) << 32) | (threadIdx.x * 7u);
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
        unsigned long long a = k ^ i;
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            unsigned long long a2 = __shfl_xor_sync(0xffffffffu, a, o);
            a = a2 > a ? a2 : a;
        }
        k = a + threadIdx.x;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = (unsigned)k; cyc[0] = (t1 - t0); }
}
__global__ void k_bar(unsigned *out, long long *cyc) {
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = 1; cyc[0] = (t1 - t0); }
}
__global__ void k_bar_lds(unsigned *out, long long *cyc) {  // sts -> bar -> lds dependent round trip
    __shared__ unsigned s[64];
    unsigned v = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
        if ((threadIdx.x & 31) == 0) s[(i & 1) * 32 + (threadIdx.x >> 5)] = v;
        __syncthreads();
        v = s[(i & 1) * 32 + (threadIdx.x & 31) % (blockDim.x >> 5)] + 1;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = v; cyc[0] = (t1 - t0); }
}
__global__ void k_lds(unsigned *out, long long *cyc) {
    __shared__ unsigned s[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = (i * 37 + 11) & 1023;
    __syncthreads();
    unsigned v = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) v = s[v];
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = v; cyc[0] = (t1 - t0); }
}
__global__ void k_vote(unsigned *out, long long *cyc, unsigned seed) {
    unsigned v = seed + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) v = __ballot_sync(0xffffffffu, (v ^ i) & 1) + threadIdx.x;
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = v; cyc[0] = (t1 - t0); }
}
__global__ void k_match(unsigned *out, long long *cyc, unsigned seed) {  // fp chain of 6 (lower bound)
    float v = seed + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) { v = fmaxf(v - 1.5f, 0.f); v = v * v; v = fmaf(v, v, 1.0f); v = fmaf(v, v, 2.0f); v = fminf(v, 7.f); }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = (unsigned)v; cyc[0] = (t1 - t0); }
}
int main() {
    unsigned *o; long long *c; cudaMalloc(&o, 64); cudaMalloc(&c, 64);
    long long h;
#define RUN(name, threads, ...) for (int r = 0; r < 2; ++r) { name<<<1, threads>>>(__VA_ARGS__); cudaDeviceSynchronize(); } cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("%-28s threads=%4d : %7.1f cycles/iter\n", #name, threads, (double)h / N);
    RUN(k_redux, 32, o, c, 5)
    RUN(k_redux, 512, o, c, 5)
    RUN(k_redux2, 32, o, c, 5)
    RUN(k_redux2, 512, o, c, 5)
    RUN(k_shfl_argmax, 32, o, c, 5)
    RUN(k_shfl_argmax, 512, o, c, 5)
    RUN(k_shfl64, 32, o, c, 5)
    RUN(k_shfl64, 512, o, c, 5)
    RUN(k_bar, 256, o, c)
    RUN(k_bar, 512, o, c)
    RUN(k_bar, 1024, o, c)
    RUN(k_bar_lds, 256, o, c)
    RUN(k_bar_lds, 512, o, c)
    RUN(k_lds, 32, o, c)
    RUN(k_vote, 32, o, c, 5)
    RUN(k_match, 32, o, c, 5)
    cudaError_t e = cudaGetLastError(); printf("%s\n", cudaGetErrorString(e));
    return 0;
}
