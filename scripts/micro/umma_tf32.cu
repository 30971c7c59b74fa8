// Bring-up test for the tcgen05 path used by csrc/sa_mlp.cu: one CTA computes
//     D1[128 x N1] = A[128 x K1] * W1[N1 x K1]^T            A, W1 in shared memory (SWIZZLE_128B, K-major), kind::tf32
//     D2[128 x N2] = relu(D1)[128 x N1] * W2[N2 x N1]^T     A operand read from TENSOR MEMORY (written back with tcgen05.st)
// and checks both against a host reference that rounds the operands to tf32 the same way (cvt.rna).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/micro/umma_tf32 scripts/micro/umma_tf32.cu
// Every wait is bounded (trap after ~1 s) so a wrong descriptor cannot hang the GPU.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);   // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                      // leading byte offset (unused for swizzled K-major), bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;            // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                      // version = 1 (sm_100)
    d |= (uint64_t)2 << 61;                      // layout type SWIZZLE_128B
    return d;
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128
__host__ __device__ inline uint32_t idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// byte offset of element (r, k) of an [R x K] fp32 matrix in the SWIZZLE_128B K-major image (K blocks of 32 elements)
__host__ __device__ inline uint32_t sw128_off(int R, int r, int k) {
    const int kb = k >> 5, c = (k & 31) >> 2;
    return (uint32_t)kb * (uint32_t)R * 128u + (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)((c ^ (r & 7)) << 4) +
           (uint32_t)(k & 3) * 4u;
}

__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t *bar, uint32_t parity) {
    for (long long it = 0; it < 20000000ll; ++it) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t *v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t *v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                 "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}

// A (128 x K1), W1 (N1 x K1), W2 (N2 x N1) row-major in global; out1 (128 x N1), out2 (128 x N2); status[0] = error code
__global__ void __launch_bounds__(128) umma_test(int K1, int N1, int N2, const float *A, const float *W1, const float *W2, float *out1,
                                                 float *out2, int *status) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int K1p = (K1 + 31) & ~31, N1p = (N1 + 31) & ~31;
    unsigned char *sA = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);   // SWIZZLE_128B atoms need 1024-byte alignment; 128 x K1p
    unsigned char *sW1 = sA + (size_t)128 * K1p * 4;           // N1 x K1p
    unsigned char *sW2 = sW1 + (size_t)N1 * K1p * 4;           // N2 x N1p   (all sizes multiples of 1024)
    // operands -> swizzled images, rounded to tf32 (zero padded in K)
    for (int i = tid; i < 128 * K1p; i += 128) { int r = i / K1p, k = i % K1p; *(float *)(sA + sw128_off(128, r, k)) = k < K1 ? to_tf32(A[r * K1 + k]) : 0.f; }
    for (int i = tid; i < N1 * K1p; i += 128) { int r = i / K1p, k = i % K1p; *(float *)(sW1 + sw128_off(N1, r, k)) = k < K1 ? to_tf32(W1[r * K1 + k]) : 0.f; }
    for (int i = tid; i < N2 * N1p; i += 128) { int r = i / N1p, k = i % N1p; *(float *)(sW2 + sw128_off(N2, r, k)) = k < N1 ? to_tf32(W2[r * N1 + k]) : 0.f; }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy smem writes -> visible to the tensor core (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmem_base_s;
    const uint32_t tD1 = tbase, tD2 = tbase + 256;
    bool ok = true;

    // ---- layer 1: both operands from shared memory ----
    if (tid == 0) {
        const uint32_t id = idesc_tf32(N1);
        for (int s = 0; s < K1p / 8; ++s) {
            const uint32_t koff = (uint32_t)(s >> 2) * 128u, inner = (uint32_t)(s & 3) * 32u;
            mma_ss(tD1, desc_sw128(smem_u32(sA) + koff * 128u + inner), desc_sw128(smem_u32(sW1) + koff * (uint32_t)N1 + inner), id, s > 0);
        }
        mma_commit(&bar);
    }
    ok = mbar_wait_bounded(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!ok) { if (tid == 0) status[0] = 1; }
    // ---- epilogue 1: D1 -> registers -> out1; relu + tf32 rounding -> back into the same TMEM columns as layer 2's A ----
    if (ok) {
        const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
        for (int c0 = 0; c0 < N1; c0 += 8) {
            uint32_t v[8];
            tmem_ld8(tD1 + lane_addr + c0, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 8; ++j) {
                out1[tid * N1 + c0 + j] = __uint_as_float(v[j]);
                v[j] = __float_as_uint(to_tf32(fmaxf(__uint_as_float(v[j]), 0.f)));
            }
            tmem_st8(tD1 + lane_addr + c0, v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- layer 2: A from tensor memory ----
    if (tid == 0 && status[0] == 0) {
        const uint32_t id = idesc_tf32(N2);
        for (int s = 0; s < N1 / 8; ++s) {
            const uint32_t koff = (uint32_t)(s >> 2) * 128u, inner = (uint32_t)(s & 3) * 32u;
            mma_ts(tD2, tD1 + (uint32_t)s * 8u, desc_sw128(smem_u32(sW2) + koff * (uint32_t)N2 + inner), id, s > 0);
        }
        mma_commit(&bar);
    }
    __syncthreads();
    if (status[0] == 0) {
        ok = mbar_wait_bounded(&bar, 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (!ok) { if (tid == 0) status[0] = 2; }
        if (ok) {
            const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
            for (int c0 = 0; c0 < N2; c0 += 8) {
                uint32_t v[8];
                tmem_ld8(tD2 + lane_addr + c0, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                for (int j = 0; j < 8; ++j) out2[tid * N2 + c0 + j] = __uint_as_float(v[j]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
}

static float tf32_host(float x) {   // round to nearest, ties away (cvt.rna): add half an ulp of the 10-bit mantissa, truncate
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return x;
    u += 0x1000u;
    u &= 0xffffe000u;
    float r;
    memcpy(&r, &u, 4);
    return r;
}

int main() {
    const int shapes[][3] = {{8, 16, 16}, {32, 16, 32}, {72, 64, 128}, {136, 128, 256}, {264, 256, 256}};
    int fails = 0;
    for (auto &sh : shapes) {
        const int K1 = sh[0], N1 = sh[1], N2 = sh[2];
        std::vector<float> A(128 * K1), W1(N1 * K1), W2(N2 * N1), o1(128 * N1), o2(128 * N2);
        srand(K1 * 131 + N1);
        for (auto &x : A) x = (rand() % 2001 - 1000) / 500.f;
        for (auto &x : W1) x = (rand() % 2001 - 1000) / 1000.f;
        for (auto &x : W2) x = (rand() % 2001 - 1000) / 1000.f;
        float *dA, *dW1, *dW2, *dO1, *dO2;
        int *dS;
        CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dW1, W1.size() * 4)); CK(cudaMalloc(&dW2, W2.size() * 4));
        CK(cudaMalloc(&dO1, o1.size() * 4)); CK(cudaMalloc(&dO2, o2.size() * 4)); CK(cudaMalloc(&dS, 4));
        CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dW1, W1.data(), W1.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dW2, W2.data(), W2.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemset(dS, 0, 4)); CK(cudaMemset(dO1, 0, o1.size() * 4)); CK(cudaMemset(dO2, 0, o2.size() * 4));
        const int K1p = (K1 + 31) & ~31, N1p = (N1 + 31) & ~31;
        const size_t smem = (size_t)128 * K1p * 4 + (size_t)N1 * K1p * 4 + (size_t)N2 * N1p * 4 + 1024;
        if (smem > 227 * 1024) { printf("shape %d %d %d: skipped (%zu B smem)\n", K1, N1, N2, smem); continue; }
        CK(cudaFuncSetAttribute(umma_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        umma_test<<<1, 128, smem>>>(K1, N1, N2, dA, dW1, dW2, dO1, dO2, dS);
        cudaError_t e = cudaDeviceSynchronize();
        int st = -1;
        if (e == cudaSuccess) CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
        if (e != cudaSuccess || st != 0) { printf("shape %d %d %d: FAILED launch (%s, status %d)\n", K1, N1, N2, cudaGetErrorString(e), st); return 3; }
        CK(cudaMemcpy(o1.data(), dO1, o1.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(o2.data(), dO2, o2.size() * 4, cudaMemcpyDeviceToHost));
        double e1 = 0, e2 = 0, m1 = 0, m2 = 0;
        std::vector<float> h1(128 * N1);
        for (int r = 0; r < 128; ++r)
            for (int n = 0; n < N1; ++n) {
                double acc = 0;
                for (int k = 0; k < K1; ++k) acc += (double)tf32_host(A[r * K1 + k]) * tf32_host(W1[n * K1 + k]);
                h1[r * N1 + n] = (float)acc;
                e1 = fmax(e1, fabs(acc - o1[r * N1 + n])); m1 = fmax(m1, fabs(acc));
            }
        for (int r = 0; r < 128; ++r)
            for (int n = 0; n < N2; ++n) {
                double acc = 0;
                for (int k = 0; k < N1; ++k) acc += (double)tf32_host(fmaxf(o1[r * N1 + k], 0.f)) * tf32_host(W2[n * N1 + k]);
                e2 = fmax(e2, fabs(acc - o2[r * N2 + n])); m2 = fmax(m2, fabs(acc));
            }
        const bool good = e1 <= 1e-4 * m1 && e2 <= 1e-4 * m2 && m1 > 0 && m2 > 0;
        printf("shape K1=%d N1=%d N2=%d: layer1 (SS) max err %.3e of %.3e, layer2 (TS) max err %.3e of %.3e  %s\n", K1, N1, N2, e1, m1, e2, m2,
               good ? "OK" : "MISMATCH");
        fails += !good;
        cudaFree(dA); cudaFree(dW1); cudaFree(dW2); cudaFree(dO1); cudaFree(dO2); cudaFree(dS);
    }
    printf(fails ? "UMMA TEST FAILED\n" : "UMMA TEST PASSED\n");
    return fails ? 1 : 0;
}
