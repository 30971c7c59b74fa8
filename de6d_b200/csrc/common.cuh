// Shared device helpers for the de6d_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define DE6D_OK 0
#define DE6D_ERR_INVALID 1   // bad argument (negative size, null pointer, unsupported shape)
#define DE6D_ERR_CUDA 2      // a CUDA runtime call or launch failed (see de6d_last_error_string)

namespace de6d {

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// Squared distance exactly as the reference's nvcc build evaluates
//   (ax-bx)*(ax-bx) + (ay-by)*(ay-by) + (az-bz)*(az-bz)
// (sampling_gpu.cu:143, ball_query_gpu.cu:39, interpolate_gpu.cu:41): the y product is a plain rounded
// FMUL, x and z are FFMAs.  Spelled with intrinsics so no other contraction can happen.
__device__ __forceinline__ float sqdist(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// Two squared distances at once on the packed-fp32 pipe (sm_100 FADD2 / FMUL2 / FFMA2): the points (a.x, a.y) against ONE
// point b, lane-wise the same IEEE operations as sqdist() above (a + (-b) rounds like a - b), hence the same bits.
__device__ __forceinline__ float2 sqdist2(float2 ax, float2 ay, float2 az, float bx, float by, float bz) {
    const float2 dx = __fadd2_rn(ax, make_float2(-bx, -bx)), dy = __fadd2_rn(ay, make_float2(-by, -by)), dz = __fadd2_rn(az, make_float2(-bz, -bz));
    return __ffma2_rn(dz, dz, __ffma2_rn(dx, dx, __fmul2_rn(dy, dy)));
}

// Monotone map float -> uint32 (total order of finite floats, -0 < +0).
__device__ __forceinline__ uint32_t f2ord(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// ---- the reference FPS tie rule --------------------------------------------------------------------------
// The reference reduces per-thread candidates with a shared-memory tree that keeps the lower slot on ties
// (sampling_gpu.cu:94-99,155-215) after a strided in-thread scan with strict '>' (:146-147).  The winner among
// equal values is therefore the point k with the smallest bit-reversed slot (k mod B), then the smallest k / B,
// with B = opt_n_threads(N) (cuda_utils.h:10-14).  prio(k) packs that order into 32 bits, smaller wins.
__device__ __forceinline__ uint32_t fps_prio(uint32_t k, uint32_t log2B) {
    uint32_t slot = k & ((1u << log2B) - 1u);
    uint32_t rev = log2B ? (__brev(slot) >> (32 - log2B)) : 0u;
    return (rev << 22) | (k >> log2B);
}
__device__ __forceinline__ uint32_t fps_prio_to_index(uint32_t prio, uint32_t log2B) {
    uint32_t rev = prio >> 22;
    uint32_t slot = log2B ? (__brev(rev) >> (32 - log2B)) : 0u;
    return ((prio & 0x3FFFFFu) << log2B) | slot;
}

// Warp arg-max of (value-bits, priority) with two REDUX instructions.  `v` must be an order-preserving
// uint32 image of the value; among equal v the smallest `prio` wins.  Returns the winning prio in `prio`.
__device__ __forceinline__ void warp_argmax(uint32_t &v, uint32_t &prio) {
    uint32_t vm = __reduce_max_sync(0xffffffffu, v);
    uint32_t p = (v == vm) ? prio : 0xffffffffu;
    prio = __reduce_min_sync(0xffffffffu, p);
    v = vm;
}

// ---- helpers shared by the on-chip sampling kernels (fps.cu, fps_features.cu) -----------------------------------------
__device__ __forceinline__ float ord2f(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__device__ __forceinline__ uint32_t part1by2(uint32_t x) {  // spread 10 bits to every third bit
    x &= 0x3ffu;
    x = (x | (x << 16)) & 0x030000ffu;
    x = (x | (x << 8)) & 0x0300f00fu;
    x = (x | (x << 4)) & 0x030c30c3u;
    x = (x | (x << 2)) & 0x09249249u;
    return x;
}

// compact priority for clouds of at most 16384 points: (bit-reversed slot, k / B) in <= 14 bits
__device__ __forceinline__ uint32_t cprio_of(uint32_t k, int log2B, int ibits) {
    uint32_t slot = k & ((1u << log2B) - 1u);
    uint32_t rev = log2B ? (__brev(slot) >> (32 - log2B)) : 0u;
    return (rev << ibits) | (k >> log2B);
}
__device__ __forceinline__ uint32_t index_of_cprio(uint32_t cp, int log2B, int ibits) {
    uint32_t rev = cp >> ibits;
    uint32_t slot = log2B ? (__brev(rev) >> (32 - log2B)) : 0u;
    return ((cp & ((1u << ibits) - 1u)) << log2B) | slot;
}

// Lower bound of sqdist(p, q) over every p inside the box: per axis the gap g = max(lo-q, q-hi, 0) satisfies
// |fl(p-q)| >= g (rounding is monotone), and fl(dy*dy), fl(dx*dx+t), fl(dz*dz+t) are monotone in |d.| and t,
// so the value below never exceeds the distance the kernel would compute for any point of the bucket.
__device__ __forceinline__ float bucket_lower_bound(float lox, float hix, float loy, float hiy, float loz, float hiz,
                                                    float qx, float qy, float qz) {
    float gx = fmaxf(fmaxf(__fsub_rn(lox, qx), __fsub_rn(qx, hix)), 0.f);
    float gy = fmaxf(fmaxf(__fsub_rn(loy, qy), __fsub_rn(qy, hiy)), 0.f);
    float gz = fmaxf(fmaxf(__fsub_rn(loz, qz), __fsub_rn(qz, hiz)), 0.f);
    return __fmaf_rn(gz, gz, __fmaf_rn(gx, gx, __fmul_rn(gy, gy)));
}

// The same bound for TWO query points at once (packed fp32): lane-wise the operations of bucket_lower_bound.
__device__ __forceinline__ float2 bucket_lower_bound2(float lox, float hix, float loy, float hiy, float loz, float hiz,
                                                      float2 qx, float2 qy, float2 qz) {
    const float2 ax = __fadd2_rn(make_float2(lox, lox), make_float2(-qx.x, -qx.y)), bx = __fadd2_rn(qx, make_float2(-hix, -hix));
    const float2 ay = __fadd2_rn(make_float2(loy, loy), make_float2(-qy.x, -qy.y)), by = __fadd2_rn(qy, make_float2(-hiy, -hiy));
    const float2 az = __fadd2_rn(make_float2(loz, loz), make_float2(-qz.x, -qz.y)), bz = __fadd2_rn(qz, make_float2(-hiz, -hiz));
    const float2 gx = make_float2(fmaxf(fmaxf(ax.x, bx.x), 0.f), fmaxf(fmaxf(ax.y, bx.y), 0.f));
    const float2 gy = make_float2(fmaxf(fmaxf(ay.x, by.x), 0.f), fmaxf(fmaxf(ay.y, by.y), 0.f));
    const float2 gz = make_float2(fmaxf(fmaxf(az.x, bz.x), 0.f), fmaxf(fmaxf(az.y, bz.y), 0.f));
    return __ffma2_rn(gz, gz, __ffma2_rn(gx, gx, __fmul2_rn(gy, gy)));
}

// ---- TMA bulk copy + mbarrier + streaming access helpers (group_gather.cu, interpolate.cu) ----------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
// invalidate a barrier object before its CTA exits (PTX: required before the memory is re-used, e.g. re-initialised by the next
// CTA on the SM; compute-sanitizer's synccheck otherwise reports the next cluster's barriers as divergent)
__device__ __forceinline__ void mbar_inval(uint64_t *bar) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
    // the spin loop lives inside one asm block, so the compiler places no reconvergence point behind it: make the lanes
    // meet again explicitly before anything warp-synchronous (REDUX, bar.sync) follows
    __syncwarp();
}
// 1-D TMA bulk copy global -> shared (SASS: UBLKCP); bytes and both addresses must be multiples of 16.
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ int4 ldg_stream_int4(const int *p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream_float4(float *p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

}  // namespace de6d

// Host-side error plumbing shared by all translation units (defined in capi.cu).
extern "C" const char *de6d_last_error_string(void);
int de6d_set_cuda_error(cudaError_t e, const char *where);
int de6d_set_error(int code, const char *msg);

void de6d_count_launch(void);

#define DE6D_CHECK_LAUNCH(where)                                  \
    do {                                                          \
        cudaError_t e__ = cudaGetLastError();                     \
        if (e__ != cudaSuccess) return de6d_set_cuda_error(e__, where); \
        de6d_count_launch();                                      \
    } while (0)

// Dynamic shared memory above 48 KB is an opt-in attribute of a (function, device) pair: set it once per pair.
// `mask` is a per-call-site static bit set of the devices already configured, read and updated atomically (entry points
// may be called from several host threads; setting the attribute twice is harmless, losing a bit only repeats it).
template <typename F>
static inline int de6d_ensure_smem(F func, int bytes, unsigned long long &mask, const char *what) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && ((__atomic_load_n(&mask, __ATOMIC_RELAXED) >> dev) & 1ull)) return DE6D_OK;
    cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();   // the failure is reported through the status code; do not leave it pending for the next CUDA user
        return de6d_set_cuda_error(e, what);
    }
    if (dev >= 0 && dev < 64) __atomic_fetch_or(&mask, 1ull << dev, __ATOMIC_RELAXED);
    return DE6D_OK;
}
