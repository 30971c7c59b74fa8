// three_nn / three_interpolate (+ gradient) for sm_100a.
//
// Replaces pointnet2_batch/src/interpolate_gpu.cu:16-149.
//   three_nn: thread per unknown point, known points streamed through shared-memory tiles (broadcast reads)
//     instead of serial global loads.  The reference keeps its three running minima in double initialised to
//     1e40; everything ever stored in them is a float, so float minima initialised to +inf with the same
//     strict '<' cascade select the same indices and store the same values (1e40 -> +inf on the final cast).
//   three_interpolate: out = fmaf(w2,p2, fmaf(w0,p0, w1*p1)) -- the contraction nvcc applies to the reference
//     expression w0*p0 + w1*p1 + w2*p2 -- with idx/weight loaded once per thread and reused across channels.
#include "common.cuh"
#include <math.h>

namespace de6d {

constexpr int NN_THREADS = 256;
constexpr int NN_TILE = 1024;

__global__ void __launch_bounds__(NN_THREADS)
three_nn_kernel(int n, int m, const float *__restrict__ unknown, const float *__restrict__ known,
                float *__restrict__ dist2, int *__restrict__ idx) {
    __shared__ float4 tile[NN_TILE];
    const int bs = blockIdx.y;
    const int q = blockIdx.x * NN_THREADS + threadIdx.x;
    const bool ok = q < n;
    unknown += (size_t)bs * n * 3;
    known += (size_t)bs * m * 3;
    const float ux = ok ? unknown[q * 3 + 0] : 0.f, uy = ok ? unknown[q * 3 + 1] : 0.f, uz = ok ? unknown[q * 3 + 2] : 0.f;
    float b1 = INFINITY, b2 = INFINITY, b3 = INFINITY;
    int i1 = 0, i2 = 0, i3 = 0;
    for (int base = 0; base < m; base += NN_TILE) {
        const int tn = min(NN_TILE, m - base);
        __syncthreads();
        for (int i = threadIdx.x; i < tn; i += NN_THREADS) {
            const float *p = known + (size_t)(base + i) * 3;
            tile[i] = make_float4(p[0], p[1], p[2], 0.f);
        }
        __syncthreads();
        if (ok) {
#pragma unroll 4
            for (int i = 0; i < tn; ++i) {
                const float4 p = tile[i];
                const float d = sqdist(ux, uy, uz, p.x, p.y, p.z);
                const int k = base + i;
                if (d < b1) { b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k; }
                else if (d < b2) { b3 = b2; i3 = i2; b2 = d; i2 = k; }
                else if (d < b3) { b3 = d; i3 = k; }
            }
        }
    }
    if (ok) {
        float *dd = dist2 + ((size_t)bs * n + q) * 3;
        int *ii = idx + ((size_t)bs * n + q) * 3;
        dd[0] = b1; dd[1] = b2; dd[2] = b3;
        ii[0] = i1; ii[1] = i2; ii[2] = i3;
    }
}

constexpr int TI_CPT = 8;

__global__ void __launch_bounds__(256)
three_interpolate_kernel(int c, int m, int n, const float *__restrict__ points, const int *__restrict__ idx,
                         const float *__restrict__ weight, float *__restrict__ out) {
    const int bs = blockIdx.z;
    const int c0 = blockIdx.y * TI_CPT;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const float *w = weight + ((size_t)bs * n + q) * 3;
    const int *ix = idx + ((size_t)bs * n + q) * 3;
    const float w0 = w[0], w1 = w[1], w2 = w[2];
    const int k0 = ix[0], k1 = ix[1], k2 = ix[2];
    const int ch = min(TI_CPT, c - c0);
#pragma unroll
    for (int g = 0; g < TI_CPT; ++g) {
        if (g < ch) {
            const float *p = points + ((size_t)bs * c + c0 + g) * m;
            const float v = __fmaf_rn(w2, __ldg(p + k2), __fmaf_rn(w0, __ldg(p + k0), __fmul_rn(w1, __ldg(p + k1))));
            out[((size_t)bs * c + c0 + g) * n + q] = v;
        }
    }
}

// Staged variant (the op is a 3-tap gather: HBM-bound if the taps do not each cost a 32-byte L2 sector): a CTA stages
// G channel rows of one cloud in shared memory with TMA bulk copies, then every thread produces 4 consecutive outputs
// per row from shared-memory taps -- idx / weight read as 128-bit vectors, 128-bit streaming stores.  Same FMA shape.
constexpr int TI_THREADS = 256;

__global__ void __launch_bounds__(TI_THREADS)
three_interpolate_staged_kernel(int c, int m, int n, int G, int m_pad, int chunk, const float *__restrict__ points,
                                const int *__restrict__ idx, const float *__restrict__ weight, float *__restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    float *rows = reinterpret_cast<float *>(smem_raw + 128);
    const int bs = blockIdx.z;
    const int c0 = blockIdx.y * G;
    const int g_here = min(G, c - c0);
    const float *src = points + ((size_t)bs * c + c0) * m;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, (uint32_t)g_here * (uint32_t)m * 4u);
        for (int g = 0; g < g_here; ++g) tma_bulk_g2s(rows + (size_t)g * m_pad, src + (size_t)g * m, (uint32_t)m * 4u, bar);
    }
    mbar_wait(bar, 0);
    __syncthreads();                           // every thread is past the wait: the barrier object is dead from here on
    if (threadIdx.x == 0) mbar_inval(bar);

    const int q0 = blockIdx.x * chunk, q1 = min(n, q0 + chunk);
    const int *ix = idx + (size_t)bs * n * 3;
    const float *wt = weight + (size_t)bs * n * 3;
    float *dst = out + ((size_t)bs * c + c0) * n;
    for (int q = q0 + 4 * threadIdx.x; q < q1; q += 4 * TI_THREADS) {   // n % 4 == 0 and chunk % 4 == 0 (host)
        const int4 i0 = ldg_stream_int4(ix + (size_t)q * 3), i1 = ldg_stream_int4(ix + (size_t)q * 3 + 4), i2 = ldg_stream_int4(ix + (size_t)q * 3 + 8);
        const float4 w0 = __ldg(reinterpret_cast<const float4 *>(wt + (size_t)q * 3)), w1 = __ldg(reinterpret_cast<const float4 *>(wt + (size_t)q * 3 + 4)),
                     w2 = __ldg(reinterpret_cast<const float4 *>(wt + (size_t)q * 3 + 8));
        const int k[12] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w, i2.x, i2.y, i2.z, i2.w};
        const float ww[12] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w};
        for (int g = 0; g < g_here; ++g) {
            const float *r = rows + (size_t)g * m_pad;
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                v[u] = __fmaf_rn(ww[3 * u + 2], r[k[3 * u + 2]], __fmaf_rn(ww[3 * u], r[k[3 * u]], __fmul_rn(ww[3 * u + 1], r[k[3 * u + 1]])));
            stg_stream_float4(dst + (size_t)g * n + q, make_float4(v[0], v[1], v[2], v[3]));
        }
    }
}

__global__ void __launch_bounds__(256)
three_interpolate_grad_kernel(int c, int n, int m, const float *__restrict__ grad_out, const int *__restrict__ idx,
                              const float *__restrict__ weight, float *__restrict__ grad_points) {
    const int bs = blockIdx.z;
    const int c0 = blockIdx.y * TI_CPT;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const float *w = weight + ((size_t)bs * n + q) * 3;
    const int *ix = idx + ((size_t)bs * n + q) * 3;
    const float w0 = w[0], w1 = w[1], w2 = w[2];
    const int k0 = ix[0], k1 = ix[1], k2 = ix[2];
    const int ch = min(TI_CPT, c - c0);
#pragma unroll
    for (int g = 0; g < TI_CPT; ++g) {
        if (g < ch) {
            const float go = grad_out[((size_t)bs * c + c0 + g) * n + q];
            float *gp = grad_points + ((size_t)bs * c + c0 + g) * m;
            atomicAdd(gp + k0, __fmul_rn(go, w0));
            atomicAdd(gp + k1, __fmul_rn(go, w1));
            atomicAdd(gp + k2, __fmul_rn(go, w2));
        }
    }
}

}  // namespace de6d

using namespace de6d;

int de6d_three_nn_grid(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx, void *workspace,
                       size_t workspace_bytes, cudaStream_t s);   // ball_query.cu

extern "C" size_t de6d_ball_query_workspace_bytes(int b, int n);

// impl: 0 automatic (grid search for >= 2048 known points), 1 brute-force scan, 2 grid.  workspace: as for ball query
// over `known` (de6d_ball_query_workspace_bytes(b, m)), or NULL (stream-ordered scratch).
extern "C" int de6d_three_nn_ex(int impl, int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx,
                                void *workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (b < 0 || n < 0 || m < 0) return de6d_set_error(DE6D_ERR_INVALID, "three_nn: negative size");
    if (b == 0 || n == 0) return DE6D_OK;
    if (!unknown || !dist2 || !idx || (m > 0 && !known)) return de6d_set_error(DE6D_ERR_INVALID, "three_nn: null pointer");
    if (b > 65535) return de6d_set_error(DE6D_ERR_INVALID, "three_nn: batch > 65535");
    if (m >= 3 && (impl == 2 || (impl == 0 && m >= 2048)))
        return de6d_three_nn_grid(b, n, m, unknown, known, dist2, idx, workspace, workspace_bytes, stream);
    dim3 grid(ceil_div(n, NN_THREADS), b);
    three_nn_kernel<<<grid, NN_THREADS, 0, stream>>>(n, m, unknown, known, dist2, idx);
    DE6D_CHECK_LAUNCH("three_nn_kernel");
    return DE6D_OK;
}

extern "C" int de6d_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx,
                             cudaStream_t stream) {
    return de6d_three_nn_ex(0, b, n, m, unknown, known, dist2, idx, nullptr, 0, stream);
}

extern "C" int de6d_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx,
                                      const float *weight, float *out, cudaStream_t stream) {
    if (b < 0 || c < 0 || n < 0 || m < 0) return de6d_set_error(DE6D_ERR_INVALID, "three_interpolate: negative size");
    if (b == 0 || c == 0 || n == 0) return DE6D_OK;
    if (!points || !idx || !weight || !out) return de6d_set_error(DE6D_ERR_INVALID, "three_interpolate: null pointer");
    if (b > 65535) return de6d_set_error(DE6D_ERR_INVALID, "three_interpolate: batch > 65535");
    // staged kernel: rows fit in shared memory, everything 16-byte aligned, enough outputs per staged row
    const int m_pad = (m + 3) & ~3;
    int G = (int)((64 * 1024) / ((size_t)m_pad * 4 + 1));
    if (G < 1 && (size_t)m_pad * 4 <= 200 * 1024) G = 1;
    if (G > c) G = c;
    if (G > 8) G = 8;
    const bool aligned = (n % 4 == 0) && (m % 4 == 0) && ((reinterpret_cast<uintptr_t>(points) & 15) == 0) &&
                         ((reinterpret_cast<uintptr_t>(idx) & 15) == 0) && ((reinterpret_cast<uintptr_t>(weight) & 15) == 0) &&
                         ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    if (G >= 1 && aligned && n >= m / 2) {
        const int cgroups = ceil_div(c, G);
        long long chunks = ceil_div_ll(8 * 148, (long long)b * cgroups);
        if (chunks < 1) chunks = 1;
        long long chunk = ceil_div_ll(n, chunks);
        long long min_chunk = (long long)m;          // write at least as much as is staged
        if (min_chunk < 1024) min_chunk = 1024;
        if (chunk < min_chunk) chunk = min_chunk;
        chunk = (chunk + 3) & ~3ll;
        chunks = ceil_div_ll(n, chunk);
        const size_t smem = 128 + (size_t)G * m_pad * 4;
        static unsigned long long devs = 0;
        if (smem > 48 * 1024)
            if (int rc = de6d_ensure_smem(three_interpolate_staged_kernel, 208 * 1024, devs, "three_interpolate smem attribute")) return rc;
        dim3 grid((unsigned)chunks, cgroups, b);
        three_interpolate_staged_kernel<<<grid, TI_THREADS, smem, stream>>>(c, m, n, G, m_pad, (int)chunk, points, idx, weight, out);
        DE6D_CHECK_LAUNCH("three_interpolate_staged_kernel");
        return DE6D_OK;
    }
    dim3 grid(ceil_div(n, 256), ceil_div(c, TI_CPT), b);
    three_interpolate_kernel<<<grid, 256, 0, stream>>>(c, m, n, points, idx, weight, out);
    DE6D_CHECK_LAUNCH("three_interpolate_kernel");
    return DE6D_OK;
}

extern "C" int de6d_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx,
                                           const float *weight, float *grad_points, cudaStream_t stream) {
    if (b < 0 || c < 0 || n < 0 || m < 0) return de6d_set_error(DE6D_ERR_INVALID, "three_interpolate_grad: negative size");
    if (b == 0 || c == 0 || n == 0) return DE6D_OK;
    if (!grad_out || !idx || !weight || !grad_points) return de6d_set_error(DE6D_ERR_INVALID, "three_interpolate_grad: null pointer");
    if (b > 65535) return de6d_set_error(DE6D_ERR_INVALID, "three_interpolate_grad: batch > 65535");
    dim3 grid(ceil_div(n, 256), ceil_div(c, TI_CPT), b);
    three_interpolate_grad_kernel<<<grid, 256, 0, stream>>>(c, n, m, grad_out, idx, weight, grad_points);
    DE6D_CHECK_LAUNCH("three_interpolate_grad_kernel");
    return DE6D_OK;
}
