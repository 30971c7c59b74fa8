// group_points / gather_points (+ their gradients) for sm_100a.
//
// Replaces pointnet2_batch/src/group_points_gpu.cu:14-72 and sampling_gpu.cu:16-71 (one thread per output
// element, the index tensor re-read once per channel, 4-byte stores).  Pure copies, so bit-exact by
// construction; the work is moving bytes:
//   out[b, c, j] = points[b, c, idx[b, j]],   j over npoints*nsample (gather_points is the nsample = 1 case)
//
// Two kernels, picked per call by the host:
//   * staged: a CTA stages G whole channel rows of one cloud (G*N*4 bytes, up to ~192 KB) in shared memory
//     with one TMA bulk copy per row (cp.async.bulk + mbarrier), then streams its slice of the index tensor
//     once (128-bit loads) and emits G output rows with 128-bit stores.  Random 4-byte reads hit shared
//     memory instead of 32-byte L2 sectors, and idx is read once per G channels instead of once per channel.
//   * direct: few outputs per source row (e.g. gathering 4096 of 16384 points) or rows too large for shared
//     memory -- 128-bit index loads, read-only-path gathers, 128-bit stores, channel loop inside the thread.
#include "common.cuh"

namespace de6d {

constexpr int GS_THREADS = 512;
// tuning knobs (scripts/group_variants.py builds the library with other values and times them on the GPU)
#ifndef DE6D_GS_U
#define DE6D_GS_U 2        // index vectors in flight per thread (r2 A/B on B200: 2 with 3 CTAs/SM is best or within 1 %)
#endif
#ifndef DE6D_GS_MINB
#define DE6D_GS_MINB 3     // CTAs per SM the register allocation must allow (512-thread CTAs with <= 64 KB of rows)
#endif
constexpr int GS_THREADS_BIG = 1024;   // one CTA per SM holding > 64 KB of rows (large clouds, few channels)

// Leading coordinate rows of the fused grouper tail (de6d_group_concat): virtual rows 0..2 of the output are
// xyz[idx] - new_xyz, rows 3.. are the feature channels.  A coordinate row is staged from the transposed cloud xyz_t
// (b,3,n) by TMA when the caller has one, else from the AoS cloud with strided loads (served by L2: clouds are small).
struct GroupXyz {
    const float *xyz;       // (b, n, 3)
    const float *xyz_t;     // (b, 3, n) or nullptr
    const float *new_xyz;   // (b, m, 3) ball centres
    int ns;                 // nsample: output slot j belongs to centre j / ns
    int m;
};

// grid: (chunks, ceil(C/G), B).  Dynamic smem: G*n_pad floats + one mbarrier.  XYZ: the first 3 of the `c` rows are the
// coordinate rows described above and `points` holds the remaining c - 3 channels (ns % 4 == 0 required).
template <bool TMA, bool XYZ, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
group_staged_kernel(int c, int n, long long ms, int G, int n_pad, long long chunk, const float *__restrict__ points,
                    const int *__restrict__ idx, float *__restrict__ out, long long out_bstride, GroupXyz gx) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    float *rows = reinterpret_cast<float *>(smem_raw + 128);

    const int bs = blockIdx.z;
    const int c0 = blockIdx.y * G;
    const int g_here = min(G, c - c0);
    const int nx = XYZ ? 3 : 0;
    const float *feat = points + (size_t)bs * (c - nx) * n;   // channel v >= nx lives at feat + (v - nx) * n

    // where virtual row v comes from: TMA-able contiguous row, or nullptr = strided from the AoS cloud
    auto row_src = [&](int v) -> const float * {
        if (!XYZ || v >= nx) return feat + (size_t)(v - nx) * n;
        return gx.xyz_t ? gx.xyz_t + ((size_t)bs * 3 + v) * n : nullptr;
    };

    if (TMA) {
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int nt = 0;
            for (int g = 0; g < g_here; ++g) nt += row_src(c0 + g) != nullptr;
            mbar_expect_tx(bar, (uint32_t)nt * (uint32_t)n * 4u);
            for (int g = 0; g < g_here; ++g) {
                const float *src = row_src(c0 + g);
                if (src) tma_bulk_g2s(rows + (size_t)g * n_pad, src, (uint32_t)n * 4u, bar);
            }
        }
    }
    bool plain = false;   // rows this CTA copies with ordinary loads
    for (int g = 0; g < g_here; ++g) {
        const float *src = row_src(c0 + g);
        if (TMA && src) continue;
        plain = true;
        float *dstrow = rows + (size_t)g * n_pad;
        if (src) {
            for (int i = threadIdx.x; i < n; i += THREADS) dstrow[i] = src[i];
        } else {
            const float *a = gx.xyz + (size_t)bs * n * 3 + (c0 + g);
            for (int i = threadIdx.x; i < n; i += THREADS) dstrow[i] = __ldg(a + (size_t)i * 3);
        }
    }
    if (plain || !TMA) __syncthreads();   // `plain` is uniform over the CTA
    if (TMA) {
        mbar_wait(bar, 0);
        __syncthreads();                       // every thread is past the wait: the barrier object is dead from here on
        if (threadIdx.x == 0) mbar_inval(bar);
    }

    const long long j0 = (long long)blockIdx.x * chunk;
    const long long j1 = min(ms, j0 + chunk);
    const int *ix = idx + (size_t)bs * ms;
    float *dst = out + (size_t)bs * out_bstride + (size_t)c0 * ms;
    const float *ctr = XYZ ? gx.new_xyz + (size_t)bs * gx.m * 3 : nullptr;
    // chunk is a multiple of 4 and ms % 4 == 0 is checked by the host for this kernel.  U index vectors are in flight per
    // thread: with few rows per CTA (layer 1: one 64 KB row) the index stream is as large as the output stream and one
    // 16-byte load per thread at a time cannot cover its latency (measured: 0.60 -> see profiles/r2 of the HBM peak).
    constexpr int U = (THREADS == GS_THREADS_BIG) ? 4 : DE6D_GS_U;
    constexpr long long STEP = 4ll * THREADS;
    for (long long j = j0 + 4ll * threadIdx.x; j < j1; j += U * STEP) {
        int4 kk[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (j + u * STEP < j1) kk[u] = ldg_stream_int4(ix + j + u * STEP);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long ju = j + u * STEP;
            if (ju >= j1) break;
            const int4 k = kk[u];
            const unsigned q = XYZ ? (unsigned)ju / (unsigned)gx.ns : 0u;   // ms < 2^31 checked by the host; 4 slots share a centre
            for (int g = 0; g < g_here; ++g) {
                const float *r = rows + (size_t)g * n_pad;
                float4 v = make_float4(r[k.x], r[k.y], r[k.z], r[k.w]);
                if (XYZ) {   // branch-free: channel rows subtract +0.0f, which leaves every float as it is
                    const float cv = (c0 + g < nx) ? __ldg(ctr + (size_t)q * 3 + (c0 + g)) : 0.f;
                    v.x = __fsub_rn(v.x, cv); v.y = __fsub_rn(v.y, cv); v.z = __fsub_rn(v.z, cv); v.w = __fsub_rn(v.w, cv);
                }
                stg_stream_float4(dst + (size_t)g * ms + ju, v);
            }
        }
    }
}

// direct gathers; CPT channels per thread (idx held in registers across them).  grid: (x, ceil(C/CPT), B)
template <int CPT>
__global__ void __launch_bounds__(256)
group_direct_kernel(int c, int n, long long ms, int vec, const float *__restrict__ points, const int *__restrict__ idx,
                    float *__restrict__ out, long long out_bstride) {
    const int bs = blockIdx.z;
    const int c0 = blockIdx.y * CPT;
    const int *ix = idx + (size_t)bs * ms;
    const float *src = points + ((size_t)bs * c + c0) * n;
    float *dst = out + (size_t)bs * out_bstride + (size_t)c0 * ms;
    const int ch = min(CPT, c - c0);
    const long long stride = (long long)gridDim.x * blockDim.x;
    if (vec) {
        for (long long j4 = (long long)blockIdx.x * blockDim.x + threadIdx.x; j4 * 4 < ms; j4 += stride) {
            const long long j = j4 * 4;
            const int4 k = *reinterpret_cast<const int4 *>(ix + j);
#pragma unroll
            for (int g = 0; g < CPT; ++g) {
                if (g < ch) {
                    const float *r = src + (size_t)g * n;
                    float4 v = make_float4(__ldg(r + k.x), __ldg(r + k.y), __ldg(r + k.z), __ldg(r + k.w));
                    *reinterpret_cast<float4 *>(dst + (size_t)g * ms + j) = v;
                }
            }
        }
    } else {
        for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < ms; j += stride) {
            const int k = ix[j];
#pragma unroll
            for (int g = 0; g < CPT; ++g)
                if (g < ch) dst[(size_t)g * ms + j] = __ldg(src + (size_t)g * n + k);
        }
    }
}

// gradient: grad_points[b, c, idx[b, j]] += grad_out[b, c, j]   (group_points_gpu.cu:14-31, sampling_gpu.cu:54-71)
template <int CPT>
__global__ void __launch_bounds__(256)
group_grad_kernel(int c, int n, long long ms, const float *__restrict__ grad_out, const int *__restrict__ idx,
                  float *__restrict__ grad_points) {
    const int bs = blockIdx.z;
    const int c0 = blockIdx.y * CPT;
    const int ch = min(CPT, c - c0);
    const int *ix = idx + (size_t)bs * ms;
    const float *g_out = grad_out + ((size_t)bs * c + c0) * ms;
    float *g_pts = grad_points + ((size_t)bs * c + c0) * n;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < ms; j += stride) {
        const int k = ix[j];
#pragma unroll
        for (int g = 0; g < CPT; ++g)
            if (g < ch) atomicAdd(g_pts + (size_t)g * n + k, g_out[(size_t)g * ms + j]);
    }
}

// Relative coordinates of the grouped points: out[b, a, j] = xyz[b, idx[b, j], a] - new_xyz[b, j / ns, a], a < 3
// (pointnet2_utils.py:410-412: grouping_operation on the transposed xyz, then the in-place centre subtraction; the
// transposed copy of xyz and the separate subtraction pass disappear).  grid (x, 1, B).
__global__ void __launch_bounds__(256)
group_xyz_center_kernel(int n, int m, int ns, int vec, const float *__restrict__ xyz, const float *__restrict__ new_xyz,
                        const int *__restrict__ idx, float *__restrict__ out, long long out_bstride) {
    const int bs = blockIdx.z;
    const long long ms = (long long)m * ns;
    const int *ix = idx + (size_t)bs * ms;
    const float *src = xyz + (size_t)bs * n * 3;
    const float *ctr = new_xyz + (size_t)bs * m * 3;
    float *dst = out + (size_t)bs * out_bstride;
    const long long stride = (long long)gridDim.x * blockDim.x;
    if (vec) {   // ns % 4 == 0: four consecutive slots share their centre
        for (long long j4 = (long long)blockIdx.x * blockDim.x + threadIdx.x; j4 * 4 < ms; j4 += stride) {
            const long long j = j4 * 4;
            const int4 k = ldg_stream_int4(ix + j);
            const float *cq = ctr + (j / ns) * 3;
            const int kk[4] = {k.x, k.y, k.z, k.w};
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float cv = __ldg(cq + a);
                float4 v;
                v.x = __fsub_rn(__ldg(src + (size_t)kk[0] * 3 + a), cv);
                v.y = __fsub_rn(__ldg(src + (size_t)kk[1] * 3 + a), cv);
                v.z = __fsub_rn(__ldg(src + (size_t)kk[2] * 3 + a), cv);
                v.w = __fsub_rn(__ldg(src + (size_t)kk[3] * 3 + a), cv);
                stg_stream_float4(dst + (size_t)a * ms + j, v);
            }
        }
    } else {
        for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < ms; j += stride) {
            const int k = ix[j];
            const float *cq = ctr + (j / ns) * 3;
#pragma unroll
            for (int a = 0; a < 3; ++a) dst[(size_t)a * ms + j] = __fsub_rn(__ldg(src + (size_t)k * 3 + a), __ldg(cq + a));
        }
    }
}

// c counts the 3 coordinate rows when gx != nullptr (points then holds c - 3 channels); returns DE6D_OK and sets
// *launched = false when the staged kernel is not applicable to the coordinate rows (caller takes the two-kernel path).
static int group_forward(int b, int c, int n, long long ms, const float *points, const int *idx, float *out,
                         long long out_bstride, int force_impl, cudaStream_t s, const GroupXyz *gx = nullptr,
                         bool *launched = nullptr) {
    if (launched) *launched = false;
    if (b < 0 || c < 0 || n < 0 || ms < 0) return de6d_set_error(DE6D_ERR_INVALID, "group/gather: negative size");
    if (b == 0 || c == 0 || ms == 0) return DE6D_OK;
    if ((!points && !(gx && c == 3)) || !idx || !out) return de6d_set_error(DE6D_ERR_INVALID, "group/gather: null pointer");
    if (b > 65535) return de6d_set_error(DE6D_ERR_INVALID, "group/gather: batch > 65535");

    // staged kernel: rows must fit, outputs per row must amortise the staging, 16-byte alignment everywhere
    const size_t smem_budget = 200 * 1024;
    const int n_pad = (n + 3) & ~3;
    const bool aligned = (ms % 4 == 0) && ((reinterpret_cast<uintptr_t>(idx) & 15) == 0) &&
                         ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    bool tma_ok = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(points) & 15) == 0);
    if (gx && gx->xyz_t) tma_ok = tma_ok && ((reinterpret_cast<uintptr_t>(gx->xyz_t) & 15) == 0);
    // channel rows per CTA: as many as fit in ~64 KB (three CTAs per SM hide the index-load latency better than one
    // CTA with 192 KB: measured +3..10 % on B200, scripts/group_tune.py), at most 8, at least one if a row fits at all
    int G = (int)(smem_budget / ((size_t)n_pad * 4 + 1));
    const int Gmax = G;
    const int G64 = (int)((64 * 1024) / ((size_t)n_pad * 4 + 1));
    if (G > 1 && G64 >= 1 && G > G64) G = G64;
    if (G > 1 && G64 < 1) G = 1;
    if (G > c) G = c;
    if (G > 8) G = 8;
    // Large clouds with few channels (layer 1: 16384 points, 3 + 1 rows of 64 KB): with one row per CTA every CTA re-reads
    // its whole index chunk for a single output row, so the SM's read traffic (row + indices) is 1.5x its write traffic
    // and the kernel stops at ~0.6 of the HBM peak (r2 measurement).  Two or three rows per CTA (one 1024-thread CTA per
    // SM, up to 200 KB of rows) halve the index re-reads.
    const bool big = (G64 < 2) && Gmax >= 2 && c >= 2;
    if (big) {
        const int groups = ceil_div(c, Gmax);
        G = ceil_div(c, groups);
    }
    bool staged = aligned && G >= 1 && ms >= 2ll * n;
    if (gx) {
        if (!(staged && gx->ns % 4 == 0 && ms < (1ll << 31))) return DE6D_OK;   // not launched: caller falls back
    } else {
        if (force_impl == 1) staged = false;
        if (force_impl == 2 && !(aligned && G >= 1)) return de6d_set_error(DE6D_ERR_INVALID, "group: staged kernel not applicable");
        if (force_impl == 2) staged = true;
    }

    if (staged) {
        const int cgroups = ceil_div(c, G);
        long long chunk, chunks;
        if (big) {
            // chunks per row group: minimise waves x per-CTA traffic, traffic = max(reads, writes) of one CTA
            const long long slots = 148;
            double best = 1e300;
            chunks = 1;
            for (long long k = 1; k <= 16; ++k) {
                long long ch = (ceil_div_ll(ms, k) + 3) & ~3ll;
                if (k > 1 && ch < (long long)n) break;              // never stage more than is written per row
                const long long ctas = (long long)b * cgroups * ceil_div_ll(ms, ch);
                const double reads = (double)G * n * 4 + (double)ch * 4, writes = (double)G * ch * 4;
                const double t = (double)ceil_div_ll(ctas, slots) * (reads > writes ? reads : writes);
                if (t < best * 0.999) { best = t; chunks = k; }
            }
            chunk = (ceil_div_ll(ms, chunks) + 3) & ~3ll;
            chunks = ceil_div_ll(ms, chunk);
        } else {
            // enough CTAs to fill the machine, but each CTA should write >= ~2x what it stages
            long long min_chunk = (long long)n * 2;
            if (min_chunk < 4096) min_chunk = 4096;
            long long want = ceil_div_ll(8 * 148, (long long)b * cgroups);   // ~8 waves of CTAs: small tail
            chunks = want < 1 ? 1 : want;
            chunk = ceil_div_ll(ms, chunks);
            if (chunk < min_chunk) chunk = min_chunk;
            chunk = (chunk + 3) & ~3ll;
            chunks = ceil_div_ll(ms, chunk);
        }
        size_t smem = 128 + (size_t)G * n_pad * 4;
        static unsigned long long devs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        dim3 grid((unsigned)chunks, cgroups, b);
        const GroupXyz none = {nullptr, nullptr, nullptr, 4, 0};
#define DE6D_GROUP_LAUNCH(T, X, TH, MB, slot)                                                                                   \
    do {                                                                                                                       \
        if (smem > 48 * 1024)                                                                                                  \
            if (int rc = de6d_ensure_smem(group_staged_kernel<T, X, TH, MB>, 208 * 1024, devs[slot], "group smem attribute")) return rc; \
        group_staged_kernel<T, X, TH, MB><<<grid, TH, smem, s>>>(c, n, ms, G, n_pad, chunk, points, idx, out, out_bstride,     \
                                                                  gx ? *gx : none);                                            \
    } while (0)
#define DE6D_GROUP_PICK(TH, MB, base)                                                                     \
    do {                                                                                                  \
        if (gx) { if (tma_ok) DE6D_GROUP_LAUNCH(true, true, TH, MB, base + 3); else DE6D_GROUP_LAUNCH(false, true, TH, MB, base + 2); } \
        else { if (tma_ok) DE6D_GROUP_LAUNCH(true, false, TH, MB, base + 1); else DE6D_GROUP_LAUNCH(false, false, TH, MB, base + 0); }  \
    } while (0)
        if (big) DE6D_GROUP_PICK(GS_THREADS_BIG, 1, 4);
        else DE6D_GROUP_PICK(GS_THREADS, DE6D_GS_MINB, 0);
#undef DE6D_GROUP_PICK
#undef DE6D_GROUP_LAUNCH
        DE6D_CHECK_LAUNCH("group_staged_kernel");
        if (launched) *launched = true;
        return DE6D_OK;
    }
    constexpr int CPT = 4;
    const int vec = aligned ? 1 : 0;
    long long work = vec ? ms / 4 : ms;
    long long bx = ceil_div_ll(work, 256);
    if (bx > 4096) bx = 4096;
    dim3 grid((unsigned)bx, ceil_div(c, CPT), b);
    group_direct_kernel<CPT><<<grid, 256, 0, s>>>(c, n, ms, vec, points, idx, out, out_bstride);
    DE6D_CHECK_LAUNCH("group_direct_kernel");
    return DE6D_OK;
}

static int group_backward(int b, int c, int n, long long ms, const float *grad_out, const int *idx, float *grad_points,
                          cudaStream_t s) {
    if (b < 0 || c < 0 || n < 0 || ms < 0) return de6d_set_error(DE6D_ERR_INVALID, "group/gather grad: negative size");
    if (b == 0 || c == 0 || ms == 0) return DE6D_OK;
    if (!grad_out || !idx || !grad_points) return de6d_set_error(DE6D_ERR_INVALID, "group/gather grad: null pointer");
    if (b > 65535) return de6d_set_error(DE6D_ERR_INVALID, "group/gather grad: batch > 65535");
    constexpr int CPT = 4;
    long long bx = ceil_div_ll(ms, 256);
    if (bx > 4096) bx = 4096;
    dim3 grid((unsigned)bx, ceil_div(c, CPT), b);
    group_grad_kernel<CPT><<<grid, 256, 0, s>>>(c, n, ms, grad_out, idx, grad_points);
    DE6D_CHECK_LAUNCH("group_grad_kernel");
    return DE6D_OK;
}

}  // namespace de6d

using namespace de6d;

extern "C" int de6d_group_points(int b, int c, int n, int npoints, int nsample, const float *points, const int *idx,
                                 float *out, cudaStream_t stream) {
    return group_forward(b, c, n, (long long)npoints * nsample, points, idx, out, (long long)c * npoints * nsample, 0, stream);
}
// impl: 0 auto, 1 direct-gather kernel, 2 shared-memory staged (TMA) kernel -- identical results
extern "C" int de6d_group_points_impl(int b, int c, int n, int npoints, int nsample, const float *points,
                                      const int *idx, float *out, int impl, cudaStream_t stream) {
    return group_forward(b, c, n, (long long)npoints * nsample, points, idx, out, (long long)c * npoints * nsample, impl, stream);
}
extern "C" int de6d_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                                      const int *idx, float *grad_points, cudaStream_t stream) {
    return group_backward(b, c, n, (long long)npoints * nsample, grad_out, idx, grad_points, stream);
}
extern "C" int de6d_gather_points(int b, int c, int n, int npoints, const float *points, const int *idx, float *out,
                                  cudaStream_t stream) {
    return group_forward(b, c, n, (long long)npoints, points, idx, out, (long long)c * npoints, 0, stream);
}
extern "C" int de6d_gather_points_grad(int b, int c, int n, int npoints, const float *grad_out, const int *idx,
                                       float *grad_points, cudaStream_t stream) {
    return group_backward(b, c, n, (long long)npoints, grad_out, idx, grad_points, stream);
}

// Fused tail of QueryAndGroup / QueryWithCntAndGroup / QueryAndGroupDilated (pointnet2_utils.py:368-387,410-424,449-463)
// for use_xyz = True: out (b, 3 + c, npoints, nsample) = cat(xyz[idx] - new_xyz, features[idx]) written once, instead of
// transpose + group + in-place subtract + group + torch.cat (three extra passes over the largest tensor of the network).
// xyz (b,n,3), new_xyz (b,npoints,3), features (b,c,n) or NULL with c == 0, idx (b,npoints,nsample).
// xyz_t: optional transposed copy of xyz, (b,3,n) (the reference module keeps one as `xyz_flipped`,
// pointnet2_modules.py:374): coordinate rows are then staged by TMA like feature rows; NULL = staged from xyz itself.
// One launch (coordinates and channels are rows of the same shared-memory staged kernel) whenever nsample % 4 == 0 and the
// output is 16-byte aligned; other shapes take a coordinate kernel + the channel kernel.
extern "C" int de6d_group_concat_t(int b, int c, int n, int npoints, int nsample, const float *xyz, const float *xyz_t,
                                   const float *new_xyz, const float *features, const int *idx, float *out, cudaStream_t stream) {
    if (b < 0 || c < 0 || n < 0 || npoints < 0 || nsample < 0) return de6d_set_error(DE6D_ERR_INVALID, "group_concat: negative size");
    const long long ms = (long long)npoints * nsample;
    if (b == 0 || ms == 0) return DE6D_OK;
    if (!xyz || !new_xyz || !idx || !out || (c > 0 && !features)) return de6d_set_error(DE6D_ERR_INVALID, "group_concat: null pointer");
    if (b > 65535) return de6d_set_error(DE6D_ERR_INVALID, "group_concat: batch > 65535");
    const long long bstride = (long long)(3 + c) * ms;
    {
        const GroupXyz gx = {xyz, xyz_t, new_xyz, nsample, npoints};
        bool launched = false;
        if (int rc = group_forward(b, 3 + c, n, ms, features, idx, out, bstride, 0, stream, &gx, &launched)) return rc;
        if (launched) return DE6D_OK;
    }
    const int vec = (nsample % 4 == 0) && ((reinterpret_cast<uintptr_t>(idx) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    long long work = vec ? ms / 4 : ms;
    long long bx = ceil_div_ll(work, 256);
    if (bx > 2048) bx = 2048;
    dim3 grid((unsigned)bx, 1, b);
    group_xyz_center_kernel<<<grid, 256, 0, stream>>>(n, npoints, nsample, vec, xyz, new_xyz, idx, out, bstride);
    DE6D_CHECK_LAUNCH("group_xyz_center_kernel");
    if (c > 0) return group_forward(b, c, n, ms, features, idx, out + 3 * ms, bstride, 0, stream);
    return DE6D_OK;
}
extern "C" int de6d_group_concat(int b, int c, int n, int npoints, int nsample, const float *xyz, const float *new_xyz,
                                 const float *features, const int *idx, float *out, cudaStream_t stream) {
    return de6d_group_concat_t(b, c, n, npoints, nsample, xyz, nullptr, new_xyz, features, idx, out, stream);
}

// new_xyz = xyz[sample_idx] in both layouts with one launch: replaces transpose(1,2).contiguous() -> gather_operation ->
// transpose(1,2).contiguous() of the SA module (pointnet2_modules.py:374,451-454).  xyz (b,n,3), idx (b,m) or NULL
// (identity: m == n, plain transposition), new_xyz (b,m,3) and/or new_xyz_t (b,3,m), either may be NULL.
namespace de6d {
__global__ void __launch_bounds__(256)
gather_xyz_kernel(int n, int m, const float *__restrict__ xyz, const int *__restrict__ idx, float *__restrict__ new_xyz,
                  float *__restrict__ new_xyz_t) {
    const int bs = blockIdx.y;
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= m) return;
    const int k = idx ? idx[(size_t)bs * m + j] : j;
    const float *p = xyz + ((size_t)bs * n + k) * 3;
    const float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
    if (new_xyz) {
        float *o = new_xyz + ((size_t)bs * m + j) * 3;
        o[0] = x; o[1] = y; o[2] = z;
    }
    if (new_xyz_t) {
        float *o = new_xyz_t + (size_t)bs * 3 * m + j;
        o[0] = x; o[m] = y; o[2 * (size_t)m] = z;
    }
}
}  // namespace de6d
extern "C" int de6d_gather_xyz(int b, int n, int m, const float *xyz, const int *idx, float *new_xyz, float *new_xyz_t,
                               cudaStream_t stream) {
    if (b < 0 || n < 0 || m < 0) return de6d_set_error(DE6D_ERR_INVALID, "gather_xyz: negative size");
    if (b == 0 || m == 0) return DE6D_OK;
    if (!xyz || (!new_xyz && !new_xyz_t)) return de6d_set_error(DE6D_ERR_INVALID, "gather_xyz: null pointer");
    if (!idx && m != n) return de6d_set_error(DE6D_ERR_INVALID, "gather_xyz: idx == NULL needs m == n");
    if (b > 65535) return de6d_set_error(DE6D_ERR_INVALID, "gather_xyz: batch > 65535");
    dim3 grid(ceil_div(m, 256), b);
    gather_xyz_kernel<<<grid, 256, 0, stream>>>(n, m, xyz, idx, new_xyz, new_xyz_t);
    DE6D_CHECK_LAUNCH("gather_xyz_kernel");
    return DE6D_OK;
}
