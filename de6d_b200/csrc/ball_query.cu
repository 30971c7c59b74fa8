// Ball query (plain / counted / dilated) for sm_100a.
//
// Replaces pointnet2_batch/src/ball_query_gpu.cu:15-130 (one thread per query walking all N points serially
// out of global memory).  Output contract kept bit for bit: the first `nsample` points in ASCENDING index
// order whose squared distance (same rounded expression, common.cuh:sqdist with the query as first operand)
// passes the radius test; the three padding rules; rows of empty balls are left untouched.
//
// Design: a CTA owns QPB = NWARP*Q consecutive queries of one cloud.  The cloud streams through shared
// memory in SoA tiles (coalesced fill); each warp keeps Q queries in registers, every lane tests one point
// of the tile against all Q queries, and __ballot_sync + popc turn the 32 results into ordered slots, so
// order is preserved without any sort and a warp stops as soon as its Q balls are full.  Hit lists are
// collected in shared memory and written as whole rows (128 B for nsample = 32).
#include "common.cuh"

namespace de6d {

enum { BQ_PLAIN = 0, BQ_CNT = 1, BQ_DILATED = 2 };

constexpr int BQ_NWARP = 8;
constexpr int BQ_Q = 8;
constexpr int BQ_QPB = BQ_NWARP * BQ_Q;   // 64 queries per CTA
constexpr int BQ_TILE = 2048;             // points per shared-memory tile (24 KB)

template <int MODE>
__global__ void __launch_bounds__(BQ_NWARP * 32)
ball_query_kernel(int n, int m, float r2_in, float r2_out, int nsample, const float *__restrict__ new_xyz,
                  const float *__restrict__ xyz, int *__restrict__ idx_cnt, int *__restrict__ idx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *sx = reinterpret_cast<float *>(smem_raw);
    float *sy = sx + BQ_TILE;
    float *sz = sy + BQ_TILE;
    int *hits = reinterpret_cast<int *>(sz + BQ_TILE);  // [QPB][nsample]

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int bs = blockIdx.y;
    const int q0 = blockIdx.x * BQ_QPB + w * BQ_Q;
    xyz += (size_t)bs * n * 3;
    new_xyz += (size_t)bs * m * 3;
    int *my_hits = hits + (size_t)(w * BQ_Q) * nsample;

    float qx[BQ_Q], qy[BQ_Q], qz[BQ_Q];
    int cnt[BQ_Q];
#pragma unroll
    for (int q = 0; q < BQ_Q; ++q) {
        const int qi = q0 + q;
        const bool ok = qi < m;
        qx[q] = ok ? new_xyz[qi * 3 + 0] : 0.f;
        qy[q] = ok ? new_xyz[qi * 3 + 1] : 0.f;
        qz[q] = ok ? new_xyz[qi * 3 + 2] : 0.f;
        cnt[q] = ok ? 0 : nsample;  // out-of-range queries are born full
    }
    const unsigned lt = (1u << lane) - 1u;

    for (int base = 0; base < n; base += BQ_TILE) {
        const int tn = min(BQ_TILE, n - base);
        __syncthreads();  // previous tile fully consumed
        for (int i = tid; i < tn * 3; i += BQ_NWARP * 32) {
            float v = xyz[(size_t)base * 3 + i];
            int p = i / 3, a = i - p * 3;
            (a == 0 ? sx : a == 1 ? sy : sz)[p] = v;
        }
        __syncthreads();
        int minc = cnt[0];
#pragma unroll
        for (int q = 1; q < BQ_Q; ++q) minc = min(minc, cnt[q]);
        bool warp_done = minc >= nsample;
        if (!warp_done) {
            for (int i0 = 0; i0 < tn; i0 += 32) {
                const int i = i0 + lane;
                const bool valid = i < tn;
                const float x = valid ? sx[i] : 0.f, y = valid ? sy[i] : 0.f, z = valid ? sz[i] : 0.f;
                int mc = nsample;
#pragma unroll
                for (int q = 0; q < BQ_Q; ++q) {
                    const float d2 = sqdist(qx[q], qy[q], qz[q], x, y, z);
                    bool hit = valid && (d2 < r2_out);
                    if (MODE == BQ_DILATED) hit = hit && (d2 >= r2_in);
                    const unsigned bal = __ballot_sync(0xffffffffu, hit);
                    if (bal) {
                        const int slot = cnt[q] + __popc(bal & lt);
                        if (hit && slot < nsample) my_hits[q * nsample + slot] = base + i;
                        cnt[q] += __popc(bal);
                    }
                    mc = min(mc, cnt[q]);
                }
                if (mc >= nsample) break;
            }
        }
        // stop streaming tiles once every warp of the CTA is full
        int mc2 = cnt[0];
#pragma unroll
        for (int q = 1; q < BQ_Q; ++q) mc2 = min(mc2, cnt[q]);
        if (__syncthreads_and(mc2 >= nsample)) break;
    }
    __syncwarp();

    // write rows: hits then padding
#pragma unroll
    for (int q = 0; q < BQ_Q; ++q) {
        const int qi = q0 + q;
        if (qi >= m) continue;
        const int c = min(cnt[q], nsample);
        int *row = idx + ((size_t)bs * m + qi) * nsample;
        if (MODE != BQ_PLAIN && lane == 0) idx_cnt[(size_t)bs * m + qi] = c;
        if (c == 0) continue;  // reference leaves the (pre-zeroed) row untouched
        for (int s = lane; s < nsample; s += 32) {
            int v;
            if (s < c) v = my_hits[q * nsample + s];
            else v = (MODE == BQ_PLAIN) ? my_hits[q * nsample] : my_hits[q * nsample + (s % c)];
            row[s] = v;
        }
    }
}

template <int MODE>
static int launch_ball_query(int b, int n, int m, float r_in, float r_out, int nsample, const float *new_xyz,
                             const float *xyz, int *idx_cnt, int *idx, cudaStream_t s) {
    if (b < 0 || n < 0 || m < 0 || nsample < 0) return de6d_set_error(DE6D_ERR_INVALID, "ball_query: negative size");
    if (b == 0 || m == 0) return DE6D_OK;
    if (!new_xyz || !idx || (n > 0 && !xyz) || (MODE != BQ_PLAIN && !idx_cnt))
        return de6d_set_error(DE6D_ERR_INVALID, "ball_query: null pointer");
    if (b > 65535) return de6d_set_error(DE6D_ERR_INVALID, "ball_query: batch > 65535");
    size_t smem = (size_t)BQ_TILE * 12 + (size_t)BQ_QPB * (nsample > 0 ? nsample : 1) * sizeof(int);
    if (smem > 200 * 1024) return de6d_set_error(DE6D_ERR_INVALID, "ball_query: nsample too large");
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(ball_query_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return de6d_set_cuda_error(e, "ball_query smem attribute");
        configured = 200 * 1024;
    }
    // radius*radius in float, as the reference kernels compute it (ball_query_gpu.cu:29,74-75,114)
    const float r2_in = r_in * r_in, r2_out = r_out * r_out;
    dim3 grid(ceil_div(m, BQ_QPB), b);
    ball_query_kernel<MODE><<<grid, BQ_NWARP * 32, smem, s>>>(n, m, r2_in, r2_out, nsample, new_xyz, xyz, idx_cnt, idx);
    DE6D_CHECK_LAUNCH("ball_query_kernel");
    return DE6D_OK;
}

}  // namespace de6d

using namespace de6d;

extern "C" int de6d_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz,
                               int *idx, cudaStream_t stream) {
    return launch_ball_query<BQ_PLAIN>(b, n, m, 0.f, radius, nsample, new_xyz, xyz, nullptr, idx, stream);
}
extern "C" int de6d_ball_query_cnt(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                                   const float *xyz, int *idx_cnt, int *idx, cudaStream_t stream) {
    return launch_ball_query<BQ_CNT>(b, n, m, 0.f, radius, nsample, new_xyz, xyz, idx_cnt, idx, stream);
}
extern "C" int de6d_ball_query_dilated(int b, int n, int m, float radius_in, float radius_out, int nsample,
                                       const float *new_xyz, const float *xyz, int *idx_cnt, int *idx,
                                       cudaStream_t stream) {
    return launch_ball_query<BQ_DILATED>(b, n, m, radius_in, radius_out, nsample, new_xyz, xyz, idx_cnt, idx, stream);
}
