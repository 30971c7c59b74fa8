// F-FPS distance matrix for sm_100a: out[b,i,j] = |xyz_i - xyz_j| + gamma * |feat_i - feat_j|.
//
// Replaces calc_dist_matrix_for_sampling (pointnet2_utils.py:36-44), which the reference evaluates with two
// torch.cdist calls (each a padded SGEMM + clamp + sqrt pass over the (B,N,N) matrix) plus a scale and an add:
// six full passes over a 64 MB/frame tensor at N = 4096.  Here one kernel produces the final matrix with a single
// write per element, and only tiles on or above the diagonal are computed -- the value is bitwise symmetric
// because (a-b)^2 == (b-a)^2 in IEEE arithmetic -- each tile being stored twice (as is and transposed, both with
// 128-bit stores).
//
// Arithmetic (restated exactly by oracle/de6d_oracle.c:orc_dist_matrix):
//   d1  = sqrtf(fmaf(dz,dz, fmaf(dx,dx, dy*dy)))                       coordinates, same shape as common.cuh:sqdist
//   acc = 0;  for ch = 0..C-1: t = f_i[ch] - f_j[ch]; acc = fmaf(t, t, acc)
//   out = d1 + gamma * sqrtf(acc)                                      separately rounded multiply and add
// Direct differences instead of the |a|^2+|b|^2-2ab expansion torch uses: no cancellation for close points (the
// ones FPS ranks), at the price of one extra FADD per channel.
#include "common.cuh"
#include <math.h>

namespace de6d {

constexpr int DM_TILE = 128;    // outputs per CTA: 128 x 128, 8 x 8 per thread (as 2 x 2 blocks of 4 x 4)
constexpr int DM_CH = 16;       // channels staged per pass
constexpr int DM_PITCH = DM_TILE + 4;   // keeps float4 rows 16-byte aligned

__global__ void __launch_bounds__(256)
dist_matrix_kernel(int n, int c, const float *__restrict__ xyz_all, const float *__restrict__ feat_all, long long fsb,
                   long long fsn, long long fsc, float gamma, float *__restrict__ out_all, int vec_ok) {
    __shared__ __align__(16) float fs[2][DM_CH][DM_PITCH];   // [0] rows-tile features, [1] columns-tile features
    __shared__ float xs[2][3][DM_TILE];
    float(*fa)[DM_PITCH] = fs[0];
    float(*fb)[DM_PITCH] = fs[1];

    const int tid = threadIdx.x;
    const int T = ceil_div(n, DM_TILE);
    // linear index over tiles with ti <= tj
    int ti = 0, rem = blockIdx.x;
    while (rem >= T - ti) { rem -= T - ti; ++ti; }
    const int tj = ti + rem;
    const int i0 = ti * DM_TILE, j0 = tj * DM_TILE;
    const int bs = blockIdx.y;
    const float *xyz = xyz_all + (size_t)bs * n * 3;
    const float *feat = feat_all ? feat_all + (long long)bs * fsb : nullptr;
    float *out = out_all + (size_t)bs * n * n;

    for (int e = tid; e < 2 * 3 * DM_TILE; e += 256) {
        const int which = e / (3 * DM_TILE), r = e - which * 3 * DM_TILE;
        const int p = r / 3, a = r - p * 3;
        const int gp = (which ? j0 : i0) + p;
        xs[which][a][p] = gp < n ? xyz[(size_t)gp * 3 + a] : 0.f;
    }

    // thread (ty, tx) owns rows {4ty..4ty+3, 64+4ty..} x columns {4tx..4tx+3, 64+4tx..}: every shared-memory read
    // is a conflict-free 128-bit load, 4 loads feed 128 arithmetic instructions
    const int ty = tid >> 4, tx = tid & 15;
    // accumulators as float2 pairs along the column index: the packed-fp32 instructions of sm_100 (FADD2 / FFMA2,
    // IEEE round-to-nearest per lane, so bit-identical to the scalar form) halve the issue slots of the inner loop.
    // The column tile is stored NEGATED in shared memory so that a - b is the single packed add a + (-b).
    float2 acc[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[r][q] = make_float2(0.f, 0.f);

    if (feat) {
        const bool point_major_fast = fsn <= fsc;   // which index is contiguous in memory: points or channels
        for (int ch0 = 0; ch0 < c; ch0 += DM_CH) {
            __syncthreads();   // previous pass fully consumed
            for (int e = tid; e < 2 * DM_CH * DM_TILE; e += 256) {
                const int which = e / (DM_CH * DM_TILE), r = e - which * DM_CH * DM_TILE;
                int p, ch;
                if (point_major_fast) { ch = r / DM_TILE; p = r - ch * DM_TILE; }
                else { p = r / DM_CH; ch = r - p * DM_CH; }
                const int gp = (which ? j0 : i0) + p, gc = ch0 + ch;
                const float v = (gp < n && gc < c) ? __ldg(feat + (long long)gp * fsn + (long long)gc * fsc) : 0.f;
                if (which) fb[ch][p] = -v; else fa[ch][p] = v;
            }
            __syncthreads();
            const int lim = min(DM_CH, c - ch0);
#pragma unroll 4
            for (int ch = 0; ch < lim; ++ch) {
                const float4 a0 = *reinterpret_cast<const float4 *>(&fa[ch][ty * 4]);
                const float4 a1 = *reinterpret_cast<const float4 *>(&fa[ch][64 + ty * 4]);
                const float4 b0 = *reinterpret_cast<const float4 *>(&fb[ch][tx * 4]);
                const float4 b1 = *reinterpret_cast<const float4 *>(&fb[ch][64 + tx * 4]);
                const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float2 nb[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const float2 a2 = make_float2(av[r], av[r]);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float2 t = __fadd2_rn(a2, nb[q]);
                        acc[r][q] = __ffma2_rn(t, t, acc[r][q]);
                    }
                }
            }
        }
    }
    __syncthreads();   // xs visible (no-feature case)

    // ---- epilogue: coordinates term, sqrt, combine; write the tile and (off the diagonal) its transpose ----
#pragma unroll
    for (int rb = 0; rb < 2; ++rb)
#pragma unroll
        for (int qb = 0; qb < 2; ++qb) {
            float res[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int pi = rb * 64 + ty * 4 + r, pj = qb * 64 + tx * 4 + q;
                    const float d1 = sqrtf(sqdist(xs[0][0][pi], xs[0][1][pi], xs[0][2][pi], xs[1][0][pj], xs[1][1][pj], xs[1][2][pj]));
                    const float2 a2 = acc[rb * 4 + r][qb * 2 + (q >> 1)];
                    res[r][q] = feat ? __fadd_rn(d1, __fmul_rn(sqrtf((q & 1) ? a2.y : a2.x), gamma)) : d1;
                }
            const int gi0 = i0 + rb * 64 + ty * 4, gj0 = j0 + qb * 64 + tx * 4;
#pragma unroll
            for (int r = 0; r < 4; ++r) {   // rows i, 4 consecutive columns j: 16 lanes cover 256 contiguous bytes
                const int gi = gi0 + r;
                if (gi >= n) continue;
                float *dst = out + (size_t)gi * n + gj0;
                if (vec_ok && gj0 + 3 < n) {
                    *reinterpret_cast<float4 *>(dst) = make_float4(res[r][0], res[r][1], res[r][2], res[r][3]);
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (gj0 + q < n) dst[q] = res[r][q];
                }
            }
            if (ti != tj) {                 // transpose: rows j, 4 consecutive columns i (full 32-byte sectors)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int gj = gj0 + q;
                    if (gj >= n) continue;
                    float *dst = out + (size_t)gj * n + gi0;
                    if (vec_ok && gi0 + 3 < n) {
                        *reinterpret_cast<float4 *>(dst) = make_float4(res[0][q], res[1][q], res[2][q], res[3][q]);
                    } else {
#pragma unroll
                        for (int r = 0; r < 4; ++r)
                            if (gi0 + r < n) dst[r] = res[r][q];
                    }
                }
            }
        }
}

}  // namespace de6d

using namespace de6d;

extern "C" int de6d_dist_matrix(int b, int n, int c, const float *xyz, const float *features, long long stride_b,
                                long long stride_n, long long stride_c, float gamma, float *out, cudaStream_t stream) {
    if (b < 0 || n < 0 || c < 0) return de6d_set_error(DE6D_ERR_INVALID, "dist_matrix: negative size");
    if (b == 0 || n == 0) return DE6D_OK;
    if (!xyz || !out) return de6d_set_error(DE6D_ERR_INVALID, "dist_matrix: null pointer");
    if (b > 65535) return de6d_set_error(DE6D_ERR_INVALID, "dist_matrix: batch > 65535");
    if (c == 0) features = nullptr;
    const int T = ceil_div(n, DM_TILE);
    const long long tiles = (long long)T * (T + 1) / 2;
    if (tiles > 2147483647ll) return de6d_set_error(DE6D_ERR_INVALID, "dist_matrix: too many points");
    const int vec_ok = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    dim3 grid((unsigned)tiles, b);
    dist_matrix_kernel<<<grid, 256, 0, stream>>>(n, c, xyz, features, stride_b, stride_n, stride_c, gamma, out, vec_ok);
    DE6D_CHECK_LAUNCH("dist_matrix_kernel");
    return DE6D_OK;
}
