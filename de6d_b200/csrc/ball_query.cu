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
#include <math.h>

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

// ---------------------------------------------------------------------------------------------------------
// Grid path (large clouds, balls that are small against the cloud: the SA layers' regime).
//
// The brute-force kernel above tests every (query, point) pair: 4.3 G distance tests per launch at 64 frames x
// 4096 queries x 16384 points, and on sparse clouds no ball ever fills, so there is no early exit.  Here a
// uniform grid is built per cloud (one CTA per cloud: bounding box, cell size >= the padded radius chosen so the
// grid has at most BQG_CAP cells, shared-memory histogram, scan, scatter of (x, y, z, index) records).  A query
// then only tests the points of the <= 3x3x3 cells its padded bounding cube touches -- with the SAME rounded
// distance expression -- and keeps the `nsample` smallest indices among the hits in a sorted list, so the result
// is the first `nsample` hits in ascending index order exactly as the reference's serial scan produces them.
//
// Exactness of the candidate set: every term of d2 = fl(dz^2 + fl(dx^2 + fl(dy^2))) is non-negative and rounding
// is monotone, so d2 < r2 implies fl(da^2) < r2 for each axis a, hence |q_a - p_a| <= r (1 + 2^-21).  The query's
// cell range is computed from fl(q_a -+ R') with R' = 1.0001 r + 1e-6 (|q_a| + r), which brackets that interval
// even after the rounding of the subtraction, and the cell function (float subtract, multiply, floor, clamp) is
// monotone, so every point that can pass the test lies in a visited cell.  Points visited but outside the ball
// fail the exact test.  Queries whose candidate count exceeds `cand_limit` (balls that swallow a large part of
// the cloud) and clouds with a non-finite bounding box fall back to the reference's in-order scan with early
// exit, inside the same kernel.
// ---------------------------------------------------------------------------------------------------------
constexpr int BQG_CAP = 32768;          // max cells per cloud (128 KB histogram in shared memory)
constexpr int BQG_CAP_BIG = 1 << 21;    // clouds of more than BQG_BIG_N points: histogram in global memory, finer cells
constexpr int BQG_BIG_N = 32768;
constexpr int BQG_BUILD_T = 1024;
constexpr int BQG_QT = 256;

struct BQGridHeader {
    float ox, oy, oz, inv_c;
    int dimx, dimy, dimz, valid;
};

// workspace of one cloud: header | cell_start[cap + 4] | sorted float4[n]
__host__ __device__ inline int bqg_cap(int n) { return n > BQG_BIG_N ? BQG_CAP_BIG : BQG_CAP; }
__host__ __device__ inline size_t bqg_sorted_off(int n) { return sizeof(BQGridHeader) + sizeof(int) * (size_t)(bqg_cap(n) + 4); }
__host__ __device__ inline size_t bqg_ws_per_cloud(int n) {
    size_t bytes = bqg_sorted_off(n) + 16 * (size_t)n;
    return (bytes + 255) & ~(size_t)255;
}

__device__ __forceinline__ int bqg_cell(float v, float o, float inv_c, int dim) {
    const int c = __float2int_rd(__fmul_rn(__fsub_rn(v, o), inv_c));   // saturating, NaN -> 0
    return min(max(c, 0), dim - 1);
}

// BIG = false: histogram / cursors in shared memory (<= BQG_CAP cells).  BIG = true: in the workspace itself
// (<= BQG_CAP_BIG cells): counts -> inclusive scan in place (tiles of 4096 with a running carry) -> scatter from the
// END of each cell with atomicSub, which leaves exactly the cell starts behind.
template <bool BIG>
__global__ void __launch_bounds__(BQG_BUILD_T)
bq_grid_build_kernel(int n, float r_abs, int cell_cap, const float *__restrict__ xyz_all, unsigned char *__restrict__ ws_all,
                     size_t ws_stride) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int *cnt = reinterpret_cast<int *>(smem_raw);            // [BQG_CAP] histogram, then scatter cursors (!BIG)
    constexpr int CAP = BIG ? BQG_CAP_BIG : BQG_CAP;
    __shared__ float red[6][BQG_BUILD_T / 32];
    __shared__ int wsum[BQG_BUILD_T / 32];
    __shared__ BQGridHeader sh;

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const float *xyz = xyz_all + (size_t)blockIdx.x * n * 3;
    unsigned char *ws = ws_all + (size_t)blockIdx.x * ws_stride;
    BQGridHeader *hdr = reinterpret_cast<BQGridHeader *>(ws);
    int *cell_start = reinterpret_cast<int *>(ws + sizeof(BQGridHeader));
    float4 *sorted = reinterpret_cast<float4 *>(ws + sizeof(BQGridHeader) + sizeof(int) * (size_t)(CAP + 4));

    // ---- bounding box ----
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    bool bad = false;
    for (int k = tid; k < n; k += BQG_BUILD_T) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = xyz[(size_t)k * 3 + a];
            bad |= !(fabsf(v) <= 1.0e30f);
            lo[a] = fminf(lo[a], v);
            hi[a] = fmaxf(hi[a], v);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
        if (lane == 0) { red[a][w] = lo[a]; red[3 + a][w] = hi[a]; }
    }
    const int anybad = __syncthreads_or(bad ? 1 : 0);
    if (tid == 0) {
        float l[3], e[3];
        for (int a = 0; a < 3; ++a) {
            float mn = INFINITY, mx = -INFINITY;
            for (int i = 0; i < BQG_BUILD_T / 32; ++i) { mn = fminf(mn, red[a][i]); mx = fmaxf(mx, red[3 + a][i]); }
            l[a] = mn; e[a] = mx - mn;
        }
        BQGridHeader h;
        h.valid = (n > 0 && !anybad && r_abs == r_abs && r_abs <= 1.0e30f) ? 1 : 0;
        float c = fmaxf(r_abs * 1.0002f, 1e-20f);
        // smallest cell size >= the padded radius with at most `cap` cells (cap <= CAP; three_nn asks for ~2 points per
        // cell).  The search starts just below the cube root of volume / cap and grows by 10 % steps.
        const int cap = min(max(cell_cap, 1), CAP);
        int dx = 1, dy = 1, dz = 1;
        if (h.valid) {
            const float vol = fmaxf(e[0], 1e-9f) * fmaxf(e[1], 1e-9f) * fmaxf(e[2], 1e-9f);
            const float c0 = 0.7f * cbrtf(vol / (float)cap);
            if (c0 == c0 && c0 < 1.0e30f) c = fmaxf(c, c0);
            for (int it = 0; it < 800; ++it) {
                const float fx = e[0] / c, fy = e[1] / c, fz = e[2] / c;
                if (fx < 30000.f && fy < 30000.f && fz < 30000.f) {
                    dx = (int)fx + 1; dy = (int)fy + 1; dz = (int)fz + 1;
                    if ((long long)dx * dy * dz <= cap) break;
                }
                c *= 1.1f;
                dx = dy = dz = 1;
            }
            if ((long long)dx * dy * dz > cap) { dx = dy = dz = 1; }
        }
        h.ox = l[0]; h.oy = l[1]; h.oz = l[2];
        h.inv_c = 1.0f / c;
        h.dimx = dx; h.dimy = dy; h.dimz = dz;
        sh = h;
        *hdr = h;
    }
    __syncthreads();
    const BQGridHeader h = sh;
    if (!h.valid) return;
    const int ncell = h.dimx * h.dimy * h.dimz;

    if (BIG) {
        // ---- histogram in the workspace (global atomics; this CTA is the only writer, reads go past L1) ----
        const int n4 = (ncell + 4) & ~3;   // cells + the closing entry, rounded up to whole int4
        for (int i = tid * 4; i < n4; i += BQG_BUILD_T * 4) *reinterpret_cast<int4 *>(cell_start + i) = make_int4(0, 0, 0, 0);
        __syncthreads();
        for (int k = tid; k < n; k += BQG_BUILD_T) {
            const float x = xyz[(size_t)k * 3], y = xyz[(size_t)k * 3 + 1], z = xyz[(size_t)k * 3 + 2];
            const int c = (bqg_cell(z, h.oz, h.inv_c, h.dimz) * h.dimy + bqg_cell(y, h.oy, h.inv_c, h.dimy)) * h.dimx +
                          bqg_cell(x, h.ox, h.inv_c, h.dimx);
            atomicAdd(cell_start + c, 1);
        }
        __threadfence();
        __syncthreads();
        // ---- inclusive scan in place: cell_start[c] = end of cell c ----
        int carry = 0;
        for (int base = 0; base < n4; base += BQG_BUILD_T * 4) {
            const int i = base + tid * 4;
            int4 v = make_int4(0, 0, 0, 0);
            if (i < n4) v = __ldcg(reinterpret_cast<const int4 *>(cell_start + i));
            v.y += v.x; v.z += v.y; v.w += v.z;
            int incl = v.w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            __syncthreads();   // wsum of the previous tile fully consumed
            if (lane == 31) wsum[w] = incl;
            __syncthreads();
            if (w == 0) {
                int t2 = wsum[lane];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, t2, o);
                    if (lane >= o) t2 += t;
                }
                wsum[lane] = t2;
            }
            __syncthreads();
            const int off = carry + incl - v.w + (w ? wsum[w - 1] : 0);
            if (i < n4) *reinterpret_cast<int4 *>(cell_start + i) = make_int4(v.x + off, v.y + off, v.z + off, v.w + off);
            carry += wsum[31];
        }
        __threadfence();
        __syncthreads();
        // ---- scatter from the end of each cell; afterwards cell_start[c] = start of cell c, cell_start[ncell] = n ----
        for (int k = tid; k < n; k += BQG_BUILD_T) {
            const float x = xyz[(size_t)k * 3], y = xyz[(size_t)k * 3 + 1], z = xyz[(size_t)k * 3 + 2];
            const int c = (bqg_cell(z, h.oz, h.inv_c, h.dimz) * h.dimy + bqg_cell(y, h.oy, h.inv_c, h.dimy)) * h.dimx +
                          bqg_cell(x, h.ox, h.inv_c, h.dimx);
            const int pos = atomicSub(cell_start + c, 1) - 1;
            sorted[pos] = make_float4(x, y, z, __int_as_float(k));
        }
        return;
    }

    // ---- histogram ----
    for (int i = tid; i < ncell; i += BQG_BUILD_T) cnt[i] = 0;
    __syncthreads();
    for (int k = tid; k < n; k += BQG_BUILD_T) {
        const float x = xyz[(size_t)k * 3], y = xyz[(size_t)k * 3 + 1], z = xyz[(size_t)k * 3 + 2];
        const int c = (bqg_cell(z, h.oz, h.inv_c, h.dimz) * h.dimy + bqg_cell(y, h.oy, h.inv_c, h.dimy)) * h.dimx +
                      bqg_cell(x, h.ox, h.inv_c, h.dimx);
        atomicAdd(&cnt[c], 1);
    }
    __syncthreads();

    // ---- exclusive scan over ncell entries: thread t owns a contiguous chunk ----
    const int chunk = ceil_div(ncell, BQG_BUILD_T);
    const int c0 = min(tid * chunk, ncell), c1 = min(c0 + chunk, ncell);
    int sum = 0;
    for (int i = c0; i < c1; ++i) sum += cnt[i];
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    if (w == 0) {
        int v = wsum[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        wsum[lane] = v;
    }
    __syncthreads();
    int run = incl - sum + (w ? wsum[w - 1] : 0);
    for (int i = c0; i < c1; ++i) {
        const int c = cnt[i];
        cnt[i] = run;          // becomes the scatter cursor
        cell_start[i] = run;
        run += c;
    }
    if (tid == BQG_BUILD_T - 1) cell_start[ncell] = n;
    __syncthreads();

    // ---- scatter (order inside a cell is irrelevant: the query keeps the smallest indices) ----
    for (int k = tid; k < n; k += BQG_BUILD_T) {
        const float x = xyz[(size_t)k * 3], y = xyz[(size_t)k * 3 + 1], z = xyz[(size_t)k * 3 + 2];
        const int c = (bqg_cell(z, h.oz, h.inv_c, h.dimz) * h.dimy + bqg_cell(y, h.oy, h.inv_c, h.dimy)) * h.dimx +
                      bqg_cell(x, h.ox, h.inv_c, h.dimx);
        const int pos = atomicAdd(&cnt[c], 1);
        sorted[pos] = make_float4(x, y, z, __int_as_float(k));
    }
}

template <int MODE>
__global__ void __launch_bounds__(BQG_QT)
bq_grid_query_kernel(int n, int m, float r_abs, float r2_in, float r2_out, int nsample, int cand_limit,
                     const float *__restrict__ new_xyz, const float *__restrict__ xyz,
                     const unsigned char *__restrict__ ws_all, size_t ws_stride, int *__restrict__ idx_cnt,
                     int *__restrict__ idx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int *lists = reinterpret_cast<int *>(smem_raw);   // [BQG_QT][nsample + 1] (odd pitch: conflict-free per-thread rows)
    const int pitch = nsample | 1;

    const int tid = threadIdx.x, lane = tid & 31;
    const int bs = blockIdx.y;
    const int q = blockIdx.x * BQG_QT + tid;
    xyz += (size_t)bs * n * 3;
    new_xyz += (size_t)bs * m * 3;
    const unsigned char *ws = ws_all + (size_t)bs * ws_stride;
    const BQGridHeader h = *reinterpret_cast<const BQGridHeader *>(ws);
    const int *cell_start = reinterpret_cast<const int *>(ws + sizeof(BQGridHeader));
    const float4 *sorted = reinterpret_cast<const float4 *>(ws + bqg_sorted_off(n));
    int *list = lists + (size_t)tid * pitch;

    int L = 0;
    if (q < m && nsample > 0) {
        const float qx = new_xyz[(size_t)q * 3], qy = new_xyz[(size_t)q * 3 + 1], qz = new_xyz[(size_t)q * 3 + 2];
        bool serial = !h.valid;
        if (!serial) {
            const float px = __fmaf_rn(1e-6f, fabsf(qx) + r_abs, r_abs * 1.0001f);
            const float py = __fmaf_rn(1e-6f, fabsf(qy) + r_abs, r_abs * 1.0001f);
            const float pz = __fmaf_rn(1e-6f, fabsf(qz) + r_abs, r_abs * 1.0001f);
            const int ix0 = bqg_cell(__fsub_rn(qx, px), h.ox, h.inv_c, h.dimx), ix1 = bqg_cell(__fadd_rn(qx, px), h.ox, h.inv_c, h.dimx);
            const int iy0 = bqg_cell(__fsub_rn(qy, py), h.oy, h.inv_c, h.dimy), iy1 = bqg_cell(__fadd_rn(qy, py), h.oy, h.inv_c, h.dimy);
            const int iz0 = bqg_cell(__fsub_rn(qz, pz), h.oz, h.inv_c, h.dimz), iz1 = bqg_cell(__fadd_rn(qz, pz), h.oz, h.inv_c, h.dimz);
            int seen = 0;
            int last = 0x7fffffff;   // largest index in a full list
            for (int iz = iz0; iz <= iz1 && !serial; ++iz) {
                for (int iy = iy0; iy <= iy1; ++iy) {
                    const int row = (iz * h.dimy + iy) * h.dimx;
                    const int s = __ldg(cell_start + row + ix0), e = __ldg(cell_start + row + ix1 + 1);
                    seen += e - s;
                    if (seen > cand_limit) { serial = true; break; }
                    // candidates four at a time: the 16-byte records are random L2 reads and the test -> insert chain behind
                    // each of them is short, so one load in flight per thread left the kernel waiting on L2 latency
                    for (int i = s; i < e; i += 4) {
                        float4 pc[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) pc[j] = __ldg(sorted + min(i + j, e - 1));
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (i + j >= e) break;
                            const float4 p = pc[j];
                            const int k = __float_as_int(p.w);
                            if (k > last) continue;   // list is full and this index cannot enter it
                            const float d2 = sqdist(qx, qy, qz, p.x, p.y, p.z);
                            bool hit = d2 < r2_out;
                            if (MODE == BQ_DILATED) hit = hit && (d2 >= r2_in);
                            if (!hit) continue;
                            int pos = (L < nsample) ? L : nsample - 1;   // full: the current largest drops out
                            while (pos > 0 && list[pos - 1] > k) { list[pos] = list[pos - 1]; --pos; }
                            list[pos] = k;
                            if (L < nsample) ++L;
                            if (L == nsample) last = list[nsample - 1];
                        }
                    }
                }
            }
        }
        if (serial) {   // the reference's scan: ascending index, stop at nsample hits
            L = 0;
            for (int k = 0; k < n; ++k) {
                const float d2 = sqdist(qx, qy, qz, xyz[(size_t)k * 3], xyz[(size_t)k * 3 + 1], xyz[(size_t)k * 3 + 2]);
                bool hit = d2 < r2_out;
                if (MODE == BQ_DILATED) hit = hit && (d2 >= r2_in);
                if (hit) {
                    list[L++] = k;
                    if (L >= nsample) break;
                }
            }
        }
    }
    __syncwarp();

    // ---- rows out, one warp-wide store per 32 slots: hits, then the variant's padding ----
    const int q_base = blockIdx.x * BQG_QT + (tid & ~31);
    for (int j = 0; j < 32; ++j) {
        const int qi = q_base + j;
        const int c = __shfl_sync(0xffffffffu, L, j);
        if (qi >= m) break;
        if (MODE != BQ_PLAIN && lane == 0) idx_cnt[(size_t)bs * m + qi] = c;
        if (c == 0) continue;   // reference leaves the (pre-zeroed) row untouched
        const int *src = lists + (size_t)((tid & ~31) + j) * pitch;
        int *row = idx + ((size_t)bs * m + qi) * nsample;
        for (int s = lane; s < nsample; s += 32) {
            int v;
            if (s < c) v = src[s];
            else v = (MODE == BQ_PLAIN) ? src[0] : src[s % c];
            row[s] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// three_nn through the same grid (interpolate_gpu.cu:16-59 scans all m known points per query).  The grid is built
// over `known` with ~2 points per cell.  A query looks at the 3x3x3 block around its cell (then 5x5x5), keeping the
// three best (distance, index) pairs in lexicographic order -- exactly what the reference's ascending scan with strict
// '<' produces.  The result is final once the third distance is provably smaller than the distance to any point
// outside the block: a point outside lies beyond a block face that is not on the grid boundary, i.e. at least `margin`
// away along that axis (0.1 % + 1e-4 cell slack for the rounding of the cell assignment).  Otherwise the query falls
// back to the reference's full scan.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void nn3_insert(float d, int k, float &b1, float &b2, float &b3, int &i1, int &i2, int &i3) {
    if (d < b1 || (d == b1 && k < i1)) { b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k; }
    else if (d < b2 || (d == b2 && k < i2)) { b3 = b2; i3 = i2; b2 = d; i2 = k; }
    else if (d < b3 || (d == b3 && k < i3)) { b3 = d; i3 = k; }
}

__global__ void __launch_bounds__(256)
three_nn_grid_kernel(int n, int m, const float *__restrict__ unknown, const float *__restrict__ known,
                     const unsigned char *__restrict__ ws_all, size_t ws_stride, float *__restrict__ dist2, int *__restrict__ idx) {
    const int bs = blockIdx.y;
    const int q = blockIdx.x * 256 + threadIdx.x;
    if (q >= n) return;
    unknown += (size_t)bs * n * 3;
    known += (size_t)bs * m * 3;
    const unsigned char *ws = ws_all + (size_t)bs * ws_stride;
    const BQGridHeader h = *reinterpret_cast<const BQGridHeader *>(ws);
    const int *cell_start = reinterpret_cast<const int *>(ws + sizeof(BQGridHeader));
    const float4 *sorted = reinterpret_cast<const float4 *>(ws + bqg_sorted_off(m));
    const float ux = unknown[(size_t)q * 3], uy = unknown[(size_t)q * 3 + 1], uz = unknown[(size_t)q * 3 + 2];
    float b1 = INFINITY, b2 = INFINITY, b3 = INFINITY;
    int i1 = 0, i2 = 0, i3 = 0;
    bool done = false;
    if (h.valid && ux == ux && uy == uy && uz == uz) {
        const float c = 1.0f / h.inv_c;
        const int cx = bqg_cell(ux, h.ox, h.inv_c, h.dimx), cy = bqg_cell(uy, h.oy, h.inv_c, h.dimy), cz = bqg_cell(uz, h.oz, h.inv_c, h.dimz);
        for (int R = 1; R <= 2 && !done; ++R) {
            b1 = b2 = b3 = INFINITY; i1 = i2 = i3 = 0;
            const int x0 = max(cx - R, 0), x1 = min(cx + R, h.dimx - 1), y0 = max(cy - R, 0), y1 = min(cy + R, h.dimy - 1);
            const int z0 = max(cz - R, 0), z1 = min(cz + R, h.dimz - 1);
            int found = 0;
            for (int iz = z0; iz <= z1; ++iz)
                for (int iy = y0; iy <= y1; ++iy) {
                    const int row = (iz * h.dimy + iy) * h.dimx;
                    const int s = __ldg(cell_start + row + x0), e = __ldg(cell_start + row + x1 + 1);
                    for (int i = s; i < e; ++i) {
                        const float4 p = __ldg(sorted + i);
                        nn3_insert(sqdist(ux, uy, uz, p.x, p.y, p.z), __float_as_int(p.w), b1, b2, b3, i1, i2, i3);
                    }
                    found += e - s;
                }
            if (found >= 3) {
                // distance from the query to the nearest block face behind which points can exist
                float margin = INFINITY;
                if (x0 > 0) margin = fminf(margin, ux - (h.ox + (float)x0 * c));
                if (x1 < h.dimx - 1) margin = fminf(margin, (h.ox + (float)(x1 + 1) * c) - ux);
                if (y0 > 0) margin = fminf(margin, uy - (h.oy + (float)y0 * c));
                if (y1 < h.dimy - 1) margin = fminf(margin, (h.oy + (float)(y1 + 1) * c) - uy);
                if (z0 > 0) margin = fminf(margin, uz - (h.oz + (float)z0 * c));
                if (z1 < h.dimz - 1) margin = fminf(margin, (h.oz + (float)(z1 + 1) * c) - uz);
                const float safe = margin * 0.999f - 1e-4f * c - 1e-6f * (fabsf(ux) + fabsf(uy) + fabsf(uz));
                done = (safe > 0.f) && (b3 < safe * safe * 0.999f);
            }
        }
    }
    if (!done) {   // the reference's scan
        b1 = b2 = b3 = INFINITY; i1 = i2 = i3 = 0;
        for (int k = 0; k < m; ++k) {
            const float d = sqdist(ux, uy, uz, known[(size_t)k * 3], known[(size_t)k * 3 + 1], known[(size_t)k * 3 + 2]);
            if (d < b1) { b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k; }
            else if (d < b2) { b3 = b2; i3 = i2; b2 = d; i2 = k; }
            else if (d < b3) { b3 = d; i3 = k; }
        }
    }
    float *dd = dist2 + ((size_t)bs * n + q) * 3;
    int *ii = idx + ((size_t)bs * n + q) * 3;
    dd[0] = b1; dd[1] = b2; dd[2] = b3;
    ii[0] = i1; ii[1] = i2; ii[2] = i3;
}

constexpr int BQG_MIN_N = 2048;   // below this the brute-force kernel is already cheap

// One uniform grid per cloud over `xyz` with cells no smaller than the padded radius `r_abs` (and at most the cell cap).
// The query kernel derives its cell range from the query's padded cube, so a grid built for radius r serves every
// radius (a larger ball simply spans more cells): the radius scales of one SA layer share one grid.
static int bq_build_grid(int b, int n, float r_abs, const float *xyz, void *ws, size_t per, cudaStream_t s) {
    static unsigned long long dev_build = 0;
    if (int rc = de6d_ensure_smem(bq_grid_build_kernel<false>, BQG_CAP * 4, dev_build, "ball_query grid smem attribute")) return rc;
    if (n > BQG_BIG_N) bq_grid_build_kernel<true><<<b, BQG_BUILD_T, 0, s>>>(n, r_abs, BQG_CAP_BIG, xyz, reinterpret_cast<unsigned char *>(ws), per);
    else bq_grid_build_kernel<false><<<b, BQG_BUILD_T, BQG_CAP * 4, s>>>(n, r_abs, BQG_CAP, xyz, reinterpret_cast<unsigned char *>(ws), per);
    DE6D_CHECK_LAUNCH("bq_grid_build_kernel");
    return DE6D_OK;
}

template <int MODE>
static int launch_ball_query(int b, int n, int m, float r_in, float r_out, int nsample, const float *new_xyz,
                             const float *xyz, int *idx_cnt, int *idx, int impl, void *workspace,
                             size_t workspace_bytes, cudaStream_t s) {
    if (b < 0 || n < 0 || m < 0 || nsample < 0) return de6d_set_error(DE6D_ERR_INVALID, "ball_query: negative size");
    if (b == 0 || m == 0) return DE6D_OK;
    if (!new_xyz || !idx || (n > 0 && !xyz) || (MODE != BQ_PLAIN && !idx_cnt))
        return de6d_set_error(DE6D_ERR_INVALID, "ball_query: null pointer");
    if (b > 65535) return de6d_set_error(DE6D_ERR_INVALID, "ball_query: batch > 65535");
    // radius*radius in float, as the reference kernels compute it (ball_query_gpu.cu:29,74-75,114)
    const float r2_in = r_in * r_in, r2_out = r_out * r_out;

    const size_t list_smem = (size_t)BQG_QT * (size_t)((nsample | 1)) * sizeof(int);
    bool grid_ok = n > 0 && nsample > 0 && list_smem <= 160 * 1024;
    bool use_grid = grid_ok && (impl == 2 || impl == 3 || (impl == 0 && n >= BQG_MIN_N));
    if ((impl == 2 || impl == 3) && !grid_ok) return de6d_set_error(DE6D_ERR_INVALID, "ball_query: grid kernel not applicable");
    if (use_grid) {
        const size_t per = bqg_ws_per_cloud(n), need = per * (size_t)b;
        void *ws = workspace;
        bool own = false;
        if (impl == 3 && !ws) return de6d_set_error(DE6D_ERR_INVALID, "ball_query: impl 3 needs the workspace de6d_ball_query_grid_build filled");
        if (!ws) {   // raw C callers without a workspace: stream-ordered scratch
            cudaError_t e = cudaMallocAsync(&ws, need, s);
            if (e != cudaSuccess) return de6d_set_cuda_error(e, "ball_query workspace");
            own = true;
        } else if (workspace_bytes < need) {
            return de6d_set_error(DE6D_ERR_INVALID, "ball_query: workspace too small");
        }
        static unsigned long long dev_query = 0;   // per call site (one per MODE instantiation)
        if (int rc = de6d_ensure_smem(bq_grid_query_kernel<MODE>, 160 * 1024, dev_query, "ball_query query smem attribute")) return rc;
        const float r_abs = fabsf(r_out);
        if (impl != 3)   // impl 3: the grid in `workspace` was built by de6d_ball_query_grid_build (shared by several radii)
            if (int rc = bq_build_grid(b, n, r_abs, xyz, ws, per, s)) return rc;
        double lim = 3.0 * sqrt((double)nsample * (double)n);
        if (lim < 1024.0) lim = 1024.0;
        if (lim > 2.0e9) lim = 2.0e9;
        dim3 grid(ceil_div(m, BQG_QT), b);
        bq_grid_query_kernel<MODE><<<grid, BQG_QT, list_smem, s>>>(n, m, r_abs, r2_in, r2_out, nsample, (int)lim, new_xyz, xyz,
                                                                 reinterpret_cast<const unsigned char *>(ws), per, idx_cnt, idx);
        DE6D_CHECK_LAUNCH("bq_grid_query_kernel");
        if (own) {
            cudaError_t e = cudaFreeAsync(ws, s);
            if (e != cudaSuccess) return de6d_set_cuda_error(e, "ball_query workspace free");
        }
        return DE6D_OK;
    }

    size_t smem = (size_t)BQ_TILE * 12 + (size_t)BQ_QPB * (nsample > 0 ? nsample : 1) * sizeof(int);
    if (smem > 200 * 1024) return de6d_set_error(DE6D_ERR_INVALID, "ball_query: nsample too large");
    static unsigned long long dev_bf = 0;
    if (smem > 48 * 1024)
        if (int rc = de6d_ensure_smem(ball_query_kernel<MODE>, 200 * 1024, dev_bf, "ball_query smem attribute")) return rc;
    dim3 grid(ceil_div(m, BQ_QPB), b);
    ball_query_kernel<MODE><<<grid, BQ_NWARP * 32, smem, s>>>(n, m, r2_in, r2_out, nsample, new_xyz, xyz, idx_cnt, idx);
    DE6D_CHECK_LAUNCH("ball_query_kernel");
    return DE6D_OK;
}

}  // namespace de6d

using namespace de6d;

// three_nn through the grid (called from interpolate.cu); workspace as for ball query over `known`, or NULL.
int de6d_three_nn_grid(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx, void *workspace,
                       size_t workspace_bytes, cudaStream_t s) {
    const size_t per = bqg_ws_per_cloud(m), need = per * (size_t)b;
    void *ws = workspace;
    bool own = false;
    if (!ws) {
        cudaError_t e = cudaMallocAsync(&ws, need, s);
        if (e != cudaSuccess) return de6d_set_cuda_error(e, "three_nn workspace");
        own = true;
    } else if (workspace_bytes < need) {
        return de6d_set_error(DE6D_ERR_INVALID, "three_nn: workspace too small");
    }
    static unsigned long long dev_build = 0;
    if (int rc = de6d_ensure_smem(bq_grid_build_kernel<false>, BQG_CAP * 4, dev_build, "three_nn grid smem attribute")) return rc;
    const int cap = m / 2 > 1 ? m / 2 : 1;
    if (m > BQG_BIG_N) bq_grid_build_kernel<true><<<b, BQG_BUILD_T, 0, s>>>(m, 0.f, cap, known, reinterpret_cast<unsigned char *>(ws), per);
    else bq_grid_build_kernel<false><<<b, BQG_BUILD_T, BQG_CAP * 4, s>>>(m, 0.f, cap, known, reinterpret_cast<unsigned char *>(ws), per);
    DE6D_CHECK_LAUNCH("bq_grid_build_kernel (three_nn)");
    dim3 grid(ceil_div(n, 256), b);
    three_nn_grid_kernel<<<grid, 256, 0, s>>>(n, m, unknown, known, reinterpret_cast<const unsigned char *>(ws), per, dist2, idx);
    DE6D_CHECK_LAUNCH("three_nn_grid_kernel");
    if (own) {
        cudaError_t e = cudaFreeAsync(ws, s);
        if (e != cudaSuccess) return de6d_set_cuda_error(e, "three_nn workspace free");
    }
    return DE6D_OK;
}

extern "C" int de6d_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz,
                               int *idx, cudaStream_t stream) {
    return launch_ball_query<BQ_PLAIN>(b, n, m, 0.f, radius, nsample, new_xyz, xyz, nullptr, idx, 0, nullptr, 0, stream);
}
extern "C" int de6d_ball_query_cnt(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                                   const float *xyz, int *idx_cnt, int *idx, cudaStream_t stream) {
    return launch_ball_query<BQ_CNT>(b, n, m, 0.f, radius, nsample, new_xyz, xyz, idx_cnt, idx, 0, nullptr, 0, stream);
}
extern "C" int de6d_ball_query_dilated(int b, int n, int m, float radius_in, float radius_out, int nsample,
                                       const float *new_xyz, const float *xyz, int *idx_cnt, int *idx,
                                       cudaStream_t stream) {
    return launch_ball_query<BQ_DILATED>(b, n, m, radius_in, radius_out, nsample, new_xyz, xyz, idx_cnt, idx, 0, nullptr, 0, stream);
}

// Scratch the automatic kernel choice needs for (b, n): 0 when the brute-force kernel will be used.
extern "C" size_t de6d_ball_query_workspace_bytes(int b, int n) {
    if (b <= 0 || n < BQG_MIN_N) return 0;
    return bqg_ws_per_cloud(n) * (size_t)b;
}

// Build the search grid of `b` clouds once, for use by several de6d_ball_query_ex(..., impl = 3, ...) calls on the same
// `xyz` (e.g. the radius scales of one SA layer, pointnet2_modules.py:462: three groupers, one cloud).  `radius`: the
// SMALLEST radius that will be queried (cells are no smaller than it; any radius is answered exactly).
extern "C" int de6d_ball_query_grid_build(int b, int n, float radius, const float *xyz, void *workspace, size_t workspace_bytes,
                                          cudaStream_t stream) {
    if (b < 0 || n < 0) return de6d_set_error(DE6D_ERR_INVALID, "ball_query_grid_build: negative size");
    if (b == 0 || n == 0) return DE6D_OK;
    if (!xyz || !workspace) return de6d_set_error(DE6D_ERR_INVALID, "ball_query_grid_build: null pointer");
    if (b > 65535) return de6d_set_error(DE6D_ERR_INVALID, "ball_query_grid_build: batch > 65535");
    const size_t per = bqg_ws_per_cloud(n);
    if (workspace_bytes < per * (size_t)b) return de6d_set_error(DE6D_ERR_INVALID, "ball_query_grid_build: workspace too small");
    return bq_build_grid(b, n, fabsf(radius), xyz, workspace, per, stream);
}
// Bytes de6d_ball_query_grid_build needs for (b, n) -- unlike de6d_ball_query_workspace_bytes not 0 for small clouds.
extern "C" size_t de6d_ball_query_grid_bytes(int b, int n) {
    if (b <= 0 || n <= 0) return 0;
    return bqg_ws_per_cloud(n) * (size_t)b;
}

// mode: 0 plain (ball_query), 1 counted (ball_query_cnt), 2 dilated.  impl: 0 auto, 1 brute-force kernel, 2 grid
// kernel (builds its grid), 3 grid kernel over the grid de6d_ball_query_grid_build left in `workspace`.  workspace: de6d_ball_query_workspace_bytes(b, n) bytes of device scratch, or NULL (the grid path then
// takes stream-ordered scratch from cudaMallocAsync).
extern "C" int de6d_ball_query_ex(int mode, int impl, int b, int n, int m, float radius_in, float radius_out, int nsample,
                                  const float *new_xyz, const float *xyz, int *idx_cnt, int *idx, void *workspace,
                                  size_t workspace_bytes, cudaStream_t stream) {
    if (mode == 0) return launch_ball_query<BQ_PLAIN>(b, n, m, 0.f, radius_out, nsample, new_xyz, xyz, nullptr, idx, impl, workspace, workspace_bytes, stream);
    if (mode == 1) return launch_ball_query<BQ_CNT>(b, n, m, 0.f, radius_out, nsample, new_xyz, xyz, idx_cnt, idx, impl, workspace, workspace_bytes, stream);
    if (mode == 2) return launch_ball_query<BQ_DILATED>(b, n, m, radius_in, radius_out, nsample, new_xyz, xyz, idx_cnt, idx, impl, workspace, workspace_bytes, stream);
    return de6d_set_error(DE6D_ERR_INVALID, "ball_query_ex: unknown mode");
}
