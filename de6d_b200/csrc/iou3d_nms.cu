// Rotated BEV overlap / IoU, fused 3-D IoU, and batched rotated / axis-aligned NMS for sm_100a.
//
// Replaces iou3d_nms/src/iou3d_nms_kernel.cu:104-372 + the host round trip of iou3d_nms.cpp:90-186
// (cudaMalloc, kernel over the full square of 64x64 tiles, blocking D2H, host greedy sweep, cudaFree) and the
// six-kernel torch composition of iou3d_nms_utils.py:48-81.
//
// Same polygon-clipping algorithm as the reference (16 edge/edge intersections, corner-in-box tests with
// MARGIN 1e-2, angular ordering about the centroid, fan area), organised differently:
//   * everything that depends on one box only -- rotated corners, cos/sin of +-heading, padded half extents,
//     area, bounding radius -- is computed once per box per tile into shared memory instead of once per pair
//     (the reference evaluates 2 + 16 sincos per pair);
//   * a bounding-circle test rejects pairs that cannot touch: for those the reference finds no intersection
//     and no contained corner, i.e. overlap == +0 exactly, so the shortcut is value-preserving;
//   * polar angles are evaluated once per polygon vertex and the vertices are put in order with a stable
//     insertion sort -- the reference's bubble sort with strict '>' is stable too, so the permutation is equal;
//   * NMS: only tiles on or above the diagonal are evaluated (the reference sweep never reads the others),
//     the suppression words stay on the device, and the LAST tile-CTA of each frame (atomic ticket) runs the
//     greedy sweep out of shared memory -- one launch for a whole batch of frames, no host synchronisation.
#include "common.cuh"
#include "box9.cuh"
#include <math.h>
#include <type_traits>

namespace de6d {

constexpr float IOU_EPS = 1e-8f;
constexpr float BOX_MARGIN = 1e-2f;

struct BoxGeo {
    float px[4], py[4];  // rotated corners (x1,y1) (x2,y1) (x2,y2) (x1,y2), reference order
    float cx, cy;
    float ic, is;        // cos(-heading), sin(-heading)
    float hx, hy;        // dx/2 + MARGIN, dy/2 + MARGIN
    float area;          // dx*dy
    float rad;           // radius of a circle around (cx,cy) containing the margin-padded box
};

__device__ __forceinline__ BoxGeo make_geo(const float *b) {
    BoxGeo g;
    const float x = b[0], y = b[1], dx = b[3], dy = b[4], ang = b[6];
    const float hx = dx / 2, hy = dy / 2;
    const float x1 = x - hx, y1 = y - hy, x2 = x + hx, y2 = y + hy;
    const float c = cosf(ang), s = sinf(ang);
    const float qx[4] = {x1, x2, x2, x1}, qy[4] = {y1, y1, y2, y2};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        g.px[k] = (qx[k] - x) * c + (qy[k] - y) * (-s) + x;
        g.py[k] = (qx[k] - x) * s + (qy[k] - y) * c + y;
    }
    g.cx = x; g.cy = y;
    g.ic = cosf(-ang); g.is = sinf(-ang);
    g.hx = dx / 2 + BOX_MARGIN; g.hy = dy / 2 + BOX_MARGIN;
    g.area = dx * dy;
    const float ex = fabsf(dx) * 0.5f + 0.03f, ey = fabsf(dy) * 0.5f + 0.03f;
    g.rad = sqrtf(ex * ex + ey * ey) * 1.001f;
    return g;
}

__device__ __forceinline__ float cross3(float p1x, float p1y, float p2x, float p2y, float p0x, float p0y) {
    return (p1x - p0x) * (p2y - p0y) - (p2x - p0x) * (p1y - p0y);
}

__device__ __forceinline__ bool corner_inside(const BoxGeo &g, float px, float py) {
    const float rx = (px - g.cx) * g.ic + (py - g.cy) * (-g.is);
    const float ry = (px - g.cx) * g.is + (py - g.cy) * g.ic;
    return fabsf(rx) < g.hx && fabsf(ry) < g.hy;
}

// edge p0->p1 against edge q0->q1; writes the crossing point
__device__ __forceinline__ bool edge_cross(float p1x, float p1y, float p0x, float p0y, float q1x, float q1y, float q0x,
                                           float q0y, float &ox, float &oy) {
    const bool boxes_touch = fminf(p0x, p1x) <= fmaxf(q0x, q1x) && fminf(q0x, q1x) <= fmaxf(p0x, p1x) &&
                             fminf(p0y, p1y) <= fmaxf(q0y, q1y) && fminf(q0y, q1y) <= fmaxf(p0y, p1y);
    if (!boxes_touch) return false;
    const float s1 = cross3(q0x, q0y, p1x, p1y, p0x, p0y);
    const float s2 = cross3(p1x, p1y, q1x, q1y, p0x, p0y);
    const float s3 = cross3(p0x, p0y, q1x, q1y, q0x, q0y);
    const float s4 = cross3(q1x, q1y, p1x, p1y, q0x, q0y);
    if (!(s1 * s2 > 0 && s3 * s4 > 0)) return false;
    const float s5 = cross3(q1x, q1y, p1x, p1y, p0x, p0y);
    if (fabsf(s5 - s1) > IOU_EPS) {
        ox = (s5 * q0x - s1 * q1x) / (s5 - s1);
        oy = (s5 * q0y - s1 * q1y) / (s5 - s1);
    } else {
        const float a0 = p0y - p1y, b0 = p1x - p0x, c0 = p0x * p1y - p1x * p0y;
        const float a1 = q0y - q1y, b1 = q1x - q0x, c1 = q0x * q1y - q1x * q0y;
        const float D = a0 * b1 - a1 * b0;
        ox = (b0 * c1 - b1 * c0) / D;
        oy = (a1 * c0 - a0 * c1) / D;
    }
    return true;
}

__device__ __forceinline__ bool cannot_touch(const BoxGeo &a, const BoxGeo &b) {
    const float ddx = a.cx - b.cx, ddy = a.cy - b.cy, r = a.rad + b.rad;
    return ddx * ddx + ddy * ddy > r * r;  // false on NaN: falls through to the full evaluation
}

__device__ float overlap_area(const BoxGeo &a, const BoxGeo &b) {
    if (cannot_touch(a, b)) return 0.f;
    float vx[16], vy[16], va[16];
    float sumx = 0.f, sumy = 0.f;
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int i1 = (i + 1) & 3;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int j1 = (j + 1) & 3;
            float ox, oy;
            if (edge_cross(a.px[i1], a.py[i1], a.px[i], a.py[i], b.px[j1], b.py[j1], b.px[j], b.py[j], ox, oy)) {
                sumx = sumx + ox; sumy = sumy + oy;
                vx[cnt] = ox; vy[cnt] = oy; ++cnt;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (corner_inside(a, b.px[k], b.py[k])) {
            sumx = sumx + b.px[k]; sumy = sumy + b.py[k];
            vx[cnt] = b.px[k]; vy[cnt] = b.py[k]; ++cnt;
        }
        if (corner_inside(b, a.px[k], a.py[k])) {
            sumx = sumx + a.px[k]; sumy = sumy + a.py[k];
            vx[cnt] = a.px[k]; vy[cnt] = a.py[k]; ++cnt;
        }
    }
    if (cnt < 3) return 0.f;  // fan over fewer than three vertices has zero area in the reference too
    const float mx = sumx / cnt, my = sumy / cnt;
    for (int i = 0; i < cnt; ++i) va[i] = atan2f(vy[i] - my, vx[i] - mx);
    for (int i = 1; i < cnt; ++i) {  // stable insertion sort, ascending angle
        const float ka = va[i], kx = vx[i], ky = vy[i];
        int j = i - 1;
        while (j >= 0 && va[j] > ka) { va[j + 1] = va[j]; vx[j + 1] = vx[j]; vy[j + 1] = vy[j]; --j; }
        va[j + 1] = ka; vx[j + 1] = kx; vy[j + 1] = ky;
    }
    float area = 0.f;
    for (int k = 0; k < cnt - 1; ++k) {
        const float ax = vx[k] - vx[0], ay = vy[k] - vy[0], bx = vx[k + 1] - vx[0], by = vy[k + 1] - vy[0];
        area += ax * by - ay * bx;
    }
    return fabsf(area) * 0.5f;
}

__device__ __forceinline__ float iou_from_overlap(float s, float sa, float sb) { return s / fmaxf(sa + sb - s, IOU_EPS); }

__device__ __forceinline__ float iou_axis_aligned(const float *a, const float *b) {  // iou_normal :314-325
    const float left = fmaxf(a[0] - a[3] / 2, b[0] - b[3] / 2), right = fminf(a[0] + a[3] / 2, b[0] + b[3] / 2);
    const float top = fmaxf(a[1] - a[4] / 2, b[1] - b[4] / 2), bottom = fminf(a[1] + a[4] / 2, b[1] + b[4] / 2);
    const float w = fmaxf(right - left, 0.f), h = fmaxf(bottom - top, 0.f);
    const float inter = w * h, sa = a[3] * a[4], sb = b[3] * b[4];
    return inter / fmaxf(sa + sb - inter, IOU_EPS);
}

// ---------------------------------------------------------------------------------------------------------
// (N, M) matrices.  MODE 0: overlap area, 1: BEV IoU, 2: 3-D IoU (fused iou3d_nms_utils.py:48-81)
// ---------------------------------------------------------------------------------------------------------
constexpr int MT = 32;  // tile edge

template <int MODE>
__global__ void __launch_bounds__(256)
iou_matrix_kernel(int na, const float *__restrict__ boxes_a, int nb, const float *__restrict__ boxes_b,
                  float *__restrict__ out) {
    __shared__ BoxGeo ga[MT], gb[MT];
    __shared__ float za[MT][3], zb[MT][3];  // zmax, zmin, volume (MODE 2)
    __shared__ unsigned short s_queue[MT * MT];
    __shared__ int s_qn;
    const int a0 = blockIdx.y * MT, b0 = blockIdx.x * MT;
    const int tid = threadIdx.x;
    if (tid < 2 * MT) {
        const bool isb = tid >= MT;
        const int i = isb ? tid - MT : tid;
        const int gi = (isb ? b0 : a0) + i;
        if (gi < (isb ? nb : na)) {
            const float *bx = (isb ? boxes_b : boxes_a) + (size_t)gi * 7;
            (isb ? gb : ga)[i] = make_geo(bx);
            if (MODE == 2) {
                float *zz = isb ? zb[i] : za[i];
                zz[0] = __fadd_rn(bx[2], bx[5] / 2);
                zz[1] = __fsub_rn(bx[2], bx[5] / 2);
                zz[2] = __fmul_rn(__fmul_rn(bx[3], bx[4]), bx[5]);
            }
        }
    }
    __syncthreads();
    // value of one matrix entry from the BEV overlap s
    auto entry = [&](float s, int i, int j) -> float {
        if (MODE == 0) return s;
        if (MODE == 1) return iou_from_overlap(s, ga[i].area, gb[j].area);
        float h = __fsub_rn(fminf(za[i][0], zb[j][0]), fmaxf(za[i][1], zb[j][1]));
        h = fmaxf(h, 0.f);  // torch.clamp(min=0): NaN propagates in torch, fmaxf drops it; inputs are finite boxes
        const float o3 = __fmul_rn(s, h);
        const float den = fmaxf(__fsub_rn(__fadd_rn(za[i][2], zb[j][2]), o3), 1e-6f);
        return __fdiv_rn(o3, den);
    };
    // Phase 1: pairs whose bounding circles are apart have overlap 0 and are written at once; the others are queued
    // and clipped densely in phase 2 (a warp would otherwise wait on every lane that needs the ~2000-instruction clip).
    if (tid == 0) s_qn = 0;
    __syncthreads();
    const int j = tid & 31, bj = b0 + j;
    if (bj < nb) {
#pragma unroll 1
        for (int r = 0; r < 4; ++r) {
            const int i = (tid >> 5) + 8 * r, ai = a0 + i;
            if (ai >= na) break;
            if (cannot_touch(ga[i], gb[j])) out[(size_t)ai * nb + bj] = entry(0.f, i, j);
            else s_queue[atomicAdd(&s_qn, 1)] = (unsigned short)(i * MT + j);
        }
    }
    __syncthreads();
    const int qn = s_qn;
    for (int q = tid; q < qn; q += 256) {
        const int code = s_queue[q], i = code / MT, jq = code % MT;
        out[(size_t)(a0 + i) * nb + b0 + jq] = entry(overlap_area(ga[i], gb[jq]), i, jq);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Batched NMS.  boxes: (F, N, 7) already sorted by descending score per frame; nvalid (optional): boxes in
// frame f beyond nvalid[f] are ignored.  keep: (F, N) int64 positions kept (ascending), num_keep: (F).
// Workspace: F * N * ceil(N/64) suppression words + F tickets.
// ---------------------------------------------------------------------------------------------------------
constexpr int NMS_T = 256;

// MODE 0: rotated BEV IoU (nms_gpu), 1: axis-aligned BEV IoU (nms_normal_gpu), boxes (F, N, 7);
// MODE 2: full-pose 3-D IoU (box9.cuh), boxes (F, N, 9) -- the pitch / roll aware NMS Det6D's 9-DoF predictions call for.
template <int MODE>
__global__ void __launch_bounds__(NMS_T)
nms_kernel(int n, const float *__restrict__ boxes_all, const int *__restrict__ nvalid, float thresh,
           unsigned long long *__restrict__ mask_all, unsigned int *__restrict__ tickets, long long *__restrict__ keep_all,
           int *__restrict__ num_keep, int sweep_in_smem) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr bool NORMAL = MODE == 1;
    constexpr int BS = MODE == 2 ? 9 : 7;                      // floats per box
    using Geo = typename std::conditional<MODE == 2, Box9Geo, BoxGeo>::type;
    __shared__ Geo grow[64], gcol[64];
    __shared__ unsigned int s_last;
    __shared__ unsigned long long s_diag[64];
    __shared__ unsigned long long s_keepbits;
    __shared__ unsigned long long s_words[64];
    __shared__ unsigned short s_queue[64 * 64];
    __shared__ int s_qn;

    const int f = blockIdx.y;
    const int cb = ceil_div(n, 64);
    const int nv = nvalid ? min(max(nvalid[f], 0), n) : n;
    const float *boxes = boxes_all + (size_t)f * n * BS;
    unsigned long long *mask = mask_all + (size_t)f * n * cb;
    const int tid = threadIdx.x;

    // linear upper-triangular tile index -> (row tile, col tile), rt <= ct
    int rt = 0, rem = blockIdx.x;
    while (rem >= cb - rt) { rem -= cb - rt; ++rt; }
    const int ct = rt + rem;

    if (rt * 64 < nv) {  // tiles entirely past the valid boxes have nothing to write that the sweep reads
        if (!NORMAL && tid < 128) {
            const bool isc = tid >= 64;
            const int i = isc ? tid - 64 : tid;
            const int gi = (isc ? ct : rt) * 64 + i;
            if (gi < nv) {
                if constexpr (MODE == 2) (isc ? gcol : grow)[i] = make_geo9(boxes + (size_t)gi * 9);
                else (isc ? gcol : grow)[i] = make_geo(boxes + (size_t)gi * 7);
            }
        }
        __syncthreads();
        const int r = tid >> 2, part = tid & 3;  // row r, columns part*16 .. +15
        const int gi = rt * 64 + r;
        if (NORMAL) {
            unsigned int bits = 0;
            if (gi < nv) {
                const int jstart = (rt == ct) ? r + 1 : 0;
                for (int jj = 0; jj < 16; ++jj) {
                    const int j = part * 16 + jj, gj = ct * 64 + j;
                    if (j < jstart || gj >= nv) continue;
                    if (iou_axis_aligned(boxes + (size_t)gi * 7, boxes + (size_t)gj * 7) > thresh) bits |= 1u << jj;
                }
            }
            // four adjacent lanes hold the four 16-bit quarters of the word
            unsigned long long word = (unsigned long long)bits << (16 * part);
            word |= __shfl_xor_sync(0xffffffffu, word, 1);
            word |= __shfl_xor_sync(0xffffffffu, word, 2);
            if (part == 0 && gi < n) mask[(size_t)gi * cb + ct] = word;
        } else {
            // Two phases, because only a few percent of the pairs survive the bounding-circle test and the polygon
            // clipping behind it is ~2000 instructions: evaluated in place, one surviving pair stalls the other 31 lanes
            // of its warp.  Phase 1 queues the surviving pairs of the tile, phase 2 hands them out densely.
            if (tid < 64) s_words[tid] = 0ull;
            if (tid == 0) s_qn = 0;
            __syncthreads();
            if (gi < nv) {
                const int jstart = (rt == ct) ? r + 1 : 0;
                unsigned int cand = 0;
                for (int jj = 0; jj < 16; ++jj) {
                    const int j = part * 16 + jj, gj = ct * 64 + j;
                    if (j < jstart || gj >= nv) continue;
                    bool far;
                    if constexpr (MODE == 2) far = cannot_touch9(grow[r], gcol[j]);
                    else far = cannot_touch(grow[r], gcol[j]);
                    if (!far) cand |= 1u << jj;
                }
                if (cand) {
                    int at = atomicAdd(&s_qn, __popc(cand));
                    while (cand) {
                        const int jj = __ffs(cand) - 1;
                        cand &= cand - 1;
                        s_queue[at++] = (unsigned short)(r * 64 + part * 16 + jj);
                    }
                }
            }
            __syncthreads();
            const int qn = s_qn;
            for (int q = tid; q < qn; q += NMS_T) {
                const int code = s_queue[q], qr = code >> 6, qj = code & 63;
                float v;
                if constexpr (MODE == 2) v = iou9(grow[qr], gcol[qj]);
                else v = iou_from_overlap(overlap_area(grow[qr], gcol[qj]), grow[qr].area, gcol[qj].area);
                if (v > thresh) atomicOr(&s_words[qr], 1ull << qj);
            }
            __syncthreads();
            if (tid < 64 && rt * 64 + tid < n) mask[(size_t)(rt * 64 + tid) * cb + ct] = s_words[tid];
        }
    } else {
        const int r = tid >> 2, gi = rt * 64 + r;
        if ((tid & 3) == 0 && gi < n) mask[(size_t)gi * cb + ct] = 0ull;
    }

    // ---- ticket: the last tile-CTA of this frame performs the sweep ----
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned int total = gridDim.x;
        const unsigned int t = atomicAdd(&tickets[f], 1u);
        s_last = (t == total - 1) ? 1u : 0u;
        if (s_last) tickets[f] = 0u;  // self-cleaning for the next launch
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();

    unsigned long long *smask = reinterpret_cast<unsigned long long *>(smem_raw);
    const unsigned long long *mrd = mask;
    if (sweep_in_smem) {
        const size_t words = (size_t)n * cb;
        for (size_t i = tid; i < words; i += NMS_T) smask[i] = __ldcg(mask + i);
        mrd = smask;
    }
    // removed-words: thread j (< cb) owns word j in a register-like smem slot
    unsigned long long *remv = sweep_in_smem ? smask + (size_t)n * cb : reinterpret_cast<unsigned long long *>(smem_raw);
    for (int j = tid; j < cb; j += NMS_T) remv[j] = 0ull;
    __syncthreads();

    long long *keep = keep_all + (size_t)f * n;
    int nk = 0;  // meaningful in thread 0
    for (int nb = 0; nb < cb; ++nb) {
        const int bn = min(64, nv - nb * 64);
        if (bn <= 0) break;
        if (tid < 64) s_diag[tid] = tid < bn ? (sweep_in_smem ? mrd[(size_t)(nb * 64 + tid) * cb + nb] : __ldcg(mrd + (size_t)(nb * 64 + tid) * cb + nb)) : 0ull;
        __syncthreads();
        if (tid == 0) {
            unsigned long long word = remv[nb], kb = 0ull;
            for (int i = 0; i < bn; ++i) {
                if (!((word >> i) & 1ull)) {
                    kb |= 1ull << i;
                    word |= s_diag[i];
                    keep[nk++] = (long long)nb * 64 + i;
                }
            }
            s_keepbits = kb;
        }
        __syncthreads();
        const unsigned long long kb = s_keepbits;
        // OR the rows of the kept boxes into the later words: thread handles word j = nb+1+tid%..., rows striped
        for (int j = nb + 1 + (tid & 31); j < cb; j += 32) {
            unsigned long long acc = 0ull;
            for (int i = tid >> 5; i < bn; i += NMS_T / 32)
                if ((kb >> i) & 1ull) acc |= sweep_in_smem ? mrd[(size_t)(nb * 64 + i) * cb + j] : __ldcg(mrd + (size_t)(nb * 64 + i) * cb + j);
            if (acc) atomicOr(&remv[j], acc);
        }
        __syncthreads();
    }
    // entries beyond num_keep are defined: position 0 (callers gather through the whole padded row)
    __shared__ int s_nk;
    if (tid == 0) { num_keep[f] = nk; s_nk = nk; }
    __syncthreads();
    for (int i = s_nk + tid; i < n; i += NMS_T) keep[i] = 0;
}

}  // namespace de6d

using namespace de6d;

static int iou_matrix_launch(int mode, int na, const float *a, int nb, const float *b, float *out, cudaStream_t s) {
    if (na < 0 || nb < 0) return de6d_set_error(DE6D_ERR_INVALID, "boxes iou: negative size");
    if (na == 0 || nb == 0) return DE6D_OK;
    if (!a || !b || !out) return de6d_set_error(DE6D_ERR_INVALID, "boxes iou: null pointer");
    dim3 grid(ceil_div(nb, MT), ceil_div(na, MT));
    if (grid.y > 65535) return de6d_set_error(DE6D_ERR_INVALID, "boxes iou: too many boxes_a");
    if (mode == 0) iou_matrix_kernel<0><<<grid, 256, 0, s>>>(na, a, nb, b, out);
    else if (mode == 1) iou_matrix_kernel<1><<<grid, 256, 0, s>>>(na, a, nb, b, out);
    else iou_matrix_kernel<2><<<grid, 256, 0, s>>>(na, a, nb, b, out);
    DE6D_CHECK_LAUNCH("iou_matrix_kernel");
    return DE6D_OK;
}

extern "C" int de6d_boxes_overlap_bev(int na, const float *boxes_a, int nb, const float *boxes_b, float *ans_overlap,
                                      cudaStream_t stream) {
    return iou_matrix_launch(0, na, boxes_a, nb, boxes_b, ans_overlap, stream);
}
extern "C" int de6d_boxes_iou_bev(int na, const float *boxes_a, int nb, const float *boxes_b, float *ans_iou,
                                  cudaStream_t stream) {
    return iou_matrix_launch(1, na, boxes_a, nb, boxes_b, ans_iou, stream);
}
extern "C" int de6d_boxes_iou3d(int na, const float *boxes_a, int nb, const float *boxes_b, float *ans_iou,
                                cudaStream_t stream) {
    return iou_matrix_launch(2, na, boxes_a, nb, boxes_b, ans_iou, stream);
}

// Full-pose IoU matrix: boxes_a (na, 9) x boxes_b (nb, 9) -> (na, nb).  32 x 32 tiles: bounding-sphere reject, then the
// surviving pairs are queued in shared memory and clipped densely (the clip is ~5000 instructions and ~2 KB of local
// memory per thread: evaluated in place it would leave most lanes of a warp waiting on a few surviving pairs).
namespace de6d {
__global__ void __launch_bounds__(256)
iou9_matrix_kernel(int na, const float *__restrict__ a, int nb, const float *__restrict__ b, float *__restrict__ out) {
    __shared__ Box9Geo ga[32], gb[32];
    __shared__ unsigned short s_queue[32 * 32];
    __shared__ int s_qn;
    const int tid = threadIdx.x;
    const int a0 = blockIdx.y * 32, b0 = blockIdx.x * 32;
    if (tid < 32) { if (a0 + tid < na) ga[tid] = make_geo9(a + (size_t)(a0 + tid) * 9); }
    else if (tid < 64) { if (b0 + tid - 32 < nb) gb[tid - 32] = make_geo9(b + (size_t)(b0 + tid - 32) * 9); }
    if (tid == 0) s_qn = 0;
    __syncthreads();
    const int j = tid & 31, bj = b0 + j;
    if (bj < nb) {
#pragma unroll 1
        for (int r = 0; r < 4; ++r) {
            const int i = (tid >> 5) + 8 * r, ai = a0 + i;
            if (ai >= na) break;
            if (cannot_touch9(ga[i], gb[j])) out[(size_t)ai * nb + bj] = 0.f;
            else s_queue[atomicAdd(&s_qn, 1)] = (unsigned short)(i * 32 + j);
        }
    }
    __syncthreads();
    const int qn = s_qn;
    for (int q = tid; q < qn; q += 256) {
        const int code = s_queue[q], i = code >> 5, jq = code & 31;
        out[(size_t)(a0 + i) * nb + b0 + jq] = iou9(ga[i], gb[jq]);
    }
}
}  // namespace de6d

extern "C" int de6d_boxes_iou3d9(int na, const float *boxes_a, int nb, const float *boxes_b, float *ans_iou, cudaStream_t stream) {
    if (na < 0 || nb < 0) return de6d_set_error(DE6D_ERR_INVALID, "boxes_iou3d9: negative size");
    if (na == 0 || nb == 0) return DE6D_OK;
    if (!boxes_a || !boxes_b || !ans_iou) return de6d_set_error(DE6D_ERR_INVALID, "boxes_iou3d9: null pointer");
    dim3 grid(de6d::ceil_div(nb, 32), de6d::ceil_div(na, 32));
    if (grid.y > 65535) return de6d_set_error(DE6D_ERR_INVALID, "boxes_iou3d9: too many boxes");
    de6d::iou9_matrix_kernel<<<grid, 256, 0, stream>>>(na, boxes_a, nb, boxes_b, ans_iou);
    DE6D_CHECK_LAUNCH("iou9_matrix_kernel");
    return DE6D_OK;
}

extern "C" size_t de6d_nms_workspace_bytes(int frames, int n) {
    if (frames <= 0 || n <= 0) return 256;
    size_t cb = (size_t)(n + 63) / 64;
    size_t mask = (size_t)frames * n * cb * 8;
    size_t tickets = ((size_t)frames * 4 + 255) & ~(size_t)255;
    return tickets + mask;
}

// The ticket words (first 256-byte-rounded frames*4 bytes of the workspace) must be zero on entry; the kernel
// leaves them zero on exit.  de6d_nms_workspace_init zeroes them (call once after allocating the workspace).
extern "C" int de6d_nms_workspace_init(int frames, void *workspace, cudaStream_t stream) {
    if (frames <= 0) return DE6D_OK;
    if (!workspace) return de6d_set_error(DE6D_ERR_INVALID, "nms: null workspace");
    size_t tickets = ((size_t)frames * 4 + 255) & ~(size_t)255;
    cudaError_t e = cudaMemsetAsync(workspace, 0, tickets, stream);
    if (e != cudaSuccess) return de6d_set_cuda_error(e, "nms workspace memset");
    return DE6D_OK;
}

extern "C" int de6d_nms_batched(int frames, int n, const float *boxes, const int *nvalid, float thresh, int mode,
                                long long *keep, int *num_keep, void *workspace, size_t workspace_bytes,
                                cudaStream_t stream) {
    if (frames < 0 || n < 0) return de6d_set_error(DE6D_ERR_INVALID, "nms: negative size");
    if (frames == 0) return DE6D_OK;
    if (!num_keep) return de6d_set_error(DE6D_ERR_INVALID, "nms: null num_keep");
    if (n == 0) {
        cudaError_t e = cudaMemsetAsync(num_keep, 0, sizeof(int) * (size_t)frames, stream);
        if (e != cudaSuccess) return de6d_set_cuda_error(e, "nms memset");
        return DE6D_OK;
    }
    if (!boxes || !keep || !workspace) return de6d_set_error(DE6D_ERR_INVALID, "nms: null pointer");
    if (workspace_bytes < de6d_nms_workspace_bytes(frames, n)) return de6d_set_error(DE6D_ERR_INVALID, "nms: workspace too small");
    if (frames > 65535) return de6d_set_error(DE6D_ERR_INVALID, "nms: more than 65535 frames per call");
    const int cb = ceil_div(n, 64);
    const long long tiles = (long long)cb * (cb + 1) / 2;
    if (tiles > 2147483647ll) return de6d_set_error(DE6D_ERR_INVALID, "nms: too many boxes");
    size_t tickets_b = ((size_t)frames * 4 + 255) & ~(size_t)255;
    unsigned int *tickets = reinterpret_cast<unsigned int *>(workspace);
    unsigned long long *mask = reinterpret_cast<unsigned long long *>(reinterpret_cast<unsigned char *>(workspace) + tickets_b);
    size_t smem_full = ((size_t)n * cb + cb) * 8;
    int in_smem = smem_full <= 160 * 1024;
    size_t smem = in_smem ? smem_full : (size_t)cb * 8;
    if (smem > 200 * 1024) return de6d_set_error(DE6D_ERR_INVALID, "nms: too many boxes for the sweep");
    static unsigned long long devs[3] = {0, 0, 0};
    if (mode < 0 || mode > 2) return de6d_set_error(DE6D_ERR_INVALID, "nms: unknown mode");
    {
        int rc = mode == 1 ? de6d_ensure_smem(nms_kernel<1>, 200 * 1024, devs[1], "nms smem attribute")
                 : mode == 2 ? de6d_ensure_smem(nms_kernel<2>, 200 * 1024, devs[2], "nms smem attribute")
                             : de6d_ensure_smem(nms_kernel<0>, 200 * 1024, devs[0], "nms smem attribute");
        if (rc) return rc;
    }
    dim3 grid((unsigned)tiles, frames);
    if (mode == 1) nms_kernel<1><<<grid, NMS_T, smem, stream>>>(n, boxes, nvalid, thresh, mask, tickets, keep, num_keep, in_smem);
    else if (mode == 2) nms_kernel<2><<<grid, NMS_T, smem, stream>>>(n, boxes, nvalid, thresh, mask, tickets, keep, num_keep, in_smem);
    else nms_kernel<0><<<grid, NMS_T, smem, stream>>>(n, boxes, nvalid, thresh, mask, tickets, keep, num_keep, in_smem);
    DE6D_CHECK_LAUNCH("nms_kernel");
    return DE6D_OK;
}
