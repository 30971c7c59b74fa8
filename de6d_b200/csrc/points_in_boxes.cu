// points_in_boxes for sm_100a.
//
// Replaces roiaware_pool3d/src/roiaware_pool3d_kernel.cu:16-36,313-336 (every thread recomputes cos/sin of
// every box for every point and compares in double) and offers the (T, M) mask variant of
// roiaware_pool3d.cpp:121-168 (the reference's host-side points_in_boxes_cpu) as a device kernel.
//
// Per frame the boxes are staged once in shared memory with everything that depends on the box alone
// precomputed by one thread per box: cos(-rz), sin(-rz) and the three bounds.  The reference compares
//   (double)|z-cz| > dz/2.0,  (double)|lx| < dx/2.0 + MARGIN,  (double)|ly| < dy/2.0 + MARGIN
// with a float on the left.  For a float a and a double bound h:  a > h  <=>  a > rd(h)  and
// a < h  <=>  a <= pred(h), where rd(h) is the largest float <= h and pred(h) the largest float < h.
// Those two floats are computed once per box (in double, exactly), so the per-pair work is float only and
// the decisions are identical.  The local coordinates use the contraction of the reference build:
//   lx = fmaf(sx, c, sy*s'), ly = fmaf(sy, c, -(sx*s'))  with c = cos(-rz), s' = -sin(-rz)   (device variant)
//   lx = sx*ca + sy*(-sa),   ly = sx*sa + sy*ca  with separate roundings                      (host variant)
#include "common.cuh"
#include <math.h>

namespace de6d {

struct BoxPre {
    float cx, cy, cz;
    float c, s;        // cos(-rz), -sin(-rz)   [host variant: ca = cos(-rz), sa = sin(-rz) stored as c, s]
    float fz, fx, fy;  // float images of the double bounds (see header)
};

__device__ __forceinline__ float strictly_below(double h) {  // largest float f with (double)f < h
    float f = __double2float_rd(h);
    if ((double)f < h) return f;
    return nextafterf(f, -INFINITY);
}

template <bool HOST_VARIANT>
__device__ __forceinline__ BoxPre precompute_box(const float *bx, float margin) {
    BoxPre p;
    p.cx = bx[0]; p.cy = bx[1]; p.cz = bx[2];
    const float rz = bx[6];
    const float ca = cosf(-rz), sa = sinf(-rz);
    p.c = ca;
    p.s = HOST_VARIANT ? sa : -sa;
    p.fz = __double2float_rd((double)bx[5] / 2.0);
    p.fx = strictly_below((double)bx[3] / 2.0 + (double)margin);
    p.fy = strictly_below((double)bx[4] / 2.0 + (double)margin);
    return p;
}

template <bool HOST_VARIANT>
__device__ __forceinline__ bool inside(const BoxPre &b, float x, float y, float z) {
    if (fabsf(__fsub_rn(z, b.cz)) > b.fz) return false;
    const float sx = __fsub_rn(x, b.cx), sy = __fsub_rn(y, b.cy);
    float lx, ly;
    if (HOST_VARIANT) {
        lx = __fadd_rn(__fmul_rn(sx, b.c), __fmul_rn(sy, -b.s));
        ly = __fadd_rn(__fmul_rn(sx, b.s), __fmul_rn(sy, b.c));
    } else {
        lx = __fmaf_rn(sx, b.c, __fmul_rn(sy, b.s));
        ly = __fmaf_rn(sy, b.c, -__fmul_rn(sx, b.s));
    }
    return (fabsf(lx) <= b.fx) & (fabsf(ly) <= b.fy);
}

constexpr int PIB_THREADS = 256;
constexpr int PIB_BOX_TILE = 512;

// (B, M) first containing box, untouched (-1 prefill) when none: points_in_boxes_kernel :313-336
__global__ void __launch_bounds__(PIB_THREADS)
points_in_boxes_first_kernel(int t, int m, const float *__restrict__ boxes, const float *__restrict__ pts,
                             int *__restrict__ out) {
    __shared__ BoxPre sb[PIB_BOX_TILE];
    const int bs = blockIdx.y;
    const int j = blockIdx.x * PIB_THREADS + threadIdx.x;
    boxes += (size_t)bs * t * 7;
    const bool ok = j < m;
    float x = 0.f, y = 0.f, z = 0.f;
    if (ok) {
        const float *p = pts + ((size_t)bs * m + j) * 3;
        x = p[0]; y = p[1]; z = p[2];
    }
    int found = -1;
    for (int base = 0; base < t; base += PIB_BOX_TILE) {
        const int tn = min(PIB_BOX_TILE, t - base);
        __syncthreads();
        for (int i = threadIdx.x; i < tn; i += PIB_THREADS) sb[i] = precompute_box<false>(boxes + (size_t)(base + i) * 7, 1e-5f);
        __syncthreads();
        if (ok && found < 0) {
            for (int i = 0; i < tn; ++i)
                if (inside<false>(sb[i], x, y, z)) { found = base + i; break; }
        }
        if (__syncthreads_and(!ok || found >= 0)) break;
    }
    if (ok && found >= 0) out[(size_t)bs * m + j] = found;
}

// (T, M) 0/1 mask over all boxes, MARGIN 1e-2, host arithmetic: points_in_boxes_cpu roiaware_pool3d.cpp:143-168
__global__ void __launch_bounds__(PIB_THREADS)
points_in_boxes_mask_kernel(int t, int m, const float *__restrict__ boxes, const float *__restrict__ pts,
                            int *__restrict__ out) {
    __shared__ BoxPre sb[16];
    const int k0 = blockIdx.y * 16;
    const int tn = min(16, t - k0);
    if (threadIdx.x < tn) sb[threadIdx.x] = precompute_box<true>(boxes + (size_t)(k0 + threadIdx.x) * 7, 1e-2f);
    __syncthreads();
    const int j = blockIdx.x * PIB_THREADS + threadIdx.x;
    if (j >= m) return;
    const float x = pts[(size_t)j * 3], y = pts[(size_t)j * 3 + 1], z = pts[(size_t)j * 3 + 2];
    for (int i = 0; i < tn; ++i) out[(size_t)(k0 + i) * m + j] = inside<true>(sb[i], x, y, z) ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------------------
// Full-pose (9-DoF) boxes [x, y, z, dx, dy, dz, rz, ry, rx]: Det6D's own target assignment
// (point_head_box6d_vote.py:198-209,284-286) calls box_utils.points_in_boxes3d (box_utils.py:110-124), which builds
// the 8 corners with scipy Rotation.from_euler('zyx', (rz, ry, rx)) in float64 and tests every point against a
// Delaunay triangulation of each box on the HOST, one box at a time; a later box overwrites an earlier one.
// Here: R = Rx(rx) Ry(ry) Rz(rz) (the same extrinsic z-y-x composition) once per box in double, the point is
// inside iff |R^T (p - c)| <= d / 2 on all three axes (the convex hull of the corners is exactly that box), and
// the LAST containing box wins.  Double arithmetic like the reference, so only points within ~1e-15 of a face can
// be classified differently.
// ---------------------------------------------------------------------------------------------------------
struct Box9Pre {
    double r[9];       // rotation matrix, row-major: world = R * local
    double c[3], h[3]; // centre, half extents
};

__global__ void __launch_bounds__(PIB_THREADS)
points_in_boxes9_kernel(int t, int m, const float *__restrict__ boxes, const float *__restrict__ pts,
                        long long *__restrict__ out) {
    constexpr int TILE = 64;
    __shared__ Box9Pre sb[TILE];
    const int bs = blockIdx.y;
    const int j = blockIdx.x * PIB_THREADS + threadIdx.x;
    boxes += (size_t)bs * t * 9;
    const bool ok = j < m;
    double x = 0, y = 0, z = 0;
    if (ok) {
        const float *p = pts + ((size_t)bs * m + j) * 3;
        x = p[0]; y = p[1]; z = p[2];
    }
    long long found = -1;
    for (int base = 0; base < t; base += TILE) {
        const int tn = min(TILE, t - base);
        __syncthreads();
        if (threadIdx.x < tn) {
            const float *b = boxes + (size_t)(base + threadIdx.x) * 9;
            Box9Pre q;
            const double rz = b[6], ry = b[7], rx = b[8];
            const double cz = cos(rz), sz = sin(rz), cy = cos(ry), sy = sin(ry), cx = cos(rx), sx = sin(rx);
            // Rx(rx) * Ry(ry) * Rz(rz)
            q.r[0] = cy * cz;                 q.r[1] = -cy * sz;                q.r[2] = sy;
            q.r[3] = sx * sy * cz + cx * sz;  q.r[4] = -sx * sy * sz + cx * cz; q.r[5] = -sx * cy;
            q.r[6] = -cx * sy * cz + sx * sz; q.r[7] = cx * sy * sz + sx * cz;  q.r[8] = cx * cy;
            q.c[0] = b[0]; q.c[1] = b[1]; q.c[2] = b[2];
            q.h[0] = (double)b[3] / 2.0; q.h[1] = (double)b[4] / 2.0; q.h[2] = (double)b[5] / 2.0;
            sb[threadIdx.x] = q;
        }
        __syncthreads();
        if (ok) {
            for (int i = 0; i < tn; ++i) {
                const Box9Pre &q = sb[i];
                const double dx = x - q.c[0], dy = y - q.c[1], dz = z - q.c[2];
                const double lx = q.r[0] * dx + q.r[3] * dy + q.r[6] * dz;   // R^T (p - c)
                const double ly = q.r[1] * dx + q.r[4] * dy + q.r[7] * dz;
                const double lz = q.r[2] * dx + q.r[5] * dy + q.r[8] * dz;
                if (fabs(lx) <= q.h[0] && fabs(ly) <= q.h[1] && fabs(lz) <= q.h[2]) found = base + i;
            }
        }
    }
    if (ok) out[(size_t)bs * m + j] = found;
}

}  // namespace de6d

using namespace de6d;

// box_utils.points_in_boxes3d (pcdet/utils/box_utils.py:110-124), batched and on the device: boxes (b,t,9)
// [x,y,z,dx,dy,dz,rz,ry,rx], pts (b,m,3) -> out (b,m) int64 = index of the LAST box containing the point, -1 if none.
extern "C" int de6d_points_in_boxes9(int b, int t, int m, const float *boxes, const float *pts, long long *out,
                                     cudaStream_t stream) {
    if (b < 0 || t < 0 || m < 0) return de6d_set_error(DE6D_ERR_INVALID, "points_in_boxes9: negative size");
    if (b == 0 || m == 0) return DE6D_OK;
    if (!pts || !out || (t > 0 && !boxes)) return de6d_set_error(DE6D_ERR_INVALID, "points_in_boxes9: null pointer");
    if (b > 65535) return de6d_set_error(DE6D_ERR_INVALID, "points_in_boxes9: batch > 65535");
    dim3 grid(ceil_div(m, PIB_THREADS), b);
    points_in_boxes9_kernel<<<grid, PIB_THREADS, 0, stream>>>(t, m, boxes, pts, out);
    DE6D_CHECK_LAUNCH("points_in_boxes9_kernel");
    return DE6D_OK;
}

extern "C" int de6d_points_in_boxes(int b, int t, int m, const float *boxes, const float *pts, int *box_idx_of_points,
                                    cudaStream_t stream) {
    if (b < 0 || t < 0 || m < 0) return de6d_set_error(DE6D_ERR_INVALID, "points_in_boxes: negative size");
    if (b == 0 || m == 0 || t == 0) return DE6D_OK;
    if (!boxes || !pts || !box_idx_of_points) return de6d_set_error(DE6D_ERR_INVALID, "points_in_boxes: null pointer");
    if (b > 65535) return de6d_set_error(DE6D_ERR_INVALID, "points_in_boxes: batch > 65535");
    dim3 grid(ceil_div(m, PIB_THREADS), b);
    points_in_boxes_first_kernel<<<grid, PIB_THREADS, 0, stream>>>(t, m, boxes, pts, box_idx_of_points);
    DE6D_CHECK_LAUNCH("points_in_boxes_first_kernel");
    return DE6D_OK;
}

extern "C" int de6d_points_in_boxes_mask(int t, int m, const float *boxes, const float *pts, int *point_indices,
                                         cudaStream_t stream) {
    if (t < 0 || m < 0) return de6d_set_error(DE6D_ERR_INVALID, "points_in_boxes_mask: negative size");
    if (m == 0 || t == 0) return DE6D_OK;
    if (!boxes || !pts || !point_indices) return de6d_set_error(DE6D_ERR_INVALID, "points_in_boxes_mask: null pointer");
    if (ceil_div(t, 16) > 65535) return de6d_set_error(DE6D_ERR_INVALID, "points_in_boxes_mask: too many boxes");
    dim3 grid(ceil_div(m, PIB_THREADS), ceil_div(t, 16));
    points_in_boxes_mask_kernel<<<grid, PIB_THREADS, 0, stream>>>(t, m, boxes, pts, point_indices);
    DE6D_CHECK_LAUNCH("points_in_boxes_mask_kernel");
    return DE6D_OK;
}
