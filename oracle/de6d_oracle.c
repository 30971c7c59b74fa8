/*
 * de6d_oracle.c -- CPU restatement of the reference's point-set-abstraction and
 * box-op kernels.  TEST INFRASTRUCTURE ONLY: linked/loaded solely by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 * The product package (de6d_b200/) never imports, links or executes this file.
 *
 * Parity status: PINNED.  Every function here is checked (tests/test_oracle_*.py,
 * tests/golden/) against outputs of the reference's own kernels, produced by
 * running the unmodified reference sources compiled by oracle/build_ref.py:
 * the two CPU entry points in this container, the CUDA ones on a B200 box
 * (tests/golden/make_golden.py is the generating script).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/core/pcdet/ops/).  Arithmetic notes:
 *   - build with -ffp-contract=off; every fused multiply-add the reference's
 *     nvcc build performs is spelled fmaf() here (shapes read off the SASS of
 *     the oracle/_ref build, nvcc 12.9 default -fmad=true):
 *       squared distance  d = fmaf(dz,dz, fmaf(dx,dx, dy*dy))
 *       interpolation     o = fmaf(w2,p2, fmaf(w0,p0, w1*p1))
 *       in-box rotation   lx = fmaf(sx,c, sy*s) ; ly = fmaf(sy,c, -(sx*s))
 *   - hidden double arithmetic is kept (S-FPS key, points_in_boxes bounds).
 *   - the rotated-IoU tree is restated operation by operation in float without
 *     contraction; the reference GPU build contracts some of it, so IoU values
 *     agree to ~1e-6 relative, not bitwise (tolerance 1e-5).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ---- block size rule: pointnet2/pointnet2_batch/src/cuda_utils.h:10-14 ---- */
ORC_API int orc_opt_n_threads(int work_size) {
    int p = (int)(log((double)work_size) / log(2.0));
    int v = 1 << p;
    if (v > 1024) v = 1024;
    if (v < 1) v = 1;
    return v;
}

static inline float sqdist(float ax, float ay, float az, float bx, float by, float bz) {
    /* (a-b) per axis; y term is the plain rounded product (SASS FMUL), x and z are FFMA */
    float dx = ax - bx, dy = ay - by, dz = az - bz;
    return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/* shared-memory tree of sampling_gpu.cu:94-99,155-215: lower slot survives ties */
static int tree_argmax(float *v, int *vi, int bs) {
    for (int half = bs >> 1; half >= 1; half >>= 1) {
        for (int t = 0; t < half; ++t) {
            float v1 = v[t], v2 = v[t + half];
            int i1 = vi[t], i2 = vi[t + half];
            v[t] = fmaxf(v1, v2);
            vi[t] = v2 > v1 ? i2 : i1;
        }
    }
    return vi[0];
}

/* D-FPS: sampling_gpu.cu:101-222 (launcher :224-266) */
ORC_API void orc_fps(int b, int n, int m, const float *xyz, float *temp, int *idx) {
    if (m <= 0) return;
    int bs = orc_opt_n_threads(n);
    float *best = (float *)malloc(sizeof(float) * bs);
    int *besti = (int *)malloc(sizeof(int) * bs);
    for (int bi = 0; bi < b; ++bi) {
        const float *p = xyz + (size_t)bi * n * 3;
        float *t = temp + (size_t)bi * n;
        int *out = idx + (size_t)bi * m;
        int old = 0;
        out[0] = 0;
        for (int j = 1; j < m; ++j) {
            for (int s = 0; s < bs; ++s) { best[s] = -1.0f; besti[s] = 0; }
            float x1 = p[old * 3], y1 = p[old * 3 + 1], z1 = p[old * 3 + 2];
            for (int k = 0; k < n; ++k) {           /* slot = k mod bs visits k ascending, as thread `slot` does */
                float d = sqdist(p[k * 3], p[k * 3 + 1], p[k * 3 + 2], x1, y1, z1);
                float d2 = fminf(d, t[k]);
                t[k] = d2;
                int s = k & (bs - 1);
                if (d2 > best[s]) { best[s] = d2; besti[s] = k; }
            }
            old = tree_argmax(best, besti, bs);
            out[j] = old;
        }
    }
    free(best); free(besti);
}

/* F-FPS on a precomputed matrix: sampling_gpu.cu:268-373 */
ORC_API void orc_fps_matrix(int b, int n, int m, const float *matrix, float *temp, int *idx) {
    if (m <= 0) return;
    int bs = orc_opt_n_threads(n);
    float *best = (float *)malloc(sizeof(float) * bs);
    int *besti = (int *)malloc(sizeof(int) * bs);
    for (int bi = 0; bi < b; ++bi) {
        const float *mat = matrix + (size_t)bi * n * n;
        float *t = temp + (size_t)bi * n;
        int *out = idx + (size_t)bi * m;
        int old = 0;
        out[0] = 0;
        for (int j = 1; j < m; ++j) {
            for (int s = 0; s < bs; ++s) { best[s] = -1.0f; besti[s] = 0; }
            const float *row = mat + (size_t)old * n;
            for (int k = 0; k < n; ++k) {
                float d2 = fminf(row[k], t[k]);
                t[k] = d2;
                int s = k & (bs - 1);
                if (d2 > best[s]) { best[s] = d2; besti[s] = k; }
            }
            old = tree_argmax(best, besti, bs);
            out[j] = old;
        }
    }
    free(best); free(besti);
}

/* S-FPS: sampling_gpu.cu:419-540.  Iteration 0 = argmax(weights); key evaluated in double (:465) */
ORC_API void orc_fps_weights(int b, int n, int m, const float *xyz, const float *weights, float *temp, int *idx) {
    if (m <= 0) return;
    int bs = orc_opt_n_threads(n);
    float *best = (float *)malloc(sizeof(float) * bs);
    int *besti = (int *)malloc(sizeof(int) * bs);
    for (int bi = 0; bi < b; ++bi) {
        const float *p = xyz + (size_t)bi * n * 3;
        const float *w = weights + (size_t)bi * n;
        float *t = temp + (size_t)bi * n;
        int *out = idx + (size_t)bi * m;
        int old = 0;
        for (int j = 0; j < m; ++j) {
            for (int s = 0; s < bs; ++s) { best[s] = -1.0f; besti[s] = 0; }
            float x1 = p[old * 3], y1 = p[old * 3 + 1], z1 = p[old * 3 + 2];
            for (int k = 0; k < n; ++k) {
                int s = k & (bs - 1);
                float key;
                if (j == 0) {
                    key = w[k];
                } else {
                    float d = sqdist(p[k * 3], p[k * 3 + 1], p[k * 3 + 2], x1, y1, z1);
                    d = fminf(d, t[k]);
                    t[k] = d;
                    key = (float)((double)d * fmax((double)w[k], 1e-12));
                }
                if (key > best[s]) { best[s] = key; besti[s] = k; }
            }
            old = tree_argmax(best, besti, bs);
            out[j] = old;
        }
    }
    free(best); free(besti);
}

/* gather_points: sampling_gpu.cu:16-32 ; grad :54-71 */
ORC_API void orc_gather_points(int b, int c, int n, int m, const float *points, const int *idx, float *out) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *src = points + ((size_t)bi * c + ci) * n;
            float *dst = out + ((size_t)bi * c + ci) * m;
            const int *ix = idx + (size_t)bi * m;
            for (int j = 0; j < m; ++j) dst[j] = src[ix[j]];
        }
}

ORC_API void orc_gather_points_grad(int b, int c, int n, int m, const float *grad_out, const int *idx, float *grad_points) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *g = grad_out + ((size_t)bi * c + ci) * m;
            float *dst = grad_points + ((size_t)bi * c + ci) * n;
            const int *ix = idx + (size_t)bi * m;
            for (int j = 0; j < m; ++j) dst[ix[j]] += g[j];
        }
}

/* ball_query: ball_query_gpu.cu:15-51 (pads with the first hit; untouched row when empty) */
ORC_API void orc_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz, int *idx) {
    float r2 = radius * radius;
    for (int bi = 0; bi < b; ++bi)
        for (int q = 0; q < m; ++q) {
            const float *c = new_xyz + ((size_t)bi * m + q) * 3;
            const float *p = xyz + (size_t)bi * n * 3;
            int *row = idx + ((size_t)bi * m + q) * nsample;
            int cnt = 0;
            for (int k = 0; k < n && cnt < nsample; ++k) {
                float d2 = sqdist(c[0], c[1], c[2], p[k * 3], p[k * 3 + 1], p[k * 3 + 2]);
                if (d2 < r2) {
                    if (cnt == 0) for (int l = 0; l < nsample; ++l) row[l] = k;
                    row[cnt++] = k;
                }
            }
        }
}

/* cyclic padding of ball_query_gpu.cu:87-90,126-129 (reads what it has just written) */
static void cyclic_pad(int *row, int cnt, int nsample) {
    for (int l = 0; cnt < nsample; ++l, ++cnt) row[cnt] = row[l];
}

/* ball_query_cnt: ball_query_gpu.cu:93-130 */
ORC_API void orc_ball_query_cnt(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz,
                                int *idx_cnt, int *idx) {
    float r2 = radius * radius;
    for (int bi = 0; bi < b; ++bi)
        for (int q = 0; q < m; ++q) {
            const float *c = new_xyz + ((size_t)bi * m + q) * 3;
            const float *p = xyz + (size_t)bi * n * 3;
            int *row = idx + ((size_t)bi * m + q) * nsample;
            int cnt = 0;
            for (int k = 0; k < n && cnt < nsample; ++k) {
                float d2 = sqdist(c[0], c[1], c[2], p[k * 3], p[k * 3 + 1], p[k * 3 + 2]);
                if (d2 < r2) row[cnt++] = k;
            }
            idx_cnt[(size_t)bi * m + q] = cnt;
            cyclic_pad(row, cnt, nsample);
        }
}

/* ball_query_dilated: ball_query_gpu.cu:53-91 */
ORC_API void orc_ball_query_dilated(int b, int n, int m, float radius_in, float radius_out, int nsample,
                                    const float *new_xyz, const float *xyz, int *idx_cnt, int *idx) {
    float rin2 = radius_in * radius_in, rout2 = radius_out * radius_out;
    for (int bi = 0; bi < b; ++bi)
        for (int q = 0; q < m; ++q) {
            const float *c = new_xyz + ((size_t)bi * m + q) * 3;
            const float *p = xyz + (size_t)bi * n * 3;
            int *row = idx + ((size_t)bi * m + q) * nsample;
            int cnt = 0;
            for (int k = 0; k < n && cnt < nsample; ++k) {
                float d2 = sqdist(c[0], c[1], c[2], p[k * 3], p[k * 3 + 1], p[k * 3 + 2]);
                if (d2 >= rin2 && d2 < rout2) row[cnt++] = k;
            }
            idx_cnt[(size_t)bi * m + q] = cnt;
            cyclic_pad(row, cnt, nsample);
        }
}

/* group_points: group_points_gpu.cu:53-72 ; grad :14-31 */
ORC_API void orc_group_points(int b, int c, int n, int npoints, int nsample, const float *points, const int *idx, float *out) {
    size_t ms = (size_t)npoints * nsample;
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *src = points + ((size_t)bi * c + ci) * n;
            float *dst = out + ((size_t)bi * c + ci) * ms;
            const int *ix = idx + (size_t)bi * ms;
            for (size_t j = 0; j < ms; ++j) dst[j] = src[ix[j]];
        }
}

ORC_API void orc_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out, const int *idx,
                                   float *grad_points) {
    size_t ms = (size_t)npoints * nsample;
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *g = grad_out + ((size_t)bi * c + ci) * ms;
            float *dst = grad_points + ((size_t)bi * c + ci) * n;
            const int *ix = idx + (size_t)bi * ms;
            for (size_t j = 0; j < ms; ++j) dst[ix[j]] += g[j];
        }
}

/* F-FPS distance matrix: calc_dist_matrix_for_sampling, pointnet2/pointnet2_batch/pointnet2_utils.py:36-44
 *   dist = cdist(xyz, xyz) + cdist(features, features) * gamma
 * torch.cdist is a third-party routine (torch 2.11.0; for N > 25 it expands |a|^2+|b|^2-2ab through a GEMM, whose
 * summation order is not specified), so this is NOT a bitwise restatement of torch: it is the mathematical
 * definition evaluated with direct differences in a fixed order, which de6d_b200's dist-matrix kernel reproduces
 * bit for bit.  Against torch.cdist the two agree to ~1e-3 absolute (torch's expansion cancels for close
 * points); the matrix is an INPUT of the F-FPS kernel under test, so this does not enter index parity.
 * features: (b, n, c) contiguous or NULL. */
ORC_API void orc_dist_matrix(int b, int n, int c, const float *xyz, const float *features, float gamma, float *out) {
    for (int bi = 0; bi < b; ++bi) {
        const float *x = xyz + (size_t)bi * n * 3;
        const float *f = features ? features + (size_t)bi * n * c : 0;
        float *o = out + (size_t)bi * n * n;
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                float d1 = sqrtf(sqdist(x[i * 3], x[i * 3 + 1], x[i * 3 + 2], x[j * 3], x[j * 3 + 1], x[j * 3 + 2]));
                if (f) {
                    float acc = 0.f;
                    for (int ch = 0; ch < c; ++ch) {
                        float t = f[(size_t)i * c + ch] - f[(size_t)j * c + ch];
                        acc = fmaf(t, t, acc);
                    }
                    float g = sqrtf(acc) * gamma;
                    d1 = d1 + g;
                }
                o[(size_t)i * n + j] = d1;
            }
    }
}

/* three_nn: interpolate_gpu.cu:16-59.  Running minima are double there; every value stored in them is a
 * float (or the 1e40 sentinel, which converts to +inf on the final store), so the compares are restated in double. */
ORC_API void orc_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx) {
    for (int bi = 0; bi < b; ++bi)
        for (int q = 0; q < n; ++q) {
            const float *u = unknown + ((size_t)bi * n + q) * 3;
            const float *kn = known + (size_t)bi * m * 3;
            double b1 = 1e40, b2 = 1e40, b3 = 1e40;
            int i1 = 0, i2 = 0, i3 = 0;
            for (int k = 0; k < m; ++k) {
                double d = (double)sqdist(u[0], u[1], u[2], kn[k * 3], kn[k * 3 + 1], kn[k * 3 + 2]);
                if (d < b1) { b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k; }
                else if (d < b2) { b3 = b2; i3 = i2; b2 = d; i2 = k; }
                else if (d < b3) { b3 = d; i3 = k; }
            }
            float *dd = dist2 + ((size_t)bi * n + q) * 3;
            int *ii = idx + ((size_t)bi * n + q) * 3;
            dd[0] = (float)b1; dd[1] = (float)b2; dd[2] = (float)b3;
            ii[0] = i1; ii[1] = i2; ii[2] = i3;
        }
}

/* three_interpolate: interpolate_gpu.cu:84-104 ; grad :127-149 */
ORC_API void orc_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx, const float *weight, float *out) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *src = points + ((size_t)bi * c + ci) * m;
            float *dst = out + ((size_t)bi * c + ci) * n;
            for (int q = 0; q < n; ++q) {
                const float *w = weight + ((size_t)bi * n + q) * 3;
                const int *ix = idx + ((size_t)bi * n + q) * 3;
                dst[q] = fmaf(w[2], src[ix[2]], fmaf(w[0], src[ix[0]], w[1] * src[ix[1]]));
            }
        }
}

ORC_API void orc_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx, const float *weight,
                                        float *grad_points) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *g = grad_out + ((size_t)bi * c + ci) * n;
            float *dst = grad_points + ((size_t)bi * c + ci) * m;
            for (int q = 0; q < n; ++q) {
                const float *w = weight + ((size_t)bi * n + q) * 3;
                const int *ix = idx + ((size_t)bi * n + q) * 3;
                dst[ix[0]] += g[q] * w[0];
                dst[ix[1]] += g[q] * w[1];
                dst[ix[2]] += g[q] * w[2];
            }
        }
}

/* ------------------------------------------------------------------------------------------------
 * Rotated BEV overlap: iou3d_nms/src/iou3d_nms_kernel.cu:15-234 (CPU twin iou3d_cpu.cpp:39-229)
 * ---------------------------------------------------------------------------------------------- */
typedef struct { float x, y; } P2;
#define IOU_EPS 1e-8f

static inline float cross3(P2 p1, P2 p2, P2 p0) { /* :39-41 */
    return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y);
}
static inline float cross2(P2 a, P2 b) { return a.x * b.y - a.y * b.x; } /* :35-37 */

static int bbox_overlap_1d(P2 p1, P2 p2, P2 q1, P2 q2) { /* check_rect_cross :43-49 */
    return fminf(p1.x, p2.x) <= fmaxf(q1.x, q2.x) && fminf(q1.x, q2.x) <= fmaxf(p1.x, p2.x) &&
           fminf(p1.y, p2.y) <= fmaxf(q1.y, q2.y) && fminf(q1.y, q2.y) <= fmaxf(p1.y, p2.y);
}

static int corner_in_box(const float *box, P2 p) { /* check_in_box2d :51-61, MARGIN 1e-2 */
    const float margin = 1e-2f;
    float c = cosf(-box[6]), s = sinf(-box[6]);
    float rx = (p.x - box[0]) * c + (p.y - box[1]) * (-s);
    float ry = (p.x - box[0]) * s + (p.y - box[1]) * c;
    return fabsf(rx) < box[3] / 2 + margin && fabsf(ry) < box[4] / 2 + margin;
}

static int seg_intersect(P2 p1, P2 p0, P2 q1, P2 q0, P2 *ans) { /* intersection :63-92 */
    if (!bbox_overlap_1d(p0, p1, q0, q1)) return 0;
    float s1 = cross3(q0, p1, p0), s2 = cross3(p1, q1, p0), s3 = cross3(p0, q1, q0), s4 = cross3(q1, p1, q0);
    if (!(s1 * s2 > 0 && s3 * s4 > 0)) return 0;
    float s5 = cross3(q1, p1, p0);
    if (fabsf(s5 - s1) > IOU_EPS) {
        ans->x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
        ans->y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
    } else {
        float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
        float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
        float D = a0 * b1 - a1 * b0;
        ans->x = (b0 * c1 - b1 * c0) / D;
        ans->y = (a1 * c0 - a0 * c1) / D;
    }
    return 1;
}

static void box_corners(const float *box, P2 *c5) { /* :108-145 incl. rotate_around_center :94-98 */
    float hx = box[3] / 2, hy = box[4] / 2;
    float x1 = box[0] - hx, y1 = box[1] - hy, x2 = box[0] + hx, y2 = box[1] + hy;
    float co = cosf(box[6]), si = sinf(box[6]);
    float px[4] = {x1, x2, x2, x1}, py[4] = {y1, y1, y2, y2};
    for (int k = 0; k < 4; ++k) {
        float ox = px[k] - box[0], oy = py[k] - box[1];
        c5[k].x = ox * co + oy * (-si) + box[0];
        c5[k].y = ox * si + oy * co + box[1];
    }
    c5[4] = c5[0];
}

ORC_API float orc_box_overlap(const float *a, const float *b) { /* box_overlap :104-225 */
    P2 ca[5], cb[5], pts[16], ctr = {0.f, 0.f};
    box_corners(a, ca);
    box_corners(b, cb);
    int cnt = 0;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            if (seg_intersect(ca[i + 1], ca[i], cb[j + 1], cb[j], &pts[cnt])) {
                ctr.x += pts[cnt].x; ctr.y += pts[cnt].y; ++cnt;
            }
    for (int k = 0; k < 4; ++k) {
        if (corner_in_box(a, cb[k])) { ctr.x += cb[k].x; ctr.y += cb[k].y; pts[cnt++] = cb[k]; }
        if (corner_in_box(b, ca[k])) { ctr.x += ca[k].x; ctr.y += ca[k].y; pts[cnt++] = ca[k]; }
    }
    ctr.x /= cnt; ctr.y /= cnt;
    for (int j = 0; j < cnt - 1; ++j)              /* bubble sort by polar angle about the centroid :199-209 */
        for (int i = 0; i < cnt - j - 1; ++i)
            if (atan2f(pts[i].y - ctr.y, pts[i].x - ctr.x) > atan2f(pts[i + 1].y - ctr.y, pts[i + 1].x - ctr.x)) {
                P2 t = pts[i]; pts[i] = pts[i + 1]; pts[i + 1] = t;
            }
    float area = 0.f;
    for (int k = 0; k < cnt - 1; ++k) {
        P2 u = {pts[k].x - pts[0].x, pts[k].y - pts[0].y}, v = {pts[k + 1].x - pts[0].x, pts[k + 1].y - pts[0].y};
        area += cross2(u, v);
    }
    return fabsf(area) / 2.0f;
}

ORC_API float orc_iou_bev(const float *a, const float *b) { /* iou_bev :227-234 */
    float sa = a[3] * a[4], sb = b[3] * b[4], s = orc_box_overlap(a, b);
    return s / fmaxf(sa + sb - s, IOU_EPS);
}

/* boxes_overlap_kernel :236-249 / boxes_iou_bev_kernel :251-265 / boxes_iou_bev_cpu iou3d_cpu.cpp:232-252 */
ORC_API void orc_boxes_overlap_bev(int na, const float *a, int nb, const float *b, float *out) {
    for (int i = 0; i < na; ++i) for (int j = 0; j < nb; ++j) out[(size_t)i * nb + j] = orc_box_overlap(a + i * 7, b + j * 7);
}
ORC_API void orc_boxes_iou_bev(int na, const float *a, int nb, const float *b, float *out) {
    for (int i = 0; i < na; ++i) for (int j = 0; j < nb; ++j) out[(size_t)i * nb + j] = orc_iou_bev(a + i * 7, b + j * 7);
}

/* boxes_iou3d_gpu: iou3d_nms_utils.py:48-81 -- each torch op rounds to float separately */
ORC_API void orc_boxes_iou3d(int na, const float *a, int nb, const float *b, float *out) {
    for (int i = 0; i < na; ++i) {
        const float *A = a + i * 7;
        float amax = A[2] + A[5] / 2, amin = A[2] - A[5] / 2, va = A[3] * A[4] * A[5];
        for (int j = 0; j < nb; ++j) {
            const float *B = b + j * 7;
            float bmax = B[2] + B[5] / 2, bmin = B[2] - B[5] / 2, vb = B[3] * B[4] * B[5];
            float bev = orc_box_overlap(A, B);
            float h = fminf(amax, bmax) - fmaxf(amin, bmin);
            if (h < 0.f) h = 0.f;
            float o3 = bev * h;
            float den = va + vb - o3;
            if (den < 1e-6f) den = 1e-6f;
            out[(size_t)i * nb + j] = o3 / den;
        }
    }
}

static float iou_axis_aligned(const float *a, const float *b) { /* iou_normal :314-325 */
    float left = fmaxf(a[0] - a[3] / 2, b[0] - b[3] / 2), right = fminf(a[0] + a[3] / 2, b[0] + b[3] / 2);
    float top = fmaxf(a[1] - a[4] / 2, b[1] - b[4] / 2), bottom = fminf(a[1] + a[4] / 2, b[1] + b[4] / 2);
    float w = fmaxf(right - left, 0.f), h = fmaxf(bottom - top, 0.f);
    float inter = w * h, sa = a[3] * a[4], sb = b[3] * b[4];
    return inter / fmaxf(sa + sb - inter, IOU_EPS);
}

/* nms_kernel :267-311 (bit i of word [row][colblk] iff IoU(row, col) > thresh, diagonal tile only col > row)
 * + host sweep iou3d_nms.cpp:113-132.  normal=1 -> nms_normal_kernel :328-372.  Returns num_to_keep. */
ORC_API int orc_nms(int n, const float *boxes, float thresh, int normal, int64_t *keep, uint64_t *mask_out) {
    int cb = (n + 63) / 64;
    uint64_t *mask = mask_out ? mask_out : (uint64_t *)malloc(sizeof(uint64_t) * (size_t)n * cb + 8);
    for (int i = 0; i < n; ++i)
        for (int c = 0; c < cb; ++c) {
            uint64_t t = 0;
            int cs = n - c * 64 < 64 ? n - c * 64 : 64;
            int start = (i / 64 == c) ? (i % 64) + 1 : 0;
            for (int k = start; k < cs; ++k) {
                const float *bi = boxes + (size_t)i * 7, *bj = boxes + (size_t)(c * 64 + k) * 7;
                float v = normal ? iou_axis_aligned(bi, bj) : orc_iou_bev(bi, bj);
                if (v > thresh) t |= 1ULL << k;
            }
            mask[(size_t)i * cb + c] = t;
        }
    uint64_t *remv = (uint64_t *)calloc(cb > 0 ? cb : 1, sizeof(uint64_t));
    int nk = 0;
    for (int i = 0; i < n; ++i) {
        int nb = i / 64, ib = i % 64;
        if (!(remv[nb] & (1ULL << ib))) {
            keep[nk++] = i;
            for (int j = nb; j < cb; ++j) remv[j] |= mask[(size_t)i * cb + j];
        }
    }
    free(remv);
    if (!mask_out) free(mask);
    return nk;
}

/* check_pt_in_box3d: roiaware_pool3d/src/roiaware_pool3d_kernel.cu:16-36 (CPU twin roiaware_pool3d.cpp:121-140).
 * Bounds compared in double (dz / 2.0, dx / 2.0 + MARGIN with MARGIN a float constant promoted to double). */
static int pt_in_box(const float *pt, const float *bx, float margin) {
    float x = pt[0], y = pt[1], z = pt[2];
    float cx = bx[0], cy = bx[1], cz = bx[2], dx = bx[3], dy = bx[4], dz = bx[5], rz = bx[6];
    if ((double)fabsf(z - cz) > (double)dz / 2.0) return 0;
    float c = cosf(rz), s = sinf(rz);            /* cos(-rz) = c, sin(-rz) = -s */
    float sx = x - cx, sy = y - cy;
    float lx = fmaf(sx, c, sy * s);
    float ly = fmaf(sy, c, -(sx * s));
    return ((double)fabsf(lx) < (double)dx / 2.0 + (double)margin) & ((double)fabsf(ly) < (double)dy / 2.0 + (double)margin);
}

/* points_in_boxes_kernel :313-336 -- first containing box, else the caller's prefill (-1) */
ORC_API void orc_points_in_boxes_gpu(int b, int t, int m, const float *boxes, const float *pts, int *out) {
    for (int bi = 0; bi < b; ++bi)
        for (int j = 0; j < m; ++j) {
            const float *p = pts + ((size_t)bi * m + j) * 3;
            for (int k = 0; k < t; ++k)
                if (pt_in_box(p, boxes + ((size_t)bi * t + k) * 7, 1e-5f)) { out[(size_t)bi * m + j] = k; break; }
        }
}

/* points_in_boxes_cpu: roiaware_pool3d.cpp:143-168 -- (T, M) 0/1 mask, MARGIN 1e-2, unfused host arithmetic */
static int pt_in_box_host(const float *pt, const float *bx) {
    const float margin = 1e-2f;
    float x = pt[0], y = pt[1], z = pt[2];
    float cx = bx[0], cy = bx[1], cz = bx[2], dx = bx[3], dy = bx[4], dz = bx[5], rz = bx[6];
    if ((double)fabsf(z - cz) > (double)dz / 2.0) return 0;
    float ca = cosf(-rz), sa = sinf(-rz);
    float sx = x - cx, sy = y - cy;
    float lx = sx * ca + sy * (-sa);
    float ly = sx * sa + sy * ca;
    return ((double)fabsf(lx) < (double)dx / 2.0 + (double)margin) & ((double)fabsf(ly) < (double)dy / 2.0 + (double)margin);
}

ORC_API void orc_points_in_boxes_cpu(int t, int m, const float *boxes, const float *pts, int *out) {
    for (int k = 0; k < t; ++k)
        for (int j = 0; j < m; ++j) out[(size_t)k * m + j] = pt_in_box_host(pts + (size_t)j * 3, boxes + (size_t)k * 7);
}

/* box_utils.points_in_boxes3d: pcdet/utils/box_utils.py:59-72 (corners from scipy Rotation.from_euler('zyx', (rz, ry, rx)),
 * float64) + :110-124 (Delaunay in_hull per box, later boxes overwrite).  Third-party pieces: scipy 1.x Rotation
 * (extrinsic z, then y, then x: R = Rx Ry Rz) and Delaunay.find_simplex >= 0, i.e. membership in the convex hull of the
 * 8 corners -- the box itself -- so the test is restated as |R^T (p - c)| <= d / 2 in double.  tests/test_oracle.py
 * pins this against scipy's own Rotation + Delaunay on random points away from faces.
 * boxes (t, 9) [x,y,z,dx,dy,dz,rz,ry,rx], pts (m, 3) -> flags (m) int64. */
ORC_API void orc_points_in_boxes9(int t, int m, const float *boxes, const float *pts, long long *flags) {
    for (int j = 0; j < m; ++j) flags[j] = -1;
    for (int i = 0; i < t; ++i) {
        const float *b = boxes + (size_t)i * 9;
        const double rz = b[6], ry = b[7], rx = b[8];
        const double cz = cos(rz), sz = sin(rz), cy = cos(ry), sy = sin(ry), cx = cos(rx), sx = sin(rx);
        const double r[9] = {cy * cz, -cy * sz, sy,
                             sx * sy * cz + cx * sz, -sx * sy * sz + cx * cz, -sx * cy,
                             -cx * sy * cz + sx * sz, cx * sy * sz + sx * cz, cx * cy};
        const double hx = (double)b[3] / 2.0, hy = (double)b[4] / 2.0, hz = (double)b[5] / 2.0;
        for (int j = 0; j < m; ++j) {
            const double dx = (double)pts[j * 3] - b[0], dy = (double)pts[j * 3 + 1] - b[1], dz = (double)pts[j * 3 + 2] - b[2];
            const double lx = r[0] * dx + r[3] * dy + r[6] * dz;
            const double ly = r[1] * dx + r[4] * dy + r[7] * dz;
            const double lz = r[2] * dx + r[5] * dy + r[8] * dz;
            if (fabs(lx) <= hx && fabs(ly) <= hy && fabs(lz) <= hz) flags[j] = i;
        }
    }
}
