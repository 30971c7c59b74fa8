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

/* ------------------------------------------------------------------------------------------------
 * Full-pose (9-DoF) box intersection volume / IoU / NMS -- SURVEY.md 8(f) rank 3, second half.
 * The reference has NO such function: PointHeadBox6DVote calls boxes_iou3d_gpu on boxes[:, 0:7]
 * (point_head_box6d_vote.py:355) and class_agnostic_nms slices [:, 0:7] (model_nms_utils.py:18), i.e. pitch and roll
 * are ignored.  Boxes follow box_utils.boxes3d_to_corners_3d (pcdet/utils/box_utils.py:59-72):
 * [x, y, z, dx, dy, dz, rz, ry, rx], corners = R * (+-d/2) + centre with R = Rotation.from_euler('zyx', (rz, ry, rx))
 * = Rx Ry Rz (the matrix orc_points_in_boxes9 uses, pinned against scipy there).
 * Parity: pinned in tests/test_oracle.py against scipy.spatial (HalfspaceIntersection + ConvexHull volume, float64) and,
 * for ry = rx = 0, against the reference's own boxes_iou3d_gpu composition (orc_boxes_iou3d).
 *
 * Algorithm (all double): work in box A's frame (A = [-ha, ha]^3 axis aligned) and clip the polyhedron A, kept as a list of
 * convex face polygons, by B's six half-spaces one after the other (Sutherland-Hodgman per face).  The points where edges
 * cross the clipping plane are collected -- computed from the inside vertex towards the outside vertex, so the two faces
 * sharing an edge produce bit-identical points -- ordered by angle about their centroid and become the cap face on that
 * plane.  By the divergence theorem V = 1/3 * sum over faces of (plane offset from A's centre) * (face area).  Because the
 * cap is built from the very cut points that removed material, coplanar or nearly coplanar faces (identical boxes, padded
 * duplicates, boxes sharing a ground plane) cost O(rounding), not a face counted twice.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { double x, y, z; } V3;
#define B9_MAXV 16
#define B9_MAXC 32

static void box9_rot(const float *b, double r[9]) {
    const double rz = b[6], ry = b[7], rx = b[8];
    const double cz = cos(rz), sz = sin(rz), cy = cos(ry), sy = sin(ry), cx = cos(rx), sx = sin(rx);
    r[0] = cy * cz;                 r[1] = -cy * sz;                r[2] = sy;
    r[3] = sx * sy * cz + cx * sz;  r[4] = -sx * sy * sz + cx * cz; r[5] = -sx * cy;
    r[6] = -cx * sy * cz + sx * sz; r[7] = cx * sy * sz + sx * cz;  r[8] = cx * cy;
}

static double poly_area(const V3 *p, int n) {
    if (n < 3) return 0.0;
    double ax = 0, ay = 0, az = 0;
    for (int i = 1; i + 1 < n; ++i) {
        const V3 u = {p[i].x - p[0].x, p[i].y - p[0].y, p[i].z - p[0].z}, v = {p[i + 1].x - p[0].x, p[i + 1].y - p[0].y, p[i + 1].z - p[0].z};
        ax += u.y * v.z - u.z * v.y; ay += u.z * v.x - u.x * v.z; az += u.x * v.y - u.y * v.x;
    }
    return 0.5 * sqrt(ax * ax + ay * ay + az * az);
}

/* clip one face polygon (in place) by n.p <= d; cut points are appended to cp */
static int clip_face(V3 *poly, int n, V3 nrm, double d, V3 *cp, int *ncp) {
    V3 out[B9_MAXV];
    double s[B9_MAXV];
    int m = 0;
    for (int i = 0; i < n; ++i) s[i] = nrm.x * poly[i].x + nrm.y * poly[i].y + nrm.z * poly[i].z - d;
    for (int i = 0; i < n; ++i) {
        const int k = (i + 1) % n;
        const int pin = s[i] <= 0.0, qin = s[k] <= 0.0;
        if (pin && m < B9_MAXV) out[m++] = poly[i];
        if (pin != qin) {
            const V3 a = pin ? poly[i] : poly[k], b = pin ? poly[k] : poly[i];   /* inside -> outside: same bits from both faces */
            const double sa = pin ? s[i] : s[k], sb = pin ? s[k] : s[i];
            const double t = sa / (sa - sb);
            const V3 x = {a.x + t * (b.x - a.x), a.y + t * (b.y - a.y), a.z + t * (b.z - a.z)};
            if (m < B9_MAXV) out[m++] = x;
            if (*ncp < B9_MAXC) cp[(*ncp)++] = x;
        }
    }
    for (int i = 0; i < m; ++i) poly[i] = out[i];
    return m;
}

/* order the cut points of one plane counter-clockwise about their centroid (in-plane basis u, v), drop exact duplicates */
static int make_cap(V3 *cp, int n, V3 nrm, V3 *cap) {
    if (n < 3) return 0;
    V3 c = {0, 0, 0};
    for (int i = 0; i < n; ++i) { c.x += cp[i].x; c.y += cp[i].y; c.z += cp[i].z; }
    c.x /= n; c.y /= n; c.z /= n;
    /* in-plane basis: u = normalised (nrm x e) with e the axis least aligned with nrm, v = nrm x u */
    const double ax = fabs(nrm.x), ay = fabs(nrm.y), az = fabs(nrm.z);
    V3 e = {0, 0, 0};
    if (ax <= ay && ax <= az) e.x = 1; else if (ay <= az) e.y = 1; else e.z = 1;
    V3 u = {nrm.y * e.z - nrm.z * e.y, nrm.z * e.x - nrm.x * e.z, nrm.x * e.y - nrm.y * e.x};
    const double ul = sqrt(u.x * u.x + u.y * u.y + u.z * u.z);
    u.x /= ul; u.y /= ul; u.z /= ul;
    const V3 v = {nrm.y * u.z - nrm.z * u.y, nrm.z * u.x - nrm.x * u.z, nrm.x * u.y - nrm.y * u.x};
    double ang[B9_MAXC];
    for (int i = 0; i < n; ++i) {
        const double px = cp[i].x - c.x, py = cp[i].y - c.y, pz = cp[i].z - c.z;
        ang[i] = atan2(px * v.x + py * v.y + pz * v.z, px * u.x + py * u.y + pz * u.z);
    }
    for (int i = 1; i < n; ++i) {           /* insertion sort */
        const double a = ang[i]; const V3 p = cp[i];
        int k = i - 1;
        while (k >= 0 && ang[k] > a) { ang[k + 1] = ang[k]; cp[k + 1] = cp[k]; --k; }
        ang[k + 1] = a; cp[k + 1] = p;
    }
    int m = 0;
    for (int i = 0; i < n; ++i) {
        if (m > 0 && cap[m - 1].x == cp[i].x && cap[m - 1].y == cp[i].y && cap[m - 1].z == cp[i].z) continue;
        if (m < B9_MAXV) cap[m++] = cp[i];
    }
    if (m > 1 && cap[m - 1].x == cap[0].x && cap[m - 1].y == cap[0].y && cap[m - 1].z == cap[0].z) --m;
    return m >= 3 ? m : 0;
}

ORC_API double orc_box9_intersection_volume(const float *a, const float *b) {
    double ra[9], rb[9], m[9];
    box9_rot(a, ra); box9_rot(b, rb);
    /* m = ra^T rb: columns = B's axes in A's frame; t = ra^T (cb - ca) */
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) m[i * 3 + j] = ra[0 * 3 + i] * rb[0 * 3 + j] + ra[1 * 3 + i] * rb[1 * 3 + j] + ra[2 * 3 + i] * rb[2 * 3 + j];
    const double dx = (double)b[0] - a[0], dy = (double)b[1] - a[1], dz = (double)b[2] - a[2];
    const double t[3] = {ra[0] * dx + ra[3] * dy + ra[6] * dz, ra[1] * dx + ra[4] * dy + ra[7] * dz, ra[2] * dx + ra[5] * dy + ra[8] * dz};
    const double ha[3] = {a[3] / 2.0, a[4] / 2.0, a[5] / 2.0}, hb[3] = {b[3] / 2.0, b[4] / 2.0, b[5] / 2.0};
    if (!(ha[0] > 0 && ha[1] > 0 && ha[2] > 0 && hb[0] > 0 && hb[1] > 0 && hb[2] > 0)) return 0.0;
    V3 face[12][B9_MAXV];
    int nv[12];
    double off[12];
    /* the six faces of A, outward normal +-e_i, offset ha_i */
    for (int i = 0; i < 3; ++i)
        for (int sgn = 0; sgn < 2; ++sgn) {
            const int f = 2 * i + sgn, u = (i + 1) % 3, v = (i + 2) % 3;
            const double s = sgn ? 1.0 : -1.0;
            const double su[4] = {-1, 1, 1, -1}, sv[4] = {-1, -1, 1, 1};
            for (int k = 0; k < 4; ++k) {
                double c[3];
                c[i] = s * ha[i]; c[u] = su[k] * ha[u]; c[v] = sv[k] * ha[v];
                face[f][k].x = c[0]; face[f][k].y = c[1]; face[f][k].z = c[2];
            }
            nv[f] = 4; off[f] = ha[i];
        }
    int nf = 6;
    for (int j = 0; j < 3; ++j)
        for (int sgn = 0; sgn < 2; ++sgn) {
            const double s = sgn ? 1.0 : -1.0;
            const V3 nr = {s * m[0 * 3 + j], s * m[1 * 3 + j], s * m[2 * 3 + j]};
            const double d = hb[j] + (nr.x * t[0] + nr.y * t[1] + nr.z * t[2]);
            V3 cp[B9_MAXC];
            int ncp = 0;
            for (int f = 0; f < nf; ++f)
                if (nv[f] > 0) nv[f] = clip_face(face[f], nv[f], nr, d, cp, &ncp);
            nv[nf] = make_cap(cp, ncp, nr, face[nf]);
            off[nf] = d;
            ++nf;
        }
    double vol3 = 0.0;
    for (int f = 0; f < nf; ++f)
        if (nv[f] >= 3) vol3 += off[f] * poly_area(face[f], nv[f]);
    const double vol = vol3 / 3.0;
    return vol > 0.0 ? vol : 0.0;
}

/* (na, 9) x (nb, 9) -> (na, nb) float: IoU = inter / max(va + vb - inter, 1e-6)  (the clamp of iou3d_nms_utils.py:79) */
ORC_API void orc_boxes_iou3d_9dof(int na, const float *a, int nb, const float *b, float *out) {
    for (int i = 0; i < na; ++i)
        for (int j = 0; j < nb; ++j) {
            const float *A = a + (size_t)i * 9, *B = b + (size_t)j * 9;
            const double inter = orc_box9_intersection_volume(A, B);
            const double va = (double)A[3] * A[4] * A[5], vb = (double)B[3] * B[4] * B[5];
            double den = va + vb - inter;
            if (den < 1e-6) den = 1e-6;
            out[(size_t)i * nb + j] = (float)(inter / den);
        }
}

/* greedy NMS over boxes sorted by descending score with the full-pose IoU (same sweep as orc_nms); returns num kept */
ORC_API int orc_nms_9dof(int n, const float *boxes, float thresh, int64_t *keep) {
    char *dead = (char *)calloc(n > 0 ? n : 1, 1);
    int nk = 0;
    for (int i = 0; i < n; ++i) {
        if (dead[i]) continue;
        keep[nk++] = i;
        for (int j = i + 1; j < n; ++j) {
            if (dead[j]) continue;
            float v;
            orc_boxes_iou3d_9dof(1, boxes + (size_t)i * 9, 1, boxes + (size_t)j * 9, &v);
            if (v > thresh) dead[j] = 1;
        }
    }
    free(dead);
    return nk;
}
