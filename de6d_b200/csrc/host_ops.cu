// Host-thread implementations of the two entry points the reference itself evaluates on the CPU:
//   boxes_iou_bev_cpu      iou3d_nms/src/iou3d_cpu.cpp:232-252     (a15; called from database_sampler.py:232-233)
//   points_in_boxes_cpu    roiaware_pool3d/src/roiaware_pool3d.cpp:143-168 (a17; kitti_dataset.py:248, box_utils.py:104)
// Both are called by the reference from forked DataLoader workers, where a CUDA context cannot be created, so these
// run on the calling thread (optionally split over `nthreads` std::threads by rows) and never touch the device.
// This is the reference's own contract for these two functions, not a fallback for a device op: every device entry
// point of this library still fails without a GPU.  The device twins (de6d_boxes_iou_bev, de6d_points_in_boxes_mask)
// stay available for callers that already hold device tensors.
//
// Arithmetic: the same float expression tree as the reference's host code (libm cosf/sinf/atan2f, no contraction:
// this TU's host pass targets baseline x86-64 / aarch64 without -ffast-math), organised differently: per-box trig and
// corners are evaluated once per box instead of once per pair, polar angles once per vertex instead of once per
// comparison -- the values compared are identical, so results are bit-identical to the reference functions
// (tests/test_host_ops.py checks that against the reference build, live, on the CPU).
#include <cmath>
#include <cstddef>
#include <thread>
#include <vector>

#include "common.cuh"

namespace de6d { namespace host {

struct V2 { float x, y; };

struct BoxGeom {
    V2 c[5];            // corners in the reference's order (x1y1, x2y1, x2y2, x1y2) rotated about the centre; c[4] = c[0]
    float cx, cy;       // centre
    float hx, hy;       // dx / 2 + MARGIN, dy / 2 + MARGIN  (check_in_box2d :73-83)
    float nc, ns;       // cos(-heading), sin(-heading)
    float area;
};

static inline float cr3(const V2& p1, const V2& p2, const V2& p0) {
    return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y);
}

static inline float fmn(float a, float b) { return a > b ? b : a; }   // the reference's own min / max (:30-36)
static inline float fmx(float a, float b) { return a > b ? a : b; }

static void geom(const float* b, BoxGeom& g) {
    const float MARGIN = 1e-2f;
    const float x1 = b[0] - b[3] / 2, y1 = b[1] - b[4] / 2, x2 = b[0] + b[3] / 2, y2 = b[1] + b[4] / 2;
    const float co = cosf(b[6]), si = sinf(b[6]);
    const float px[4] = {x1, x2, x2, x1}, py[4] = {y1, y1, y2, y2};
    for (int k = 0; k < 4; ++k) {
        const float ox = px[k] - b[0], oy = py[k] - b[1];
        g.c[k].x = ox * co + oy * (-si) + b[0];
        g.c[k].y = ox * si + oy * co + b[1];
    }
    g.c[4] = g.c[0];
    g.cx = b[0]; g.cy = b[1];
    g.hx = b[3] / 2 + MARGIN; g.hy = b[4] / 2 + MARGIN;
    g.nc = cosf(-b[6]); g.ns = sinf(-b[6]);
    g.area = b[3] * b[4];
}

static inline bool inside(const BoxGeom& g, const V2& p) {
    const float rx = (p.x - g.cx) * g.nc + (p.y - g.cy) * (-g.ns);
    const float ry = (p.x - g.cx) * g.ns + (p.y - g.cy) * g.nc;
    return fabsf(rx) < g.hx && fabsf(ry) < g.hy;
}

static inline bool cross_point(const V2& p1, const V2& p0, const V2& q1, const V2& q0, V2& ans) {
    const float EPS = 1e-8f;
    if (!(fmn(p0.x, p1.x) <= fmx(q0.x, q1.x) && fmn(q0.x, q1.x) <= fmx(p0.x, p1.x) &&
          fmn(p0.y, p1.y) <= fmx(q0.y, q1.y) && fmn(q0.y, q1.y) <= fmx(p0.y, p1.y))) return false;
    const float s1 = cr3(q0, p1, p0), s2 = cr3(p1, q1, p0), s3 = cr3(p0, q1, q0), s4 = cr3(q1, p1, q0);
    if (!(s1 * s2 > 0 && s3 * s4 > 0)) return false;
    const float s5 = cr3(q1, p1, p0);
    if (fabsf(s5 - s1) > EPS) {
        ans.x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
        ans.y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
    } else {
        const float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
        const float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
        const float D = a0 * b1 - a1 * b0;
        ans.x = (b0 * c1 - b1 * c0) / D;
        ans.y = (a1 * c0 - a0 * c1) / D;
    }
    return true;
}

// box_overlap (iou3d_cpu.cpp:128-222): vertices of the intersection polygon = edge crossings + contained corners,
// ordered by polar angle about their mean with the reference's bubble sort (a stable pass structure: which of two
// vertices with equal angles comes first is part of the result), fan area.
static float overlap(const BoxGeom& A, const BoxGeom& B) {
    V2 v[16];
    float ang[16];
    V2 ctr = {0.f, 0.f};
    int n = 0;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            if (cross_point(A.c[i + 1], A.c[i], B.c[j + 1], B.c[j], v[n])) {
                ctr.x += v[n].x; ctr.y += v[n].y; ++n;
            }
    for (int k = 0; k < 4; ++k) {
        if (inside(A, B.c[k])) { ctr.x += B.c[k].x; ctr.y += B.c[k].y; v[n++] = B.c[k]; }
        if (inside(B, A.c[k])) { ctr.x += A.c[k].x; ctr.y += A.c[k].y; v[n++] = A.c[k]; }
    }
    ctr.x /= n; ctr.y /= n;
    for (int k = 0; k < n; ++k) ang[k] = atan2f(v[k].y - ctr.y, v[k].x - ctr.x);
    for (int j = 0; j < n - 1; ++j)
        for (int i = 0; i < n - j - 1; ++i)
            if (ang[i] > ang[i + 1]) {
                const V2 t = v[i]; v[i] = v[i + 1]; v[i + 1] = t;
                const float a = ang[i]; ang[i] = ang[i + 1]; ang[i + 1] = a;
            }
    float area = 0.f;
    for (int k = 0; k < n - 1; ++k) {
        const V2 u = {v[k].x - v[0].x, v[k].y - v[0].y}, w = {v[k + 1].x - v[0].x, v[k + 1].y - v[0].y};
        area += u.x * w.y - u.y * w.x;
    }
    return fabsf(area) / 2.0f;
}

template <class F>
static void rows_parallel(int rows, int nthreads, F&& body) {
    if (nthreads <= 1 || rows < 2 * nthreads) { body(0, rows); return; }
    std::vector<std::thread> th;
    const int per = (rows + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; ++t) {
        const int lo = t * per, hi = lo + per < rows ? lo + per : rows;
        if (lo < hi) th.emplace_back([=, &body] { body(lo, hi); });
    }
    for (auto& x : th) x.join();
}

}}  // namespace de6d::host

using namespace de6d::host;

extern "C" int de6d_boxes_iou_bev_host(int n, const float* boxes_a, int m, const float* boxes_b, float* ans_iou, int nthreads) {
    if (n < 0 || m < 0) return de6d_set_error(DE6D_ERR_INVALID, "de6d_boxes_iou_bev_host: negative size");
    if (n == 0 || m == 0) return 0;
    if (!boxes_a || !boxes_b || !ans_iou) return de6d_set_error(DE6D_ERR_INVALID, "de6d_boxes_iou_bev_host: null pointer");
    std::vector<BoxGeom> ga(n), gb(m);
    for (int i = 0; i < n; ++i) geom(boxes_a + (size_t)i * 7, ga[i]);
    for (int j = 0; j < m; ++j) geom(boxes_b + (size_t)j * 7, gb[j]);
    rows_parallel(n, nthreads, [&](int lo, int hi) {
        for (int i = lo; i < hi; ++i)
            for (int j = 0; j < m; ++j) {
                const float s = overlap(ga[i], gb[j]);
                ans_iou[(size_t)i * m + j] = s / fmaxf(ga[i].area + gb[j].area - s, 1e-8f);   // iou_bev :224-229
            }
    });
    return 0;
}

extern "C" int de6d_points_in_boxes_mask_host(int t, int m, const float* boxes, const float* pts, int* mask, int nthreads) {
    if (t < 0 || m < 0) return de6d_set_error(DE6D_ERR_INVALID, "de6d_points_in_boxes_mask_host: negative size");
    if (t == 0 || m == 0) return 0;
    if (!boxes || !pts || !mask) return de6d_set_error(DE6D_ERR_INVALID, "de6d_points_in_boxes_mask_host: null pointer");
    rows_parallel(t, nthreads, [&](int lo, int hi) {
        for (int k = lo; k < hi; ++k) {
            // check_pt_in_box3d (roiaware_pool3d.cpp:121-140): bounds compared in double (dz / 2.0, dx / 2.0 + MARGIN)
            const float* b = boxes + (size_t)k * 7;
            const float cx = b[0], cy = b[1], cz = b[2];
            const double hz = (double)b[5] / 2.0, hx = (double)b[3] / 2.0 + (double)1e-2f, hy = (double)b[4] / 2.0 + (double)1e-2f;
            const float ca = cosf(-b[6]), sa = sinf(-b[6]);
            int* row = mask + (size_t)k * m;
            for (int j = 0; j < m; ++j) {
                const float x = pts[(size_t)j * 3], y = pts[(size_t)j * 3 + 1], z = pts[(size_t)j * 3 + 2];
                int in = 0;
                if (!((double)fabsf(z - cz) > hz)) {
                    const float sx = x - cx, sy = y - cy;
                    const float lx = sx * ca + sy * (-sa);
                    const float ly = sx * sa + sy * ca;
                    in = ((double)fabsf(lx) < hx) & ((double)fabsf(ly) < hy);
                }
                row[j] = in;
            }
        }
    });
    return 0;
}
