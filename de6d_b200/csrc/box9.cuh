// Full-pose (9-DoF) box geometry: intersection volume of two oriented boxes [x, y, z, dx, dy, dz, rz, ry, rx] with
// R = Rx(rx) Ry(ry) Rz(rz) -- the convention of pcdet/utils/box_utils.py:59-72 (scipy Rotation.from_euler('zyx', (rz, ry, rx)))
// that Det6D's head predicts (point_head_box6d_vote.py: PointBinResidual6DCoder).  SURVEY.md 8(f) rank 3: the reference
// evaluates IoU / NMS on boxes[:, 0:7] only, ignoring pitch and roll (point_head_box6d_vote.py:355, model_nms_utils.py:18).
//
// Algorithm (oracle/de6d_oracle.c: orc_box9_intersection_volume states it in double and is pinned against scipy's
// HalfspaceIntersection + ConvexHull): in A's frame the polyhedron A, kept as convex face polygons, is clipped by B's six
// half-spaces; the cut points of each plane (computed from the inside vertex towards the outside one, so the two faces
// sharing an edge produce identical bits) are ordered about their centroid and become the cap face; V = 1/3 * sum(offset *
// area).  Coplanar faces (identical / padded boxes, shared ground plane) cost O(rounding) instead of a face counted twice.
#pragma once
#include "common.cuh"
#include <math.h>

namespace de6d {

struct Box9Geo {
    float c[3];     // centre
    float h[3];     // half extents
    float r[9];     // rotation, row-major: world = r * local + c
    float rad;      // bounding-sphere radius
    float vol;
};

__device__ __forceinline__ Box9Geo make_geo9(const float *b) {
    Box9Geo g;
    g.c[0] = b[0]; g.c[1] = b[1]; g.c[2] = b[2];
    g.h[0] = b[3] * 0.5f; g.h[1] = b[4] * 0.5f; g.h[2] = b[5] * 0.5f;
    float sz, cz, sy, cy, sx, cx;
    sincosf(b[6], &sz, &cz); sincosf(b[7], &sy, &cy); sincosf(b[8], &sx, &cx);
    g.r[0] = cy * cz;                 g.r[1] = -cy * sz;                g.r[2] = sy;
    g.r[3] = sx * sy * cz + cx * sz;  g.r[4] = -sx * sy * sz + cx * cz; g.r[5] = -sx * cy;
    g.r[6] = -cx * sy * cz + sx * sz; g.r[7] = cx * sy * sz + sx * cz;  g.r[8] = cx * cy;
    g.rad = sqrtf(g.h[0] * g.h[0] + g.h[1] * g.h[1] + g.h[2] * g.h[2]) * 1.0001f;
    g.vol = b[3] * b[4] * b[5];
    return g;
}

__device__ __forceinline__ bool cannot_touch9(const Box9Geo &a, const Box9Geo &b) {
    const float dx = a.c[0] - b.c[0], dy = a.c[1] - b.c[1], dz = a.c[2] - b.c[2];
    const float rr = a.rad + b.rad;
    return dx * dx + dy * dy + dz * dz > rr * rr * 1.0001f;   // also true for NaN-free far pairs only; NaN falls through to the clip
}

constexpr int B9_MAXV = 12;   // vertices per face: a quad cut by <= 6 planes has <= 10, a cap of a <= 12-face polytope <= 11
constexpr int B9_MAXC = 24;   // cut points per plane before duplicates are dropped (each crossing edge is seen by two faces)

struct P3 { float x, y, z; };

__device__ __forceinline__ float poly_area9(const P3 *p, int n) {
    if (n < 3) return 0.f;
    float ax = 0.f, ay = 0.f, az = 0.f;
    for (int i = 1; i + 1 < n; ++i) {
        const float ux = p[i].x - p[0].x, uy = p[i].y - p[0].y, uz = p[i].z - p[0].z;
        const float vx = p[i + 1].x - p[0].x, vy = p[i + 1].y - p[0].y, vz = p[i + 1].z - p[0].z;
        ax += uy * vz - uz * vy; ay += uz * vx - ux * vz; az += ux * vy - uy * vx;
    }
    return 0.5f * sqrtf(ax * ax + ay * ay + az * az);
}

// monotone in the polar angle of (x, y) over (-pi, pi]: cheaper than atan2f and all that ordering needs
__device__ __forceinline__ float pseudo_angle(float y, float x) {
    const float d = fabsf(x) + fabsf(y);
    const float p = d > 0.f ? x / d : 1.f;       // 1 .. -1 as the angle goes 0 .. pi
    return y < 0.f ? p - 1.f : 1.f - p;          // (-2, 0) below the axis, [0, 2] above
}

// Intersection volume of two boxes.  ~2 KB of thread-local polygon storage: call it from few, densely packed threads.
__device__ float box9_intersection_volume(const Box9Geo &A, const Box9Geo &B) {
    if (!(A.h[0] > 0.f && A.h[1] > 0.f && A.h[2] > 0.f && B.h[0] > 0.f && B.h[1] > 0.f && B.h[2] > 0.f)) return 0.f;
    // m = Ra^T Rb (columns: B's axes in A's frame), t = Ra^T (cb - ca)
    float m[9], t[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) m[i * 3 + j] = A.r[i] * B.r[j] + A.r[3 + i] * B.r[3 + j] + A.r[6 + i] * B.r[6 + j];
        t[i] = A.r[i] * (B.c[0] - A.c[0]) + A.r[3 + i] * (B.c[1] - A.c[1]) + A.r[6 + i] * (B.c[2] - A.c[2]);
    }
    P3 face[12][B9_MAXV];
    int nv[12];
    float off[12];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int sgn = 0; sgn < 2; ++sgn) {
            const int f = 2 * i + sgn, u = (i + 1) % 3, v = (i + 2) % 3;
            const float s = sgn ? 1.f : -1.f;
            const float su[4] = {-1.f, 1.f, 1.f, -1.f}, sv[4] = {-1.f, -1.f, 1.f, 1.f};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float c[3];
                c[i] = s * A.h[i]; c[u] = su[k] * A.h[u]; c[v] = sv[k] * A.h[v];
                face[f][k].x = c[0]; face[f][k].y = c[1]; face[f][k].z = c[2];
            }
            nv[f] = 4; off[f] = A.h[i];
        }
    }
    int nf = 6;
    for (int pl = 0; pl < 6; ++pl) {
        const int j = pl >> 1;
        const float s = (pl & 1) ? 1.f : -1.f;
        const float nx = s * m[j], ny = s * m[3 + j], nz = s * m[6 + j];
        const float d = B.h[j] + (nx * t[0] + ny * t[1] + nz * t[2]);
        P3 cp[B9_MAXC];
        int ncp = 0;
        for (int f = 0; f < nf; ++f) {
            const int n = nv[f];
            if (n <= 0) continue;
            P3 out[B9_MAXV];
            float sd[B9_MAXV];
            int mo = 0;
            for (int i = 0; i < n; ++i) sd[i] = nx * face[f][i].x + ny * face[f][i].y + nz * face[f][i].z - d;
            for (int i = 0; i < n; ++i) {
                const int k = (i + 1 == n) ? 0 : i + 1;
                const bool pin = sd[i] <= 0.f, qin = sd[k] <= 0.f;
                if (pin && mo < B9_MAXV) out[mo++] = face[f][i];
                if (pin != qin) {
                    const P3 a = pin ? face[f][i] : face[f][k], b = pin ? face[f][k] : face[f][i];
                    const float sa = pin ? sd[i] : sd[k], sb = pin ? sd[k] : sd[i];
                    const float tt = sa / (sa - sb);
                    P3 x;
                    x.x = __fmaf_rn(tt, b.x - a.x, a.x); x.y = __fmaf_rn(tt, b.y - a.y, a.y); x.z = __fmaf_rn(tt, b.z - a.z, a.z);
                    if (mo < B9_MAXV) out[mo++] = x;
                    if (ncp < B9_MAXC) cp[ncp++] = x;
                }
            }
            for (int i = 0; i < mo; ++i) face[f][i] = out[i];
            nv[f] = mo;
        }
        // cap: cut points ordered about their centroid in the plane's basis (u, v), exact duplicates dropped
        int nc = 0;
        if (ncp >= 3) {
            float cx = 0.f, cy = 0.f, cz = 0.f;
            for (int i = 0; i < ncp; ++i) { cx += cp[i].x; cy += cp[i].y; cz += cp[i].z; }
            const float inv = 1.f / (float)ncp;
            cx *= inv; cy *= inv; cz *= inv;
            const float ax = fabsf(nx), ay = fabsf(ny), az = fabsf(nz);
            float ex = 0.f, ey = 0.f, ez = 0.f;
            if (ax <= ay && ax <= az) ex = 1.f; else if (ay <= az) ey = 1.f; else ez = 1.f;
            float ux = ny * ez - nz * ey, uy = nz * ex - nx * ez, uz = nx * ey - ny * ex;
            const float ul = rsqrtf(ux * ux + uy * uy + uz * uz);
            ux *= ul; uy *= ul; uz *= ul;
            const float vx = ny * uz - nz * uy, vy = nz * ux - nx * uz, vz = nx * uy - ny * ux;
            float ang[B9_MAXC];
            for (int i = 0; i < ncp; ++i) {
                const float px = cp[i].x - cx, py = cp[i].y - cy, pz = cp[i].z - cz;
                ang[i] = pseudo_angle(px * vx + py * vy + pz * vz, px * ux + py * uy + pz * uz);
            }
            for (int i = 1; i < ncp; ++i) {
                const float a = ang[i];
                const P3 p = cp[i];
                int k = i - 1;
                while (k >= 0 && ang[k] > a) { ang[k + 1] = ang[k]; cp[k + 1] = cp[k]; --k; }
                ang[k + 1] = a; cp[k + 1] = p;
            }
            for (int i = 0; i < ncp; ++i) {
                if (nc > 0 && face[nf][nc - 1].x == cp[i].x && face[nf][nc - 1].y == cp[i].y && face[nf][nc - 1].z == cp[i].z) continue;
                if (nc < B9_MAXV) face[nf][nc++] = cp[i];
            }
            if (nc > 1 && face[nf][nc - 1].x == face[nf][0].x && face[nf][nc - 1].y == face[nf][0].y && face[nf][nc - 1].z == face[nf][0].z) --nc;
            if (nc < 3) nc = 0;
        }
        nv[nf] = nc;
        off[nf] = d;
        ++nf;
    }
    float vol3 = 0.f;
    for (int f = 0; f < 12; ++f)
        if (nv[f] >= 3) vol3 += off[f] * poly_area9(face[f], nv[f]);
    const float vol = vol3 * (1.f / 3.f);
    return vol > 0.f ? vol : 0.f;
}

// IoU with the reference composition's clamp (iou3d_nms_utils.py:79: clamp(min=1e-6) on the union volume)
__device__ __forceinline__ float iou9(const Box9Geo &A, const Box9Geo &B) {
    const float inter = box9_intersection_volume(A, B);
    return inter / fmaxf(A.vol + B.vol - inter, 1e-6f);
}

}  // namespace de6d
