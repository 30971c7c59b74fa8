"""ctypes/numpy front end of the CPU oracle (oracle/de6d_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never by de6d_b200/.
Parity status: pinned (see the header of de6d_oracle.c and tests/golden/).

Each function mirrors one native entry point of the reference, with the
reference's own argument order (pointnet2_api.cpp:11-30, iou3d_nms_api.cpp:11-17,
roiaware_pool3d.cpp:172-177) but numpy arrays in / numpy arrays out, and it
applies the same pre-initialisation the reference Python wrappers apply
(temp=1e10, idx=0, box index=-1 ...; pointnet2_utils.py:25-26,294,322-323).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libde6d_oracle.so")
_SRC = os.path.join(_HERE, "de6d_oracle.c")
_lib = None

_f = C.POINTER(C.c_float)
_i = C.POINTER(C.c_int)


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
                               "-fvisibility=hidden", "-o", _SO, _SRC, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_box_overlap.restype = C.c_float
        _lib.orc_iou_bev.restype = C.c_float
        _lib.orc_nms.restype = C.c_int
        _lib.orc_opt_n_threads.restype = C.c_int
    return _lib


def _fp(a):
    return a.ctypes.data_as(_f)


def _ip(a):
    return a.ctypes.data_as(_i)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def opt_n_threads(n):
    return lib().orc_opt_n_threads(int(n))


def furthest_point_sample(xyz, npoint, return_temp=False):
    xyz = _f32(xyz); B, N, _ = xyz.shape
    temp = np.full((B, N), 1e10, np.float32); out = np.zeros((B, npoint), np.int32)
    lib().orc_fps(B, N, npoint, _fp(xyz), _fp(temp), _ip(out))
    return (out, temp) if return_temp else out


def furthest_point_sample_matrix(matrix, npoint, return_temp=False):
    matrix = _f32(matrix); B, N, _ = matrix.shape
    temp = np.full((B, N), 1e10, np.float32); out = np.zeros((B, npoint), np.int32)
    lib().orc_fps_matrix(B, N, npoint, _fp(matrix), _fp(temp), _ip(out))
    return (out, temp) if return_temp else out


def furthest_point_sample_weights(xyz, weights, npoint, return_temp=False):
    xyz = _f32(xyz); weights = _f32(weights); B, N, _ = xyz.shape
    temp = np.full((B, N), 1e10, np.float32); out = np.zeros((B, npoint), np.int32)
    lib().orc_fps_weights(B, N, npoint, _fp(xyz), _fp(weights), _fp(temp), _ip(out))
    return (out, temp) if return_temp else out


def calc_dist_matrix_for_sampling(xyz, features=None, gamma=1.0):
    """pointnet2_utils.py:36-44 with direct differences in a fixed order (see orc_dist_matrix).
    xyz (B, N, 3), features (B, N, C) or None -> (B, N, N)."""
    xyz = _f32(xyz); B, N, _ = xyz.shape
    out = np.zeros((B, N, N), np.float32)
    if features is not None:
        features = _f32(features); Cc = features.shape[2]
        lib().orc_dist_matrix(B, N, Cc, _fp(xyz), _fp(features), C.c_float(gamma), _fp(out))
    else:
        lib().orc_dist_matrix(B, N, 0, _fp(xyz), None, C.c_float(gamma), _fp(out))
    return out


def gather_operation(features, idx):
    features = _f32(features); idx = _i32(idx)
    B, Cc, N = features.shape; M = idx.shape[1]
    out = np.zeros((B, Cc, M), np.float32)
    lib().orc_gather_points(B, Cc, N, M, _fp(features), _ip(idx), _fp(out))
    return out


def gather_operation_grad(grad_out, idx, N):
    grad_out = _f32(grad_out); idx = _i32(idx)
    B, Cc, M = grad_out.shape
    g = np.zeros((B, Cc, N), np.float32)
    lib().orc_gather_points_grad(B, Cc, N, M, _fp(grad_out), _ip(idx), _fp(g))
    return g


def ball_query(radius, nsample, xyz, new_xyz):
    xyz = _f32(xyz); new_xyz = _f32(new_xyz)
    B, N, _ = xyz.shape; M = new_xyz.shape[1]
    idx = np.zeros((B, M, nsample), np.int32)
    lib().orc_ball_query(B, N, M, C.c_float(radius), nsample, _fp(new_xyz), _fp(xyz), _ip(idx))
    return idx


def ball_query_cnt(radius, nsample, xyz, new_xyz):
    xyz = _f32(xyz); new_xyz = _f32(new_xyz)
    B, N, _ = xyz.shape; M = new_xyz.shape[1]
    idx = np.zeros((B, M, nsample), np.int32); cnt = np.zeros((B, M), np.int32)
    lib().orc_ball_query_cnt(B, N, M, C.c_float(radius), nsample, _fp(new_xyz), _fp(xyz), _ip(cnt), _ip(idx))
    return cnt, idx


def ball_query_dilated(radius_in, radius_out, nsample, xyz, new_xyz):
    xyz = _f32(xyz); new_xyz = _f32(new_xyz)
    B, N, _ = xyz.shape; M = new_xyz.shape[1]
    idx = np.zeros((B, M, nsample), np.int32); cnt = np.zeros((B, M), np.int32)
    lib().orc_ball_query_dilated(B, N, M, C.c_float(radius_in), C.c_float(radius_out), nsample,
                                 _fp(new_xyz), _fp(xyz), _ip(cnt), _ip(idx))
    return cnt, idx


def grouping_operation(features, idx):
    features = _f32(features); idx = _i32(idx)
    B, Cc, N = features.shape; _, M, ns = idx.shape
    out = np.zeros((B, Cc, M, ns), np.float32)
    lib().orc_group_points(B, Cc, N, M, ns, _fp(features), _ip(idx), _fp(out))
    return out


def grouping_operation_grad(grad_out, idx, N):
    grad_out = _f32(grad_out); idx = _i32(idx)
    B, Cc, M, ns = grad_out.shape
    g = np.zeros((B, Cc, N), np.float32)
    lib().orc_group_points_grad(B, Cc, N, M, ns, _fp(grad_out), _ip(idx), _fp(g))
    return g


def three_nn(unknown, known):
    """Returns (sqrt(dist2), idx) like pointnet2_utils.py:152-181."""
    unknown = _f32(unknown); known = _f32(known)
    B, n, _ = unknown.shape; m = known.shape[1]
    d2 = np.zeros((B, n, 3), np.float32); idx = np.zeros((B, n, 3), np.int32)
    lib().orc_three_nn(B, n, m, _fp(unknown), _fp(known), _fp(d2), _ip(idx))
    return np.sqrt(d2), idx


def three_interpolate(features, idx, weight):
    features = _f32(features); idx = _i32(idx); weight = _f32(weight)
    B, c, m = features.shape; n = idx.shape[1]
    out = np.zeros((B, c, n), np.float32)
    lib().orc_three_interpolate(B, c, m, n, _fp(features), _ip(idx), _fp(weight), _fp(out))
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    grad_out = _f32(grad_out); idx = _i32(idx); weight = _f32(weight)
    B, c, n = grad_out.shape
    g = np.zeros((B, c, m), np.float32)
    lib().orc_three_interpolate_grad(B, c, n, m, _fp(grad_out), _ip(idx), _fp(weight), _fp(g))
    return g


def boxes_overlap_bev(a, b):
    a = _f32(a); b = _f32(b)
    out = np.zeros((a.shape[0], b.shape[0]), np.float32)
    lib().orc_boxes_overlap_bev(a.shape[0], _fp(a), b.shape[0], _fp(b), _fp(out))
    return out


def boxes_iou_bev(a, b):
    a = _f32(a); b = _f32(b)
    out = np.zeros((a.shape[0], b.shape[0]), np.float32)
    lib().orc_boxes_iou_bev(a.shape[0], _fp(a), b.shape[0], _fp(b), _fp(out))
    return out


boxes_bev_iou_cpu = boxes_iou_bev


def boxes_iou3d(a, b):
    a = _f32(a); b = _f32(b)
    out = np.zeros((a.shape[0], b.shape[0]), np.float32)
    lib().orc_boxes_iou3d(a.shape[0], _fp(a), b.shape[0], _fp(b), _fp(out))
    return out


def _nms_sorted(boxes_sorted, thresh, normal, return_mask=False):
    n = boxes_sorted.shape[0]
    keep = np.zeros(max(n, 1), np.int64)
    cb = (n + 63) // 64
    mask = np.zeros((n, cb), np.uint64) if return_mask else None
    nk = lib().orc_nms(n, _fp(boxes_sorted), C.c_float(thresh), int(normal),
                       keep.ctypes.data_as(C.POINTER(C.c_int64)),
                       mask.ctypes.data_as(C.POINTER(C.c_uint64)) if return_mask else None)
    return (keep[:nk], mask) if return_mask else keep[:nk]


def nms_sorted(boxes_sorted, thresh, normal=False, return_mask=False):
    """Native-level nms (iou3d_nms.cpp:90-136): boxes already score-sorted; returns kept positions."""
    return _nms_sorted(_f32(boxes_sorted), thresh, normal, return_mask)


def nms_gpu(boxes, scores, thresh, pre_maxsize=None, normal=False):
    """iou3d_nms_utils.py:84-99 (stable descending sort, as torch.sort is asked to be in the tests)."""
    boxes = _f32(boxes)
    order = np.argsort(-np.asarray(scores, np.float32), kind="stable")
    if pre_maxsize is not None:
        order = order[:pre_maxsize]
    keep = _nms_sorted(np.ascontiguousarray(boxes[order]), thresh, normal)
    return order[keep]


def points_in_boxes_gpu(points, boxes):
    points = _f32(points); boxes = _f32(boxes)
    B, M, _ = points.shape; T = boxes.shape[1]
    out = np.full((B, M), -1, np.int32)
    lib().orc_points_in_boxes_gpu(B, T, M, _fp(boxes), _fp(points), _ip(out))
    return out


def points_in_boxes_cpu(points, boxes):
    points = _f32(points); boxes = _f32(boxes)
    out = np.zeros((boxes.shape[0], points.shape[0]), np.int32)
    lib().orc_points_in_boxes_cpu(boxes.shape[0], points.shape[0], _fp(boxes), _fp(points), _ip(out))
    return out


def boxes_iou3d_9dof(a, b):
    """Full-pose IoU: (N, 9) x (M, 9) [x, y, z, dx, dy, dz, rz, ry, rx] -> (N, M) float32 (double arithmetic inside)."""
    a = _f32(a); b = _f32(b)
    assert a.shape[1] == 9 and b.shape[1] == 9
    out = np.zeros((a.shape[0], b.shape[0]), np.float32)
    lib().orc_boxes_iou3d_9dof(a.shape[0], _fp(a), b.shape[0], _fp(b), _fp(out))
    return out


def box9_intersection_volume(a, b):
    a = _f32(a); b = _f32(b)
    fn = lib().orc_box9_intersection_volume
    fn.restype = C.c_double
    return float(fn(_fp(a), _fp(b)))


def nms_9dof(boxes, scores, thresh, pre_maxsize=None):
    """Greedy NMS with the full-pose IoU; same calling convention as nms_gpu (returns indices into `boxes`)."""
    boxes = _f32(boxes)
    order = np.argsort(-np.asarray(scores, np.float32), kind="stable")
    if pre_maxsize is not None:
        order = order[:pre_maxsize]
    bs = np.ascontiguousarray(boxes[order])
    keep = np.zeros(max(len(order), 1), np.int64)
    nk = lib().orc_nms_9dof(len(order), _fp(bs), C.c_float(thresh), keep.ctypes.data_as(C.POINTER(C.c_int64)))
    return order[keep[:nk]]


def points_in_boxes3d(points, boxes3d):
    """box_utils.points_in_boxes3d (pcdet/utils/box_utils.py:110-124): points (n, 3+), boxes (m, 9) -> (n,) int64."""
    pts = _f32(np.asarray(points)[:, :3]); boxes = _f32(boxes3d)
    assert boxes.shape[-1] == 9
    out = np.full(pts.shape[0], -1, np.int64)
    lib().orc_points_in_boxes9(boxes.shape[0], pts.shape[0], _fp(boxes), _fp(pts), out.ctypes.data_as(C.POINTER(C.c_int64)))
    return out


def break_up_pc(points, batch_size):
    """pointnet2_backbone.py:193-222 restated with numpy: points (B*N, 4+C) rows [batch_idx, x, y, z, features...] ->
    (batch_idx (B,N) f32, xyz (B,N,3), features (B,C,N) or None); asserts that every frame holds the same number of rows
    like the reference (:214-218)."""
    pc = _f32(points)
    bidx = pc[:, 0]
    cnt = np.array([(bidx == b).sum() for b in range(batch_size)])
    assert cnt.min() == cnt.max()
    xyz = np.ascontiguousarray(pc[:, 1:4]).reshape(batch_size, -1, 3)
    feats = None
    if pc.shape[1] > 4:
        feats = np.ascontiguousarray(pc[:, 4:].reshape(batch_size, -1, pc.shape[1] - 4).transpose(0, 2, 1))
    return bidx.reshape(batch_size, -1).astype(np.float32), xyz, feats
