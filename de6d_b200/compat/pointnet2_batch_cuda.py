"""Replacement for the reference extension `pointnet2_batch_cuda` (pointnet2_api.cpp:11-30).
Same names, same positional arguments, tensors pre-allocated by the caller; every function returns 1 like
the reference wrappers do.  `grid_query_wrapper` is not provided: no Python code in the reference calls it."""
import torch

from .._lib import load
from ._common import call, dev, need, stream_ptr

f32, i32 = torch.float32, torch.int32


def _ball_query(mode, b, n, m, radius_in, radius_out, nsample, new_xyz, xyz, idx_cnt, idx, impl=0, grid_ws=None):
    """All three variants go through de6d_ball_query_ex with scratch from torch's allocator (stream-ordered on the
    current stream, CUDA-graph capturable).  grid_ws: a workspace tensor de6d_ball_query_grid_build already filled for
    this `xyz` (BallQueryGrid in pointnet2_utils) -- the query then skips its own grid build (impl 3)."""
    need(new_xyz, b * m * 3, "new_xyz"); need(xyz, b * n * 3, "xyz"); need(idx, b * m * nsample, "idx")
    if idx_cnt is not None:
        need(idx_cnt, b * m, "idx_cnt")
    if grid_ws is not None:
        impl, ws, ws_bytes = 3, grid_ws, grid_ws.numel()
        if ws_bytes < int(load().de6d_ball_query_grid_bytes(b, n)) or ws.device != xyz.device:
            raise ValueError("ball query grid workspace does not belong to a (%d, %d, 3) cloud batch on %s" % (b, n, xyz.device))
    else:
        ws_bytes = int(load().de6d_ball_query_workspace_bytes(b, n)) if impl == 0 else 0
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=xyz.device) if ws_bytes else None
    call("de6d_ball_query_ex", mode, impl, b, n, m, radius_in, radius_out, nsample, dev(new_xyz, "new_xyz", f32),
         dev(xyz, "xyz", f32), None if idx_cnt is None else dev(idx_cnt, "idx_cnt", i32), dev(idx, "idx", i32),
         None if ws is None else ws.data_ptr(), ws_bytes, stream_ptr())
    return 1


def ball_query_grid_build(b, n, radius, xyz, workspace):
    """Extension of this package: fill `workspace` (uint8, de6d_ball_query_grid_bytes(b, n)) with the search grid of xyz."""
    need(xyz, b * n * 3, "xyz")
    call("de6d_ball_query_grid_build", b, n, float(radius), dev(xyz, "xyz", f32), dev(workspace, "workspace", torch.uint8),
         workspace.numel(), stream_ptr())
    return 1


def ball_query_wrapper(b, n, m, radius, nsample, new_xyz, xyz, idx):
    return _ball_query(0, b, n, m, 0.0, radius, nsample, new_xyz, xyz, None, idx)


def ball_query_cnt_wrapper(b, n, m, radius, nsample, new_xyz, xyz, idx_cnt, idx):
    return _ball_query(1, b, n, m, 0.0, radius, nsample, new_xyz, xyz, idx_cnt, idx)


def ball_query_dilated_wrapper(b, n, m, radius_in, radius_out, nsample, new_xyz, xyz, idx_cnt, idx):
    return _ball_query(2, b, n, m, radius_in, radius_out, nsample, new_xyz, xyz, idx_cnt, idx)


def group_points_wrapper(b, c, n, npoints, nsample, points, idx, out):
    need(points, b * c * n, "points"); need(idx, b * npoints * nsample, "idx"); need(out, b * c * npoints * nsample, "out")
    call("de6d_group_points", b, c, n, npoints, nsample, dev(points, "points", f32), dev(idx, "idx", i32),
         dev(out, "out", f32), stream_ptr())
    return 1


def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out, idx, grad_points):
    need(grad_out, b * c * npoints * nsample, "grad_out"); need(idx, b * npoints * nsample, "idx"); need(grad_points, b * c * n, "grad_points")
    call("de6d_group_points_grad", b, c, n, npoints, nsample, dev(grad_out, "grad_out", f32), dev(idx, "idx", i32),
         dev(grad_points, "grad_points", f32), stream_ptr())
    return 1


def gather_points_wrapper(b, c, n, npoints, points, idx, out):
    need(points, b * c * n, "points"); need(idx, b * npoints, "idx"); need(out, b * c * npoints, "out")
    call("de6d_gather_points", b, c, n, npoints, dev(points, "points", f32), dev(idx, "idx", i32),
         dev(out, "out", f32), stream_ptr())
    return 1


def gather_points_grad_wrapper(b, c, n, npoints, grad_out, idx, grad_points):
    need(grad_out, b * c * npoints, "grad_out"); need(idx, b * npoints, "idx"); need(grad_points, b * c * n, "grad_points")
    call("de6d_gather_points_grad", b, c, n, npoints, dev(grad_out, "grad_out", f32), dev(idx, "idx", i32),
         dev(grad_points, "grad_points", f32), stream_ptr())
    return 1


def farthest_point_sampling_wrapper(b, n, m, xyz, temp, idx):
    need(xyz, b * n * 3, "xyz"); need(temp, b * n, "temp"); need(idx, b * m, "idx")
    call("de6d_furthest_point_sampling", b, n, m, dev(xyz, "xyz", f32), dev(temp, "temp", f32), dev(idx, "idx", i32),
         stream_ptr())
    return 1


def furthest_point_sampling_matrix_wrapper(b, n, m, matrix, temp, idx):
    need(matrix, b * n * n, "matrix"); need(temp, b * n, "temp"); need(idx, b * m, "idx")
    call("de6d_furthest_point_sampling_matrix", b, n, m, dev(matrix, "matrix", f32), dev(temp, "temp", f32),
         dev(idx, "idx", i32), stream_ptr())
    return 1


def furthest_point_sampling_weights_wrapper(b, n, m, xyz, weights, temp, idx):
    need(xyz, b * n * 3, "xyz"); need(weights, b * n, "weights"); need(temp, b * n, "temp"); need(idx, b * m, "idx")
    call("de6d_furthest_point_sampling_weights", b, n, m, dev(xyz, "xyz", f32), dev(weights, "weights", f32),
         dev(temp, "temp", f32), dev(idx, "idx", i32), stream_ptr())
    return 1


def three_nn_wrapper(b, n, m, unknown, known, dist2, idx, impl=0):
    need(unknown, b * n * 3, "unknown"); need(known, b * m * 3, "known"); need(dist2, b * n * 3, "dist2"); need(idx, b * n * 3, "idx")
    ws_bytes = int(load().de6d_ball_query_workspace_bytes(b, m)) if impl == 0 else 0   # grid over `known` for m >= 2048
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=known.device) if ws_bytes else None
    call("de6d_three_nn_ex", impl, b, n, m, dev(unknown, "unknown", f32), dev(known, "known", f32), dev(dist2, "dist2", f32),
         dev(idx, "idx", i32), None if ws is None else ws.data_ptr(), ws_bytes, stream_ptr())
    return 1


def three_interpolate_wrapper(b, c, m, n, points, idx, weight, out):
    need(points, b * c * m, "points"); need(idx, b * n * 3, "idx"); need(weight, b * n * 3, "weight"); need(out, b * c * n, "out")
    call("de6d_three_interpolate", b, c, m, n, dev(points, "points", f32), dev(idx, "idx", i32),
         dev(weight, "weight", f32), dev(out, "out", f32), stream_ptr())
    return 1


def three_interpolate_grad_wrapper(b, c, n, m, grad_out, idx, weight, grad_points):
    need(grad_out, b * c * n, "grad_out"); need(idx, b * n * 3, "idx"); need(weight, b * n * 3, "weight"); need(grad_points, b * c * m, "grad_points")
    call("de6d_three_interpolate_grad", b, c, n, m, dev(grad_out, "grad_out", f32), dev(idx, "idx", i32),
         dev(weight, "weight", f32), dev(grad_points, "grad_points", f32), stream_ptr())
    return 1
