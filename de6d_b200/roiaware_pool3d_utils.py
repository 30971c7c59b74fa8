"""Host-side mirror of the `points_in_boxes_*` functions of pcdet/ops/roiaware_pool3d/roiaware_pool3d_utils.py."""
import numpy as np
import torch

from .compat import roiaware_pool3d_cuda as _ext


def _to_torch(x):
    if isinstance(x, np.ndarray):
        return torch.from_numpy(x).float(), True
    return x, False


def points_in_boxes_cpu(points, boxes):
    """(reference :9-25) points (M, 3), boxes (T, 7), numpy or CPU tensors -> (T, M) int32 0/1 mask.  Evaluated on the
    calling host thread like the reference (safe inside forked DataLoader workers); bit-identical to it."""
    assert boxes.shape[1] == 7
    assert points.shape[1] == 3
    points, is_numpy = _to_torch(points)
    boxes, is_numpy = _to_torch(boxes)
    point_indices = points.new_zeros((boxes.shape[0], points.shape[0]), dtype=torch.int)
    _ext.points_in_boxes_cpu(boxes.float().contiguous(), points.float().contiguous(), point_indices)
    return point_indices.numpy() if is_numpy else point_indices


def points_in_boxes_gpu(points, boxes):
    """(reference :28-41) points (B, M, 3), boxes (B, T, 7) -> (B, M) int32 index of the first containing box, -1 if none."""
    assert boxes.shape[0] == points.shape[0]
    assert boxes.shape[2] == 7 and points.shape[2] == 3
    batch_size, num_points, _ = points.shape
    box_idxs_of_pts = points.new_zeros((batch_size, num_points), dtype=torch.int).fill_(-1)
    _ext.points_in_boxes_gpu(boxes.contiguous(), points.contiguous(), box_idxs_of_pts)
    return box_idxs_of_pts


def points_in_boxes_mask_gpu(points, boxes):
    """Device twin of points_in_boxes_cpu for tensors already resident on the GPU: points (M, 3), boxes (T, 7) CUDA
    float32 -> (T, M) int32 0/1 mask with the same MARGIN 1e-2 arithmetic (not in the reference)."""
    from ._lib import call
    from .compat._common import dev, stream_ptr
    assert boxes.shape[1] == 7 and points.shape[1] == 3
    boxes, points = boxes.contiguous(), points.contiguous()
    out = torch.zeros((boxes.shape[0], points.shape[0]), dtype=torch.int32, device=points.device)
    call("de6d_points_in_boxes_mask", boxes.shape[0], points.shape[0], dev(boxes, "boxes", torch.float32),
         dev(points, "points", torch.float32), out.data_ptr(), stream_ptr())
    return out
