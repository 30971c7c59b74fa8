"""Replacement for the `points_in_boxes_*` entries of the reference extension `roiaware_pool3d_cuda`
(roiaware_pool3d.cpp:172-177).  The RoI-aware pooling forward/backward entries belong to the PartA2 head,
are out of this path's scope and raise."""
import torch

from ._common import call, dev, stream_ptr

f32, i32 = torch.float32, torch.int32


def points_in_boxes_gpu(boxes_tensor, pts_tensor, box_idx_of_points_tensor):
    b, t = boxes_tensor.size(0), boxes_tensor.size(1)
    m = pts_tensor.size(1)
    if pts_tensor.size(0) != b or box_idx_of_points_tensor.numel() < b * m:
        raise ValueError("points_in_boxes_gpu: inconsistent shapes")
    call("de6d_points_in_boxes", b, t, m, dev(boxes_tensor, "boxes", f32), dev(pts_tensor, "pts", f32),
         dev(box_idx_of_points_tensor, "box_idx_of_points", i32), stream_ptr())
    return 1


def points_in_boxes_cpu(boxes_tensor, pts_tensor, pts_indices_tensor):
    """Host tensors in, host tensor out, like the reference (roiaware_pool3d.cpp:143-168); the test itself
    runs on the B200 (de6d_points_in_boxes_mask)."""
    if boxes_tensor.is_cuda or pts_tensor.is_cuda or pts_indices_tensor.is_cuda:
        raise ValueError("points_in_boxes_cpu takes CPU tensors")
    t, m = boxes_tensor.size(0), pts_tensor.size(0)
    bx = boxes_tensor.to(device="cuda", dtype=f32).contiguous()
    pt = pts_tensor.to(device="cuda", dtype=f32).contiguous()
    out = torch.zeros((t, m), dtype=i32, device="cuda")
    call("de6d_points_in_boxes_mask", t, m, bx.data_ptr(), pt.data_ptr(), out.data_ptr(), stream_ptr())
    pts_indices_tensor.copy_(out)
    return 1


def forward(*args, **kwargs):
    raise NotImplementedError("roiaware_pool3d forward (PartA2 RoI head) is outside the Det6D hot path")


def backward(*args, **kwargs):
    raise NotImplementedError("roiaware_pool3d backward (PartA2 RoI head) is outside the Det6D hot path")
