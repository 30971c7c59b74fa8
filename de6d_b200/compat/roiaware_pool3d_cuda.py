"""Replacement for the `points_in_boxes_*` entries of the reference extension `roiaware_pool3d_cuda`
(roiaware_pool3d.cpp:172-177).  The RoI-aware pooling forward/backward entries belong to the PartA2 head,
are out of this path's scope and raise."""
import torch

from ._common import call, dev, stream_ptr

f32, i32 = torch.float32, torch.int32
HOST_THREADS = 1


def points_in_boxes_gpu(boxes_tensor, pts_tensor, box_idx_of_points_tensor):
    b, t = boxes_tensor.size(0), boxes_tensor.size(1)
    m = pts_tensor.size(1)
    if pts_tensor.size(0) != b or box_idx_of_points_tensor.numel() < b * m:
        raise ValueError("points_in_boxes_gpu: inconsistent shapes")
    call("de6d_points_in_boxes", b, t, m, dev(boxes_tensor, "boxes", f32), dev(pts_tensor, "pts", f32),
         dev(box_idx_of_points_tensor, "box_idx_of_points", i32), stream_ptr())
    return 1


def points_in_boxes_cpu(boxes_tensor, pts_tensor, pts_indices_tensor):
    """Host tensors in, host tensor out, evaluated on the calling host thread like the reference
    (roiaware_pool3d.cpp:143-168): its callers (kitti_dataset.py:248, box_utils.py:104, database_sampler) run inside
    forked DataLoader workers.  de6d_points_in_boxes_mask_host, bit-identical to the reference function.
    (The device twin for resident tensors is de6d_points_in_boxes_mask / roiaware_pool3d_utils.points_in_boxes_mask_gpu.)"""
    if boxes_tensor.is_cuda or pts_tensor.is_cuda or pts_indices_tensor.is_cuda:
        raise ValueError("points_in_boxes_cpu takes CPU tensors")
    t, m = boxes_tensor.size(0), pts_tensor.size(0)
    if boxes_tensor.dtype != f32 or pts_tensor.dtype != f32 or pts_indices_tensor.dtype != i32:
        raise TypeError("points_in_boxes_cpu: boxes / pts must be float32 and pts_indices int32")
    if boxes_tensor.dim() != 2 or boxes_tensor.size(1) != 7 or pts_tensor.dim() != 2 or pts_tensor.size(1) != 3:
        raise ValueError("points_in_boxes_cpu: boxes (T, 7), pts (M, 3)")
    if pts_indices_tensor.numel() < t * m or not pts_indices_tensor.is_contiguous():
        raise ValueError("pts_indices must be a contiguous (T, M) int tensor")
    # the reference reads .data<float>() without a contiguity check (its CHECK_CONTIGUOUS lines are commented out);
    # a strided view would be misread there, here it is made contiguous first
    call("de6d_points_in_boxes_mask_host", t, m, boxes_tensor.contiguous().data_ptr(), pts_tensor.contiguous().data_ptr(),
         pts_indices_tensor.data_ptr(), HOST_THREADS)
    return 1


def forward(*args, **kwargs):
    raise NotImplementedError("roiaware_pool3d forward (PartA2 RoI head) is outside the Det6D hot path")


def backward(*args, **kwargs):
    raise NotImplementedError("roiaware_pool3d backward (PartA2 RoI head) is outside the Det6D hot path")
