"""Replacement for the reference extension `iou3d_nms_cuda` (iou3d_nms_api.cpp:11-17)."""
import torch

from .._lib import load
from ._common import call, dev, stream_ptr

f32 = torch.float32


def _boxes(t, name):
    if t.dim() != 2 or t.size(1) != 7:
        raise ValueError("%s must have shape (N, 7)" % name)
    return dev(t, name, f32)


def boxes_overlap_bev_gpu(boxes_a, boxes_b, ans_overlap):
    call("de6d_boxes_overlap_bev", boxes_a.size(0), _boxes(boxes_a, "boxes_a"), boxes_b.size(0), _boxes(boxes_b, "boxes_b"),
         dev(ans_overlap, "ans_overlap", f32), stream_ptr())
    return 1


def boxes_iou_bev_gpu(boxes_a, boxes_b, ans_iou):
    call("de6d_boxes_iou_bev", boxes_a.size(0), _boxes(boxes_a, "boxes_a"), boxes_b.size(0), _boxes(boxes_b, "boxes_b"),
         dev(ans_iou, "ans_iou", f32), stream_ptr())
    return 1


def boxes_iou3d_gpu(boxes_a, boxes_b, ans_iou):
    """Extension of this package: the fused form of iou3d_nms_utils.boxes_iou3d_gpu (python lines 48-81)."""
    call("de6d_boxes_iou3d", boxes_a.size(0), _boxes(boxes_a, "boxes_a"), boxes_b.size(0), _boxes(boxes_b, "boxes_b"),
         dev(ans_iou, "ans_iou", f32), stream_ptr())
    return 1


def boxes_iou_bev_cpu(boxes_a, boxes_b, ans_iou):
    """The reference evaluates this on the host (iou3d_cpu.cpp:232-252).  Here the same arithmetic runs on the
    B200: host tensors are staged to the device, computed with de6d_boxes_iou_bev and copied back."""
    if boxes_a.is_cuda or boxes_b.is_cuda or ans_iou.is_cuda:
        raise ValueError("boxes_iou_bev_cpu takes CPU tensors")
    if not (boxes_a.is_contiguous() and boxes_b.is_contiguous()):
        raise ValueError("boxes must be contiguous")
    a = boxes_a.to(device="cuda", dtype=f32)
    b = boxes_b.to(device="cuda", dtype=f32)
    out = torch.zeros((a.size(0), b.size(0)), dtype=f32, device="cuda")
    boxes_iou_bev_gpu(a, b, out)
    ans_iou.copy_(out)
    return 1


def _nms(boxes, keep, thresh, normal):
    n = boxes.size(0)
    ptr = _boxes(boxes, "boxes")
    if keep.is_cuda or keep.dtype != torch.int64 or not keep.is_contiguous():
        raise ValueError("keep must be a contiguous CPU LongTensor (reference contract, iou3d_nms_utils.py:97)")
    if keep.numel() < n:
        raise ValueError("keep has fewer elements than boxes")
    if n == 0:
        return 0
    lib = load()
    ws_bytes = lib.de6d_nms_workspace_bytes(1, n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=boxes.device)
    s = stream_ptr()
    call("de6d_nms_workspace_init", 1, ws.data_ptr(), s)
    keep_dev = torch.empty(n, dtype=torch.int64, device=boxes.device)
    num = torch.zeros(1, dtype=torch.int32, device=boxes.device)
    call("de6d_nms_batched", 1, n, ptr, None, float(thresh), int(normal), keep_dev.data_ptr(), num.data_ptr(),
         ws.data_ptr(), ws_bytes, s)
    num_out = int(num.item())  # the reference API returns a host int: this sync is part of its contract
    keep[:num_out].copy_(keep_dev[:num_out])
    return num_out


def nms_gpu(boxes, keep, nms_overlap_thresh):
    return _nms(boxes, keep, nms_overlap_thresh, 0)


def nms_normal_gpu(boxes, keep, nms_overlap_thresh):
    return _nms(boxes, keep, nms_overlap_thresh, 1)
