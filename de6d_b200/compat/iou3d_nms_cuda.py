"""Replacement for the reference extension `iou3d_nms_cuda` (iou3d_nms_api.cpp:11-17)."""
import torch

from .._lib import load
from ._common import call, dev, stream_ptr

f32 = torch.float32
HOST_THREADS = 1   # the reference loops are single-threaded; raise for bulk host-side use outside DataLoader workers


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


def boxes_iou3d9_gpu(boxes_a, boxes_b, ans_iou):
    """Extension of this package: full-pose IoU of (N, 9) x (M, 9) boxes [x, y, z, dx, dy, dz, rz, ry, rx]."""
    for t, name in ((boxes_a, "boxes_a"), (boxes_b, "boxes_b")):
        if t.dim() != 2 or t.size(1) != 9:
            raise ValueError("%s must have shape (N, 9)" % name)
    if ans_iou.numel() < boxes_a.size(0) * boxes_b.size(0):
        raise ValueError("ans_iou is too small")
    call("de6d_boxes_iou3d9", boxes_a.size(0), dev(boxes_a, "boxes_a", f32), boxes_b.size(0), dev(boxes_b, "boxes_b", f32),
         dev(ans_iou, "ans_iou", f32), stream_ptr())
    return 1


def boxes_iou_bev_cpu(boxes_a, boxes_b, ans_iou):
    """Host tensors in, host tensor out, evaluated on the calling host thread like the reference
    (iou3d_cpu.cpp:232-252) -- its callers run in forked DataLoader workers (database_sampler.py:232-233) where no CUDA
    context may be created.  de6d_boxes_iou_bev_host: the reference's host arithmetic, bit-identical results.
    (Callers that hold device tensors use boxes_iou_bev_gpu.)"""
    if boxes_a.is_cuda or boxes_b.is_cuda or ans_iou.is_cuda:
        raise ValueError("boxes_iou_bev_cpu takes CPU tensors")
    for t, name in ((boxes_a, "boxes_a"), (boxes_b, "boxes_b"), (ans_iou, "ans_iou")):
        if t.dtype != f32:
            raise TypeError("%s must be float32, got %s" % (name, t.dtype))
        if not t.is_contiguous():
            raise ValueError("%s must be contiguous" % name)
    if boxes_a.dim() != 2 or boxes_a.size(1) != 7 or boxes_b.dim() != 2 or boxes_b.size(1) != 7:
        raise ValueError("boxes must have shape (N, 7)")
    if ans_iou.numel() < boxes_a.size(0) * boxes_b.size(0):
        raise ValueError("ans_iou is too small")
    call("de6d_boxes_iou_bev_host", boxes_a.size(0), boxes_a.data_ptr(), boxes_b.size(0), boxes_b.data_ptr(),
         ans_iou.data_ptr(), HOST_THREADS)
    return 1


def _nms(boxes, keep, thresh, mode):
    n = boxes.size(0)
    width = 9 if mode == 2 else 7
    if boxes.dim() != 2 or boxes.size(1) != width:
        raise ValueError("boxes must have shape (N, %d)" % width)
    ptr = dev(boxes, "boxes", f32)
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
    call("de6d_nms_batched", 1, n, ptr, None, float(thresh), int(mode), keep_dev.data_ptr(), num.data_ptr(),
         ws.data_ptr(), ws_bytes, s)
    num_out = int(num.item())  # the reference API returns a host int: this sync is part of its contract
    keep[:num_out].copy_(keep_dev[:num_out])
    return num_out


def nms9_gpu(boxes, keep, nms_overlap_thresh):
    """Extension of this package: nms_gpu's contract with the full-pose IoU on (N, 9) boxes."""
    return _nms(boxes, keep, nms_overlap_thresh, 2)


def nms_gpu(boxes, keep, nms_overlap_thresh):
    return _nms(boxes, keep, nms_overlap_thresh, 0)


def nms_normal_gpu(boxes, keep, nms_overlap_thresh):
    return _nms(boxes, keep, nms_overlap_thresh, 1)
