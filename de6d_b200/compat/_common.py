"""Tensor checks and pointer plumbing shared by the compat modules."""
import torch

from .._lib import call  # noqa: F401  (re-exported)


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def dev(t, name, dtype):
    """CHECK_INPUT of the reference (ball_query.cpp:17-29, iou3d_nms.cpp:14-26) as an exception: CUDA,
    contiguous, expected dtype, on the current device (kernels are launched on the current stream)."""
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise ValueError("%s must be a CUDA tensor" % name)
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if t.device.index != torch.cuda.current_device():
        raise ValueError("%s is on cuda:%d but the current device is cuda:%d" % (name, t.device.index, torch.cuda.current_device()))
    return t.data_ptr()


def need(t, numel, name):
    if t.numel() < numel:
        raise ValueError("%s has %d elements, kernel needs %d" % (name, t.numel(), numel))
