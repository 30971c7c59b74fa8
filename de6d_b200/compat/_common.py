"""Tensor checks and pointer plumbing shared by the compat modules.  Kept lean: in eager mode the reference calls these
ops thousands of times per second and a wrapper that costs more host time than the kernel takes on the device would
cap the speed-up (the raw-stream / device queries below are the C-level calls behind torch.cuda.current_stream())."""
import torch

from .._lib import call  # noqa: F401  (re-exported)

_get_device = torch._C._cuda_getDevice if hasattr(torch._C, "_cuda_getDevice") else torch.cuda.current_device
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr():
    if _raw_stream is not None:
        return _raw_stream(_get_device())
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
    if t.get_device() != _get_device():
        raise ValueError("%s is on cuda:%d but the current device is cuda:%d" % (name, t.get_device(), _get_device()))
    return t.data_ptr()


def need(t, numel, name):
    if t.numel() < numel:
        raise ValueError("%s has %d elements, kernel needs %d" % (name, t.numel(), numel))
