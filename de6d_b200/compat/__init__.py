"""Drop-in replacements for the reference's three pybind11 extension modules, same function names and
positional signatures:

    pcdet.ops.pointnet2.pointnet2_batch.pointnet2_batch_cuda  ->  de6d_b200.compat.pointnet2_batch_cuda
    pcdet.ops.iou3d_nms.iou3d_nms_cuda                        ->  de6d_b200.compat.iou3d_nms_cuda
    pcdet.ops.roiaware_pool3d.roiaware_pool3d_cuda            ->  de6d_b200.compat.roiaware_pool3d_cuda

`install()` registers them in sys.modules under the reference's module paths so an unmodified checkout of
the reference (pointnet2_utils.py, iou3d_nms_utils.py, roiaware_pool3d_utils.py and everything above them)
imports these instead of its own extensions.  See INTEGRATION.md.
"""
import sys

from . import iou3d_nms_cuda, pointnet2_batch_cuda, roiaware_pool3d_cuda

_TARGETS = {
    "pcdet.ops.pointnet2.pointnet2_batch.pointnet2_batch_cuda": pointnet2_batch_cuda,
    "pcdet.ops.iou3d_nms.iou3d_nms_cuda": iou3d_nms_cuda,
    "pcdet.ops.roiaware_pool3d.roiaware_pool3d_cuda": roiaware_pool3d_cuda,
}


def install():
    """Make `from . import pointnet2_batch_cuda` (etc.) inside the reference resolve to this package."""
    for name, mod in _TARGETS.items():
        sys.modules[name] = mod
    return dict(_TARGETS)
