"""Mirror of the full-pose box helper of pcdet/utils/box_utils.py that Det6D's head calls during target assignment
(point_head_box6d_vote.py:198-209,284-286) -- on the device instead of host numpy + scipy Delaunay."""
import numpy as np
import torch

from ._lib import call


def points_in_boxes3d(points, boxes3d):
    """(reference box_utils.py:110-124) points (n, 3+C), boxes3d (m, 9) [x, y, z, dx, dy, dz, rz, ry, rx] ->
    (n,) int64: index of the last box containing the point, -1 if none.  numpy in -> numpy out; torch in -> torch out
    on the points' device."""
    assert boxes3d.shape[-1] == 9
    is_numpy = isinstance(points, np.ndarray)
    if is_numpy:
        pts = torch.from_numpy(np.ascontiguousarray(points[:, :3], dtype=np.float32)).cuda()
        bx = torch.from_numpy(np.ascontiguousarray(boxes3d, dtype=np.float32)).cuda()
    else:
        pts = points[:, :3].detach().to(device="cuda", dtype=torch.float32).contiguous()
        bx = torch.as_tensor(boxes3d).detach().to(device="cuda", dtype=torch.float32).contiguous()
    flags = points_in_boxes3d_batched(pts.unsqueeze(0), bx.unsqueeze(0))[0]
    return flags.cpu().numpy() if is_numpy else flags.to(points.device)


@torch.no_grad()
def points_in_boxes3d_batched(points, boxes3d):
    """points (B, n, 3), boxes3d (B, m, 9) CUDA float32 -> (B, n) int64, one launch for the whole batch."""
    assert points.is_cuda and boxes3d.is_cuda and points.dtype == torch.float32 and boxes3d.dtype == torch.float32
    assert points.shape[0] == boxes3d.shape[0] and points.shape[2] == 3 and boxes3d.shape[2] == 9
    from .compat._common import dev, stream_ptr
    points, boxes3d = points.contiguous(), boxes3d.contiguous()
    B, n, _ = points.shape
    out = torch.empty((B, n), dtype=torch.int64, device=points.device)
    call("de6d_points_in_boxes9", B, boxes3d.shape[1], n, dev(boxes3d, "boxes3d", torch.float32), dev(points, "points", torch.float32),
         out.data_ptr(), stream_ptr())
    return out
