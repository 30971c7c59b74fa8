"""Host-side mirror of the reference op wrappers pcdet/ops/pointnet2/pointnet2_batch/pointnet2_utils.py:
the same public names, argument meaning and return values, on top of de6d_b200.compat.pointnet2_batch_cuda
(hand-written sm_100a kernels behind the C ABI).  Output/scratch tensors are allocated here with the same
pre-initialisation the reference relies on (temp = 1e10, idx = 0, idx_cnt = 0, grads = 0).
"""
from typing import Tuple

import torch
import torch.nn as nn
from torch.autograd import Function

from .compat import pointnet2_batch_cuda as _ext


def _new(ref: torch.Tensor, shape, dtype):
    return torch.empty(shape, dtype=dtype, device=ref.device)


class FarthestPointSampling(Function):
    """D-FPS (reference :10-33).  xyz (B, N, 3) -> (B, npoint) int32; index 0 is always selected first."""

    @staticmethod
    def forward(ctx, xyz: torch.Tensor, npoint: int) -> torch.Tensor:
        assert xyz.is_contiguous()
        B, N, _ = xyz.size()
        out = _new(xyz, (B, npoint), torch.int32)
        temp = _new(xyz, (B, N), torch.float32).fill_(1e10)
        _ext.farthest_point_sampling_wrapper(B, N, npoint, xyz, temp, out)
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, a=None):
        return None, None


farthest_point_sample = furthest_point_sample = FarthestPointSampling.apply


def _strided_features(features, B, N, device):
    """(B, N, C) float32 CUDA view with arbitrary non-negative strides on `device` (the permuted (B, C, N) backbone tensor)."""
    if not (isinstance(features, torch.Tensor) and features.is_cuda and features.dtype == torch.float32):
        raise TypeError("features must be a float32 CUDA tensor")
    if features.dim() != 3 or tuple(features.shape[:2]) != (B, N) or features.device != device:
        raise ValueError("features must have shape (B, N, C) on the device of xyz")
    if min(features.stride()) < 0:
        raise ValueError("features must not have negative strides")
    return features.data_ptr()


@torch.no_grad()
def calc_dist_matrix_for_sampling(xyz: torch.Tensor, features: torch.Tensor = None, gamma: float = 1.0):
    """F-FPS input (reference :36-44): pairwise L2 of coordinates plus gamma * pairwise L2 of features, (B, N, N).
    xyz (B, N, 3), features (B, N, C) -- any strides, e.g. the permuted view of a (B, C, N) tensor.  One kernel with
    direct differences instead of the reference's two torch.cdist (GEMM expansion) + scale + add passes; values
    agree with torch.cdist to its own rounding error (~1e-3 abs for close points, where the expansion cancels)."""
    from ._lib import call
    from .compat._common import stream_ptr
    assert xyz.is_cuda and xyz.dtype == torch.float32 and xyz.dim() == 3 and xyz.size(2) == 3
    xyz = xyz.contiguous()
    B, N, _ = xyz.shape
    px = _chk(xyz, "xyz", torch.float32)
    out = torch.empty((B, N, N), dtype=torch.float32, device=xyz.device)
    if features is not None:
        fptr, (sb, sn, sc) = _strided_features(features, B, N, xyz.device), features.stride()
        call("de6d_dist_matrix", B, N, features.size(2), px, fptr, sb, sn, sc, float(gamma), out.data_ptr(), stream_ptr())
    else:
        call("de6d_dist_matrix", B, N, 0, px, None, 0, 0, 0, float(gamma), out.data_ptr(), stream_ptr())
    return out


@torch.no_grad()
def furthest_point_sample_matrix(matrix: torch.Tensor, npoint: int) -> torch.Tensor:
    """F-FPS on a (B, N, N) distance matrix (reference :69-86)."""
    assert matrix.is_contiguous()
    B, N, _ = matrix.size()
    out = _new(matrix, (B, npoint), torch.int32)
    temp = _new(matrix, (B, N), torch.float32).fill_(1e10)
    _ext.furthest_point_sampling_matrix_wrapper(B, N, npoint, matrix, temp, out)
    return out


@torch.no_grad()
def furthest_point_sample_features(xyz: torch.Tensor, features: torch.Tensor, gamma: float, npoint: int,
                                   cluster_size: int = 0, prune: int = 0) -> torch.Tensor:
    """F-FPS without the (B, N, N) matrix: identical indices to
        furthest_point_sample_matrix(calc_dist_matrix_for_sampling(xyz, features, gamma), npoint)
    (the reference's call pair, pointnet2_modules.py:383-388).  xyz (B, N, 3), features (B, N, C) with any strides.
    One thread-block cluster (6 or 8 CTAs) per cloud evaluates only the selected rows out of distributed shared memory; shapes that do not
    fit on chip take the two-call form.  cluster_size: 0 = automatic, 6 / 8 pin the cluster size; prune: 0 = automatic, 1 = dense
    kernel, 2 = pruned kernel (64-point buckets skipped when their bounding box proves nothing can change), 3 = pruned kernel with the
    remaining buckets evaluated by the whole CTA -- tests, tuning."""
    from ._lib import call, load
    from .compat._common import stream_ptr
    B, N, _ = xyz.shape
    px = _chk(xyz, "xyz", torch.float32, (B, N, 3))
    C = 0 if features is None else features.size(2)
    if not load().de6d_furthest_point_sampling_features_fits(N, C):
        return furthest_point_sample_matrix(calc_dist_matrix_for_sampling(xyz, features, gamma), npoint)
    out = _new(xyz, (B, npoint), torch.int32)
    temp = _new(xyz, (B, N), torch.float32).fill_(1e10)
    if features is None:
        fptr, (sb, sn, sc) = None, (0, 0, 0)
    else:
        fptr, (sb, sn, sc) = _strided_features(features, B, N, xyz.device), features.stride()
    if cluster_size or prune:
        call("de6d_furthest_point_sampling_features_impl", B, N, C, npoint, px, fptr, sb, sn, sc, float(gamma),
             temp.data_ptr(), out.data_ptr(), int(cluster_size), int(prune), stream_ptr())
    else:
        call("de6d_furthest_point_sampling_features", B, N, C, npoint, px, fptr, sb, sn, sc, float(gamma),
             temp.data_ptr(), out.data_ptr(), stream_ptr())
    return out


@torch.no_grad()
def furthest_point_sample_weights(xyz: torch.Tensor, weights: torch.Tensor, npoint: int) -> torch.Tensor:
    """S-FPS (reference :89-109): first pick = argmax(weights), then argmax of min-dist * max(w, 1e-12)."""
    assert xyz.is_contiguous()
    assert weights.is_contiguous()
    B, N, _ = xyz.size()
    out = _new(xyz, (B, npoint), torch.int32)
    temp = _new(xyz, (B, N), torch.float32).fill_(1e10)
    _ext.furthest_point_sampling_weights_wrapper(B, N, npoint, xyz, weights, temp, out)
    return out


class GatherOperation(Function):
    """features (B, C, N), idx (B, npoint) -> (B, C, npoint) (reference :115-146)."""

    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        assert features.is_contiguous()
        assert idx.is_contiguous()
        B, npoint = idx.size()
        _, C, N = features.size()
        out = _new(features, (B, C, npoint), torch.float32)
        _ext.gather_points_wrapper(B, C, N, npoint, features, idx, out)
        ctx.for_backwards = (idx, C, N)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, C, N = ctx.for_backwards
        B, npoint = idx.size()
        grad_features = torch.zeros((B, C, N), dtype=torch.float32, device=grad_out.device)
        _ext.gather_points_grad_wrapper(B, C, N, npoint, grad_out.detach().contiguous(), idx, grad_features)
        return grad_features, None


gather_operation = GatherOperation.apply


class ThreeNN(Function):
    """unknown (B, N, 3), known (B, M, 3) -> (sqrt of the 3 smallest squared distances, their indices)
    (reference :152-178)."""

    @staticmethod
    def forward(ctx, unknown: torch.Tensor, known: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        assert unknown.is_contiguous()
        assert known.is_contiguous()
        B, N, _ = unknown.size()
        m = known.size(1)
        dist2 = _new(unknown, (B, N, 3), torch.float32)
        idx = _new(unknown, (B, N, 3), torch.int32)
        _ext.three_nn_wrapper(B, N, m, unknown, known, dist2, idx)
        dist = torch.sqrt(dist2)
        ctx.mark_non_differentiable(dist, idx)
        return dist, idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = ThreeNN.apply


class ThreeInterpolate(Function):
    """features (B, C, M), idx/weight (B, n, 3) -> (B, C, n) (reference :184-226)."""

    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
        assert features.is_contiguous()
        assert idx.is_contiguous()
        assert weight.is_contiguous()
        B, c, m = features.size()
        n = idx.size(1)
        ctx.three_interpolate_for_backward = (idx, weight, m)
        out = _new(features, (B, c, n), torch.float32)
        _ext.three_interpolate_wrapper(B, c, m, n, features, idx, weight, out)
        return out

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        idx, weight, m = ctx.three_interpolate_for_backward
        B, c, n = grad_out.size()
        grad_features = torch.zeros((B, c, m), dtype=torch.float32, device=grad_out.device)
        _ext.three_interpolate_grad_wrapper(B, c, n, m, grad_out.detach().contiguous(), idx, weight, grad_features)
        return grad_features, None, None


three_interpolate = ThreeInterpolate.apply


class GroupingOperation(Function):
    """features (B, C, N), idx (B, npoint, nsample) -> (B, C, npoint, nsample) (reference :232-270)."""

    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        assert features.is_contiguous()
        assert idx.is_contiguous()
        B, nfeatures, nsample = idx.size()
        _, C, N = features.size()
        out = _new(features, (B, C, nfeatures, nsample), torch.float32)
        _ext.group_points_wrapper(B, C, N, nfeatures, nsample, features, idx, out)
        ctx.for_backwards = (idx, N)
        return out

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        idx, N = ctx.for_backwards
        B, C, npoint, nsample = grad_out.size()
        grad_features = torch.zeros((B, C, N), dtype=torch.float32, device=grad_out.device)
        _ext.group_points_grad_wrapper(B, C, N, npoint, nsample, grad_out.detach().contiguous(), idx, grad_features)
        return grad_features, None


grouping_operation = GroupingOperation.apply


class BallQuery(Function):
    """(reference :276-301) first `nsample` neighbours (ascending index) within `radius`, padded with the first."""

    @staticmethod
    def forward(ctx, radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
        assert new_xyz.is_contiguous()
        assert xyz.is_contiguous()
        B, N, _ = xyz.size()
        npoint = new_xyz.size(1)
        idx = torch.zeros((B, npoint, nsample), dtype=torch.int32, device=xyz.device)
        _ext.ball_query_wrapper(B, N, npoint, radius, nsample, new_xyz, xyz, idx)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


ball_query = BallQuery.apply


class BallQueryGrid:
    """Search grid of a batch of clouds, built once and shared by every ball query over the same `xyz` -- the radius
    scales of one SA layer (pointnet2_modules.py:462-463 runs one grouper per scale on the same cloud).

        grid = BallQueryGrid(xyz, min(radii))                         # one build kernel
        idx_cnt, idx = ball_query_cnt(r, ns, xyz, new_xyz, grid=grid)   # per scale: query kernel only

    Results are identical with and without a grid (tests/test_parity_gpu.py).  Clouds below 2048 points are answered by
    the brute-force kernel, which needs no grid: `BallQueryGrid.wanted(n)` tells."""

    MIN_N = 2048

    @staticmethod
    def wanted(n):
        return n >= BallQueryGrid.MIN_N

    def __init__(self, xyz: torch.Tensor, radius: float):
        from ._lib import load
        assert xyz.is_cuda and xyz.is_contiguous() and xyz.dtype == torch.float32 and xyz.dim() == 3 and xyz.size(2) == 3
        self.b, self.n = xyz.size(0), xyz.size(1)
        self.xyz_ptr = xyz.data_ptr()
        self.ws = torch.empty(int(load().de6d_ball_query_grid_bytes(self.b, self.n)), dtype=torch.uint8, device=xyz.device)
        _ext.ball_query_grid_build(self.b, self.n, radius, xyz, self.ws)

    def check(self, xyz):
        if xyz.data_ptr() != self.xyz_ptr or xyz.size(0) != self.b or xyz.size(1) != self.n:
            raise ValueError("BallQueryGrid was built for a different xyz tensor")
        return self.ws


@torch.no_grad()
def ball_query_cnt(radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor, grid: "BallQueryGrid" = None):
    """(reference :307-327) -> (idx_cnt (B, npoint), idx (B, npoint, nsample)); hit list repeated cyclically.
    grid (not in the reference): a BallQueryGrid of `xyz` to reuse."""
    assert new_xyz.is_contiguous()
    assert xyz.is_contiguous()
    B, N, _ = xyz.size()
    npoint = new_xyz.size(1)
    idx = torch.zeros((B, npoint, nsample), dtype=torch.int32, device=xyz.device)
    idx_cnt = torch.zeros((B, npoint), dtype=torch.int32, device=xyz.device)
    if grid is None:
        _ext.ball_query_cnt_wrapper(B, N, npoint, radius, nsample, new_xyz, xyz, idx_cnt, idx)
    else:
        _ext._ball_query(1, B, N, npoint, 0.0, radius, nsample, new_xyz, xyz, idx_cnt, idx, grid_ws=grid.check(xyz))
    return idx_cnt, idx


@torch.no_grad()
def ball_query_dilated(radius_in: float, radius_out: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor,
                       grid: "BallQueryGrid" = None):
    """(reference :330-351) shell radius_in <= d < radius_out."""
    assert new_xyz.is_contiguous()
    assert xyz.is_contiguous()
    B, N, _ = xyz.size()
    npoint = new_xyz.size(1)
    idx_cnt = torch.zeros((B, npoint), dtype=torch.int32, device=xyz.device)
    idx = torch.zeros((B, npoint, nsample), dtype=torch.int32, device=xyz.device)
    if grid is None:
        _ext.ball_query_dilated_wrapper(B, N, npoint, radius_in, radius_out, nsample, new_xyz, xyz, idx_cnt, idx)
    else:
        _ext._ball_query(2, B, N, npoint, radius_in, radius_out, nsample, new_xyz, xyz, idx_cnt, idx, grid_ws=grid.check(xyz))
    return idx_cnt, idx


def _chk(t, name, dtype, shape=None):
    """CHECK_INPUT for the entry points that are not in the reference's pybind table: CUDA, contiguous, dtype, on the
    current device (the kernel is launched on the current stream), optional exact shape.  Returns data_ptr()."""
    from .compat._common import dev
    ptr = dev(t, name, dtype)
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError("%s must have shape %s, got %s" % (name, tuple(shape), tuple(t.shape)))
    return ptr


def group_concat(xyz, new_xyz, features, idx, xyz_t=None):
    """cat(xyz[idx] - new_xyz, features[idx]) -> (B, 3 + C, npoint, nsample) in one kernel pass (no autograd).
    Bit-identical to the reference composition: copies plus one fp32 subtraction.  xyz_t: optional (B, 3, N) transposed
    copy of xyz (the SA module's `xyz_flipped`): coordinate rows are then TMA-staged like the channels."""
    from ._lib import call
    from .compat._common import stream_ptr
    if xyz.dim() != 3 or idx.dim() != 3:
        raise ValueError("group_concat: xyz (B, N, 3), idx (B, npoint, nsample)")
    B, N, _ = xyz.shape
    _, M, ns = idx.shape
    C = 0 if features is None else features.size(1)
    out = torch.empty((B, 3 + C, M, ns), dtype=torch.float32, device=xyz.device)
    call("de6d_group_concat_t", B, C, N, M, ns, _chk(xyz, "xyz", torch.float32, (B, N, 3)),
         None if xyz_t is None else _chk(xyz_t, "xyz_t", torch.float32, (B, 3, N)),
         _chk(new_xyz, "new_xyz", torch.float32, (B, M, 3)),
         None if features is None else _chk(features, "features", torch.float32, (B, C, N)),
         _chk(idx, "idx", torch.int32, (B, M, ns)), out.data_ptr(), stream_ptr())
    return out


@torch.no_grad()
def gather_xyz(xyz, sample_idx, want_transposed=False):
    """new_xyz = xyz[sample_idx]: (B, N, 3), (B, M) int32 -> (B, M, 3) [and (B, 3, M)] with one launch instead of the SA
    module's transpose -> gather_operation -> transpose (pointnet2_modules.py:374,451-454).  Pure copies.
    sample_idx None: only the transposed copy of xyz itself is produced (returns (xyz, xyz_t))."""
    from ._lib import call
    from .compat._common import stream_ptr
    B, N, _ = xyz.shape
    px = _chk(xyz, "xyz", torch.float32, (B, N, 3))
    if sample_idx is None:
        xyz_t = torch.empty((B, 3, N), dtype=torch.float32, device=xyz.device)
        call("de6d_gather_xyz", B, N, N, px, None, None, xyz_t.data_ptr(), stream_ptr())
        return xyz, xyz_t
    M = sample_idx.size(1)
    new_xyz = torch.empty((B, M, 3), dtype=torch.float32, device=xyz.device)
    new_t = torch.empty((B, 3, M), dtype=torch.float32, device=xyz.device) if want_transposed else None
    call("de6d_gather_xyz", B, N, M, px, _chk(sample_idx, "sample_idx", torch.int32, (B, M)), new_xyz.data_ptr(),
         None if new_t is None else new_t.data_ptr(), stream_ptr())
    return (new_xyz, new_t) if want_transposed else new_xyz


def _needs_grad(*ts):
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in ts)


def _assemble(xyz, new_xyz, features, idx, use_xyz):
    """Shared tail of the three grouper modules (reference :368-387, :410-424, :449-463).  Inference (nothing
    requires grad) with use_xyz takes the fused single-pass kernel; otherwise the reference composition, whose
    autograd graph (grouping_operation backward) is kept as is."""
    if use_xyz and xyz.dtype == torch.float32 and not _needs_grad(xyz, new_xyz, features) and \
            (features is None or features.dtype == torch.float32):
        return group_concat(xyz.contiguous(), new_xyz.contiguous(), None if features is None else features.contiguous(), idx)
    xyz_trans = xyz.transpose(1, 2).contiguous()
    grouped_xyz = grouping_operation(xyz_trans, idx)  # (B, 3, npoint, nsample)
    grouped_xyz -= new_xyz.transpose(1, 2).unsqueeze(-1)
    if features is not None:
        grouped_features = grouping_operation(features, idx)
        if use_xyz:
            return torch.cat([grouped_xyz, grouped_features], dim=1)  # (B, 3 + C, npoint, nsample)
        return grouped_features
    assert use_xyz, "Cannot have not features and not use xyz as a feature!"
    return grouped_xyz


class QueryAndGroup(nn.Module):
    """(reference :354-387)"""

    def __init__(self, radius: float, nsample: int, use_xyz: bool = True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor, features: torch.Tensor = None):
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        return _assemble(xyz, new_xyz, features, idx, self.use_xyz)


class QueryWithCntAndGroup(nn.Module):
    """(reference :390-424) -- the grouper Det6D / SASA use."""

    def __init__(self, radius: float, nsample: int, use_xyz: bool = True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor, features: torch.Tensor = None):
        idx_cnt, idx = ball_query_cnt(self.radius, self.nsample, xyz, new_xyz)
        return idx_cnt, _assemble(xyz, new_xyz, features, idx, self.use_xyz)


class QueryAndGroupDilated(nn.Module):
    """(reference :427-463)"""

    def __init__(self, radius_in: float, radius_out: float, nsample: int, use_xyz: bool = True):
        super().__init__()
        self.radius_in, self.radius_out, self.nsample, self.use_xyz = radius_in, radius_out, nsample, use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor, features: torch.Tensor = None):
        idx_cnt, idx = ball_query_dilated(self.radius_in, self.radius_out, self.nsample, xyz, new_xyz)
        return idx_cnt, _assemble(xyz, new_xyz, features, idx, self.use_xyz)


class GroupAll(nn.Module):
    """(reference :466-488) no kernel involved."""

    def __init__(self, use_xyz: bool = True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz: torch.Tensor, new_xyz: torch.Tensor, features: torch.Tensor = None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is None:
            return grouped_xyz
        grouped_features = features.unsqueeze(2)
        if self.use_xyz:
            return torch.cat([grouped_xyz, grouped_features], dim=1)
        return grouped_features
