"""Input staging in front of the sampling path: host mirror of the reference's
`DataProcessor.sample_points` (pcdet/datasets/processor/data_processor.py:145-177) and of `break_up_pc` + the per-frame
count / view / permute of `PointNet2FSMSG.forward` (pcdet/models/backbones_3d/pointnet2_backbone.py:193-222), with the
data movement done by one device kernel (`de6d_stage_points`, csrc/stage_points.cu).

The random draw of `sample_points` stays on the host with numpy, in the reference's call order, so a seeded run picks
the same points; only the index list travels, the gather itself happens on the device in the same pass that splits
xyz / features into the layouts the SA layers read.
"""
import numpy as np
import torch

from ._lib import call
from .compat._common import stream_ptr


def sample_points_choice(points, num_points, rng=None):
    """The index list `choice` of data_processor.py:145-177 (`points[choice]` is what the reference stores): keep every
    point beyond 40 m and fill with random near points, or draw uniformly when the far points alone exceed the budget;
    pad by re-drawing when the frame is short.  `rng` defaults to the global `np.random` state like the reference, and
    the draws are issued in the same order, so `np.random.seed(s)` reproduces the reference's selection."""
    rng = np.random if rng is None else rng
    n = len(points)
    if num_points == -1:
        return np.arange(n, dtype=np.int32)
    if num_points < n:
        depth = np.linalg.norm(points[:, 0:3], axis=1)
        near = depth < 40.0
        far_idx = np.where(near == 0)[0]
        near_idx = np.where(near == 1)[0]
        if num_points > len(far_idx):
            pick = rng.choice(near_idx, num_points - len(far_idx), replace=False)
            choice = np.concatenate((pick, far_idx), axis=0) if len(far_idx) > 0 else pick
        else:
            choice = rng.choice(np.arange(0, n, dtype=np.int32), num_points, replace=False)
        rng.shuffle(choice)
        return choice
    choice = np.arange(0, n, dtype=np.int32)
    if num_points > n:
        short = num_points - n
        extra = rng.choice(choice, short, replace=short > n)
        choice = np.concatenate((choice, extra), axis=0)
    rng.shuffle(choice)
    return choice


def _stage(src, B, N, lead, choice, want_batch_idx, check):
    assert src.is_cuda and src.dtype == torch.float32 and src.dim() == 2 and src.is_contiguous()
    C = src.shape[1] - 3 - lead
    assert C >= 0, "rows need at least %d columns" % (3 + lead)
    dev = src.device
    xyz = torch.empty((B, N, 3), dtype=torch.float32, device=dev)
    feats = torch.empty((B, C, N), dtype=torch.float32, device=dev) if C > 0 else None
    bidx = torch.empty((B, N), dtype=torch.float32, device=dev) if want_batch_idx else None
    status = torch.empty(2, dtype=torch.int32, device=dev) if check else None
    if choice is not None:
        assert choice.is_cuda and choice.dtype == torch.int32 and choice.is_contiguous() and choice.numel() == B * N
    call("de6d_stage_points", B, N, C, lead, src.shape[0], src.data_ptr(),
         choice.data_ptr() if choice is not None else None, xyz.data_ptr(),
         feats.data_ptr() if feats is not None else None, bidx.data_ptr() if bidx is not None else None,
         status.data_ptr() if status is not None else None, stream_ptr())
    if check:
        bad_frame, bad_row = status.tolist()          # the one host sync (the reference does B + 2 of them)
        assert bad_row == 0, "%d sample indices outside the source" % bad_row
        assert bad_frame == 0, "frames do not hold the same number of points (%d misplaced rows)" % bad_frame
    return bidx, xyz, feats


@torch.no_grad()
def break_up_pc(points, batch_size, check=True):
    """points (B*N, 4+C) CUDA f32 rows [batch_idx, x, y, z, features...] (the collated `batch_dict['points']`) ->
    (batch_idx (B,N) f32, xyz (B,N,3), features (B,C,N) or None): what pointnet2_backbone.py:193-222 computes with
    slicing copies, a Python loop of B `.sum()` syncs, `.view` and `.permute(0,2,1).contiguous()`.  With `check` the
    reference's `assert xyz_batch_cnt.min() == xyz_batch_cnt.max()` is evaluated from a device counter (one sync);
    `check=False` is sync-free."""
    assert points.shape[0] % batch_size == 0, "frames do not hold the same number of points"
    return _stage(points, batch_size, points.shape[0] // batch_size, 1, None, True, check)


@torch.no_grad()
def stage_frames(rows, choice, check=True):
    """rows (total, 3+C) CUDA f32: the raw frames concatenated (no batch column); choice (B, N) CUDA i32: per frame the
    `sample_points_choice` list plus that frame's row offset -> (xyz (B,N,3), features (B,C,N) or None).  Replaces
    host `points[choice]`, collate, H2D of the padded array and `break_up_pc` with one H2D of the raw rows + indices
    and one kernel."""
    B, N = choice.shape
    _, xyz, feats = _stage(rows, B, N, 0, choice, False, check)
    return xyz, feats


def collate_choice(frames, num_points, rng=None):
    """Host helper: list of (n_i, 3+C) numpy frames -> (rows (sum n_i, 3+C) f32, choice (B, num_points) i32 global row
    indices) drawn frame by frame in order, as the reference's dataloader would with one worker."""
    offs = np.cumsum([0] + [len(f) for f in frames])
    choice = np.stack([sample_points_choice(f, num_points, rng).astype(np.int64) + offs[k]
                       for k, f in enumerate(frames)]).astype(np.int32)
    rows = np.ascontiguousarray(np.concatenate(frames, axis=0), dtype=np.float32)
    return rows, choice
