"""Multi-GPU plumbing: frames are independent units, so a batch is sharded contiguously across ranks with no
collective on the data path; one all_gather of fixed-shape padded detections at the end replaces the
reference's pickle-files-on-a-shared-filesystem merge (pcdet/utils/common_utils.py:212-233).
One process per GPU, torch.distributed (NCCL on GPUs, gloo in CPU tests)."""
import os
from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of `total` frames for `rank`; the first total % world ranks get one extra."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init_from_env(backend: str = None):
    """Reads RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun).  Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def gather_detections(keep: torch.Tensor, num: torch.Tensor, frames_per_rank: int = None):
    """All-gather per-rank padded detections.  keep (F_local, K) int64, num (F_local,) int32; every rank must
    pass the same F_local (pad the last shard).  Returns (keep_all (world*F_local, K), num_all (world*F_local,))
    on every rank; with world == 1 the inputs are returned unchanged."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return keep, num
    world = dist.get_world_size()
    keep_all = torch.empty((world * keep.shape[0],) + tuple(keep.shape[1:]), dtype=keep.dtype, device=keep.device)
    num_all = torch.empty((world * num.shape[0],), dtype=num.dtype, device=num.device)
    dist.all_gather_into_tensor(keep_all, keep.contiguous())
    dist.all_gather_into_tensor(num_all, num.contiguous())
    return keep_all, num_all


def gather_detection_boxes(boxes: torch.Tensor, scores: torch.Tensor, keep: torch.Tensor, num: torch.Tensor):
    """The end-of-batch collective SURVEY.md 8(e) describes: every rank contributes fixed-shape padded detections
    (F_local, K, D) boxes + (F_local, K) scores + (F_local,) counts and receives the whole batch's (replaces the reference's
    pickle files on a shared filesystem between two barriers, pcdet/utils/common_utils.py:212-233).
    boxes (F_local, n, D), scores (F_local, n): the frame's proposals; keep (F_local, K) int64 kept indices padded beyond
    num[f]; num (F_local,) int32.  Returns (boxes_all (world*F_local, K, D), scores_all (world*F_local, K), num_all); rows
    beyond num are zero.  Three all_gather_into_tensor calls, no host synchronisation; with world == 1 no collective."""
    F, K = keep.shape
    valid = torch.arange(K, device=keep.device).unsqueeze(0) < num.unsqueeze(1)
    kc = keep.clamp(min=0)
    det_boxes = torch.gather(boxes, 1, kc.unsqueeze(-1).expand(-1, -1, boxes.size(2))) * valid.unsqueeze(-1)
    det_scores = torch.gather(scores, 1, kc) * valid
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return det_boxes, det_scores, num
    world = dist.get_world_size()
    boxes_all = torch.empty((world * F, K, boxes.size(2)), dtype=boxes.dtype, device=boxes.device)
    scores_all = torch.empty((world * F, K), dtype=scores.dtype, device=scores.device)
    num_all = torch.empty((world * F,), dtype=num.dtype, device=num.device)
    dist.all_gather_into_tensor(boxes_all, det_boxes.contiguous())
    dist.all_gather_into_tensor(scores_all, det_scores.contiguous())
    dist.all_gather_into_tensor(num_all, num.contiguous())
    return boxes_all, scores_all, num_all


def max_over_ranks(value: float, device=None) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
