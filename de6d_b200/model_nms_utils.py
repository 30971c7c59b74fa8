"""Mirror of pcdet/models/model_utils/model_nms_utils.py (the caller of nms_gpu in Detector3DTemplate.post_processing,
detector3d_template.py:199-282) plus its batched, host-synchronisation-free form.

`class_agnostic_nms` keeps the reference signature and return values for one frame.  `class_agnostic_nms_batched`
runs score threshold -> top-k -> sort -> rotated NMS -> index mapping for a whole batch with one sort, one gather and
ONE kernel launch, returning padded device tensors; per frame it selects exactly what the reference's Python loop
selects (scores within a frame must be distinct for the orders to be comparable, as in the reference itself, whose
result under ties depends on torch.topk / torch.sort internals).
"""
import torch

from . import iou3d_nms_utils


def _cfg(nms_config, key, default=None):
    if isinstance(nms_config, dict):
        return nms_config.get(key, default)
    return getattr(nms_config, key, default)


def class_agnostic_nms(box_scores, box_preds, nms_config, score_thresh=None):
    """(reference :6-25) box_scores (N), box_preds (N, 7+C) -> (selected indices into the inputs, their scores)."""
    src_box_scores = box_scores
    if score_thresh is not None:
        scores_mask = (box_scores >= score_thresh)
        box_scores = box_scores[scores_mask]
        box_preds = box_preds[scores_mask]
    selected = []
    if box_scores.shape[0] > 0:
        box_scores_nms, indices = torch.topk(box_scores, k=min(_cfg(nms_config, "NMS_PRE_MAXSIZE"), box_scores.shape[0]))
        boxes_for_nms = box_preds[indices]
        nms_type = _cfg(nms_config, "NMS_TYPE", "nms_gpu")
        # NMS_TYPE "nms_gpu_9dof" (not in the reference) keeps rz, ry, rx: the reference slices [:, 0:7] (:18) and so
        # suppresses with the yaw-only BEV IoU even though Det6D's boxes carry pitch and roll
        width = 9 if nms_type == "nms_gpu_9dof" else 7
        keep_idx, _ = getattr(iou3d_nms_utils, nms_type)(
            boxes_for_nms[:, 0:width], box_scores_nms, _cfg(nms_config, "NMS_THRESH"))
        selected = indices[keep_idx[:_cfg(nms_config, "NMS_POST_MAXSIZE")]]
    if score_thresh is not None:
        original_idxs = scores_mask.nonzero().view(-1)
        selected = original_idxs[selected]
    return selected, src_box_scores[selected]


class BatchedClassAgnosticNMS:
    """op = BatchedClassAgnosticNMS(frames, n, nms_config); selected, scores, num = op(box_scores, box_preds, score_thresh)

    box_scores (F, n), box_preds (F, n, 7+C) -> selected (F, K) int64 indices into each frame's boxes (descending
    score, first num[f] valid, the rest -1), scores (F, K), num (F) int32, with K = min(NMS_POST_MAXSIZE, n, PRE).
    Static buffers: reusable and CUDA-graph capturable."""

    def __init__(self, frames, n, nms_config, device="cuda"):
        self.frames, self.n = frames, n
        self.pre = min(int(_cfg(nms_config, "NMS_PRE_MAXSIZE")), n)
        self.post = min(int(_cfg(nms_config, "NMS_POST_MAXSIZE")), self.pre)
        self.thresh = float(_cfg(nms_config, "NMS_THRESH"))
        self.normal = _cfg(nms_config, "NMS_TYPE", "nms_gpu") == "nms_normal_gpu"
        self.width = 9 if _cfg(nms_config, "NMS_TYPE", "nms_gpu") == "nms_gpu_9dof" else 7
        self.nms = iou3d_nms_utils.BatchedNMS(frames, self.pre, device=device, box_dim=self.width)

    @torch.no_grad()
    def __call__(self, box_scores, box_preds, score_thresh=None):
        F, n = self.frames, self.n
        assert box_scores.shape == (F, n) and box_preds.shape[:2] == (F, n)
        scores = box_scores
        if score_thresh is not None:
            passed = scores >= score_thresh
            scores = torch.where(passed, scores, torch.full_like(scores, float("-inf")))
            nvalid = passed.sum(1).clamp(max=self.pre).to(torch.int32)
        else:
            nvalid = torch.full((F,), self.pre, dtype=torch.int32, device=scores.device)
        top_scores, order = torch.sort(scores, dim=1, descending=True)
        order, top_scores = order[:, :self.pre], top_scores[:, :self.pre]
        boxes = torch.gather(box_preds[..., 0:self.width], 1, order.unsqueeze(-1).expand(-1, -1, self.width)).contiguous()
        keep_pos, num = self.nms(boxes, None, self.thresh, nvalid=nvalid.contiguous(), normal=self.normal, presorted=True)
        keep_pos = keep_pos[:, :self.post]
        num = num.clamp(max=self.post)
        valid = torch.arange(self.post, device=scores.device).unsqueeze(0) < num.unsqueeze(1)
        selected = torch.where(valid, torch.gather(order, 1, keep_pos), torch.full_like(keep_pos, -1))
        sel_scores = torch.where(valid, torch.gather(box_scores, 1, selected.clamp(min=0)), torch.zeros_like(top_scores[:, :self.post]))
        return selected, sel_scores, num


def class_agnostic_nms_batched(box_scores, box_preds, nms_config, score_thresh=None):
    """Functional form for one-off calls."""
    op = BatchedClassAgnosticNMS(box_scores.shape[0], box_scores.shape[1], nms_config, device=box_scores.device)
    return op(box_scores, box_preds, score_thresh)
