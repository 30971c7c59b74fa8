"""Host-side mirror of pcdet/ops/iou3d_nms/iou3d_nms_utils.py: same names and return values, B200 kernels
underneath.  Additions (not in the reference): `nms_gpu_batched` -- the sync-free batched form used by the
op chain -- and the fused single-kernel `boxes_iou3d_gpu`.
"""
import numpy as np
import torch

from ._lib import call, load
from .compat import iou3d_nms_cuda as _ext


def _to_torch(x):
    if isinstance(x, np.ndarray):
        return torch.from_numpy(x).float(), True
    return x, False


def boxes_bev_iou_cpu(boxes_a, boxes_b):
    """(reference :12-28) CPU tensors / numpy in, same kind out; evaluated on the calling host thread like the reference
    (safe inside forked DataLoader workers), bit-identical to it.  `boxes_iou_bev` is the device form."""
    boxes_a, is_numpy = _to_torch(boxes_a)
    boxes_b, _ = _to_torch(boxes_b)
    assert not (boxes_a.is_cuda or boxes_b.is_cuda), 'Only support CPU tensors'
    assert boxes_a.shape[1] == 7 and boxes_b.shape[1] == 7
    ans_iou = boxes_a.new_zeros(torch.Size((boxes_a.shape[0], boxes_b.shape[0])))
    _ext.boxes_iou_bev_cpu(boxes_a.float().contiguous(), boxes_b.float().contiguous(), ans_iou)
    return ans_iou.numpy() if is_numpy else ans_iou


def boxes_iou_bev(boxes_a, boxes_b):
    """(reference :31-45) rotated BEV IoU, (N, 7) x (M, 7) -> (N, M)."""
    assert boxes_a.shape[1] == boxes_b.shape[1] == 7
    ans_iou = torch.zeros((boxes_a.shape[0], boxes_b.shape[0]), dtype=torch.float32, device=boxes_a.device)
    _ext.boxes_iou_bev_gpu(boxes_a.contiguous(), boxes_b.contiguous(), ans_iou)
    return ans_iou


def boxes_iou3d_gpu(boxes_a, boxes_b):
    """(reference :48-81) 3-D IoU = BEV overlap x height overlap / union volume, one fused kernel that rounds
    after every step exactly where the reference's separate torch kernels do."""
    assert boxes_a.shape[1] == boxes_b.shape[1] == 7
    ans = torch.zeros((boxes_a.shape[0], boxes_b.shape[0]), dtype=torch.float32, device=boxes_a.device)
    _ext.boxes_iou3d_gpu(boxes_a.contiguous(), boxes_b.contiguous(), ans)
    return ans


def boxes_iou3d_9dof_gpu(boxes_a, boxes_b):
    """Full-pose IoU (not in the reference, whose boxes_iou3d_gpu ignores pitch and roll): boxes (N, 9) / (M, 9)
    [x, y, z, dx, dy, dz, rz, ry, rx], rotation as box_utils.boxes3d_to_corners_3d (pcdet/utils/box_utils.py:59-72) -> (N, M).
    For ry = rx = 0 it is the exact-geometry value of what boxes_iou3d_gpu approximates (the reference's BEV clipping
    pads its corner tests by 1e-2 m, so the two agree to ~1e-3)."""
    assert boxes_a.shape[1] == boxes_b.shape[1] == 9
    ans = torch.zeros((boxes_a.shape[0], boxes_b.shape[0]), dtype=torch.float32, device=boxes_a.device)
    _ext.boxes_iou3d9_gpu(boxes_a.contiguous(), boxes_b.contiguous(), ans)
    return ans


def nms_gpu_9dof(boxes, scores, thresh, pre_maxsize=None, **kwargs):
    """nms_gpu (reference :84-99) on (N, 9) boxes with the full-pose 3-D IoU: returns (kept indices, None)."""
    assert boxes.shape[1] == 9
    order = scores.sort(0, descending=True)[1]
    if pre_maxsize is not None:
        order = order[:pre_maxsize]
    boxes = boxes[order].contiguous()
    keep = torch.empty(boxes.size(0), dtype=torch.int64)
    num_out = _ext.nms9_gpu(boxes, keep, thresh)
    return order[keep[:num_out].to(boxes.device)].contiguous(), None


def nms_gpu(boxes, scores, thresh, pre_maxsize=None, **kwargs):
    """(reference :84-99) returns (indices into `boxes` of the kept boxes in descending score order, None)."""
    assert boxes.shape[1] == 7
    order = scores.sort(0, descending=True)[1]
    if pre_maxsize is not None:
        order = order[:pre_maxsize]
    boxes = boxes[order].contiguous()
    keep = torch.empty(boxes.size(0), dtype=torch.int64)
    num_out = _ext.nms_gpu(boxes, keep, thresh)
    return order[keep[:num_out].to(boxes.device)].contiguous(), None


def nms_normal_gpu(boxes, scores, thresh, **kwargs):
    """(reference :102-116) axis-aligned variant."""
    assert boxes.shape[1] == 7
    order = scores.sort(0, descending=True)[1]
    boxes = boxes[order].contiguous()
    keep = torch.empty(boxes.size(0), dtype=torch.int64)
    num_out = _ext.nms_normal_gpu(boxes, keep, thresh)
    return order[keep[:num_out].to(boxes.device)].contiguous(), None


class BatchedNMS:
    """Sync-free NMS over a whole batch of frames: one sort, one gather, one kernel launch; outputs stay on the
    device as padded tensors.  Equivalent to calling `nms_gpu` frame by frame (detector3d_template.py:199-282
    does that in a Python loop with a malloc + D2H + host sweep + H2D per frame).

        nms = BatchedNMS(frames, n)                      # owns the workspace, reusable / graph-capturable
        keep, num = nms(boxes, scores, thresh)           # keep (F, n) int64 indices into boxes[f], num (F) int32
    """

    def __init__(self, frames, n, device="cuda", box_dim=7):
        self.frames, self.n, self.box_dim = frames, n, box_dim      # box_dim 9: full-pose IoU (nms_gpu_9dof)
        lib = load()
        self.ws_bytes = int(lib.de6d_nms_workspace_bytes(frames, n))
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=device)
        call("de6d_nms_workspace_init", frames, self.ws.data_ptr(), torch.cuda.current_stream().cuda_stream)
        self.keep_pos = torch.zeros((frames, n), dtype=torch.int64, device=device)
        self.num = torch.zeros(frames, dtype=torch.int32, device=device)

    def __call__(self, boxes, scores, thresh, nvalid=None, normal=False, presorted=False):
        F, n = self.frames, self.n
        D = self.box_dim
        assert boxes.shape == (F, n, D) and boxes.dtype == torch.float32
        assert not (normal and D == 9), "the axis-aligned variant takes 7-value boxes"
        if presorted:
            order, sorted_boxes = None, boxes.contiguous()
        else:
            order = scores.sort(1, descending=True)[1]
            sorted_boxes = torch.gather(boxes, 1, order.unsqueeze(-1).expand(-1, -1, D)).contiguous()
        if nvalid is not None:
            if not (nvalid.is_cuda and nvalid.dtype == torch.int32 and nvalid.is_contiguous() and nvalid.numel() == F
                    and nvalid.device == boxes.device):
                raise ValueError("nvalid must be a contiguous int32 tensor of %d elements on %s" % (F, boxes.device))
        if not boxes.is_cuda or boxes.device != self.ws.device:
            raise ValueError("boxes must live on %s" % self.ws.device)
        call("de6d_nms_batched", F, n, sorted_boxes.data_ptr(), None if nvalid is None else nvalid.data_ptr(),
             float(thresh), 2 if D == 9 else int(bool(normal)), self.keep_pos.data_ptr(), self.num.data_ptr(), self.ws.data_ptr(),
             self.ws_bytes, torch.cuda.current_stream().cuda_stream)
        if order is None:
            return self.keep_pos, self.num
        # positions -> original indices; entries beyond num[f] are padding: the kernel writes position 0 there, so the
        # padded tail of the returned row repeats the frame's top-scoring box index
        return torch.gather(order, 1, self.keep_pos), self.num


def nms_gpu_batched(boxes, scores, thresh, nvalid=None, normal=False):
    """Functional form of BatchedNMS for one-off calls: boxes (F, n, 7), scores (F, n)."""
    op = BatchedNMS(boxes.shape[0], boxes.shape[1], device=boxes.device, box_dim=boxes.shape[2])
    keep, num = op(boxes, scores, thresh, nvalid=nvalid, normal=normal)
    return keep.clone(), num.clone()
