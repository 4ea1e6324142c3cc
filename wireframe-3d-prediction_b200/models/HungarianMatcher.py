"""HungarianMatcher (DETR bbox matcher) -- drop-in for the reference's models/HungarianMatcher.py.
The helper functions keep their reference semantics (plain tensor expressions, usable on any device);
the matcher's cost blocks and assignments run on the device through libwf_b200."""
import torch
from torch import nn

from wf_b200 import ops
from wf_b200._lib import call
from models.WireframeHungarianMatcher import _raise_status, _to_pairs


def box_cxcywh_to_xyxy(x):
    x_c, y_c, w, h = x.unbind(-1)
    return torch.stack([(x_c - 0.5 * w), (y_c - 0.5 * h), (x_c + 0.5 * w), (y_c + 0.5 * h)], dim=-1)


def box_iou(boxes1, boxes2):
    area1 = (boxes1[:, 2] - boxes1[:, 0]) * (boxes1[:, 3] - boxes1[:, 1])
    area2 = (boxes2[:, 2] - boxes2[:, 0]) * (boxes2[:, 3] - boxes2[:, 1])
    lt = torch.max(boxes1[:, None, :2], boxes2[:, :2])
    rb = torch.min(boxes1[:, None, 2:], boxes2[:, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[:, :, 0] * wh[:, :, 1]
    union = area1[:, None] + area2 - inter
    return inter / union, union


def generalized_box_iou(boxes1, boxes2):
    assert (boxes1[:, 2:] >= boxes1[:, :2]).all()
    assert (boxes2[:, 2:] >= boxes2[:, :2]).all()
    iou, union = box_iou(boxes1, boxes2)
    lt = torch.min(boxes1[:, None, :2], boxes2[:, :2])
    rb = torch.max(boxes1[:, None, 2:], boxes2[:, 2:])
    wh = (rb - lt).clamp(min=0)
    area = wh[:, :, 0] * wh[:, :, 1]
    return iou - (area - union) / area


class HungarianMatcher(nn.Module):
    def __init__(self, cost_class: float = 1, cost_bbox: float = 1, cost_giou: float = 1):
        super().__init__()
        self.cost_class = cost_class
        self.cost_bbox = cost_bbox
        self.cost_giou = cost_giou
        assert cost_class != 0 or cost_bbox != 0 or cost_giou != 0, "all costs cant be 0"

    @torch.no_grad()
    def forward(self, outputs, targets):
        logits = ops._f32c(outputs["pred_logits"])
        boxes = ops._f32c(outputs["pred_boxes"])
        ops._need_cuda(logits)
        bs, nq, ncls = logits.shape
        dev = logits.device
        sizes = [len(v["boxes"]) for v in targets]
        labels = torch.cat([v["labels"].reshape(-1) for v in targets]).to(dev, torch.int64).contiguous()
        tb = torch.cat([v["boxes"].reshape(-1, 4) for v in targets]).to(dev, torch.float32).contiguous()
        off = [0]
        for t in sizes:
            off.append(off[-1] + t)
        ld = max(max(sizes), 1)
        toff = torch.tensor(off, dtype=torch.int32, device=dev)
        cost = torch.zeros(bs, nq, ld, device=dev, dtype=torch.float32)
        if tb.numel() > 0:
            call("wf_detr_matcher_cost", ops._p(logits), ops._p(boxes), ops._p(labels), ops._p(tb), ops._p(toff), bs, nq,
                 ncls, float(self.cost_class), float(self.cost_bbox), float(self.cost_giou), ops._p(cost), ld, ops._s())
            ops._count()
        nr = torch.full((bs,), nq, dtype=torch.int32, device=dev)
        nc = torch.tensor(sizes, dtype=torch.int32, device=dev)
        col, status = ops.lsap_batched(cost, nr, nc)
        _raise_status(status)
        return _to_pairs(col.cpu(), sizes, nq)


def build_matcher(args):
    return HungarianMatcher(cost_class=args.set_cost_class, cost_bbox=args.set_cost_bbox, cost_giou=args.set_cost_giou)
