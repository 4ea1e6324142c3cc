"""WireframeHungarianMatcher -- drop-in for the reference's models/WireframeHungarianMatcher.py.
Cost blocks (L1 cdist + |existence difference|) and the assignment are computed on the device; only the
final index arrays come back (the reference moves the whole (B*V, sum T) cost matrix to the CPU)."""
import torch
from torch import nn

from wf_b200 import ops
from wf_b200._lib import LSAP_INFEASIBLE, LSAP_INVALID, call


def _raise_status(status):
    st = status.tolist()
    if any(s == LSAP_INFEASIBLE for s in st):
        raise ValueError("cost matrix is infeasible")
    if any(s == LSAP_INVALID for s in st):
        raise ValueError("matrix contains invalid numeric entries")


def _to_pairs(col, sizes, nq):
    """col_of_row (B, nq) on host -> scipy-style (row_idx, col_idx) int64 tensors (rows ascending)."""
    out = []
    for b, t in enumerate(sizes):
        c = col[b, :nq]
        rows = torch.nonzero(c >= 0).flatten() if t < nq else torch.arange(nq)
        out.append((rows.to(torch.int64), c[rows].to(torch.int64)))
    return out


class WireframeHungarianMatcher(nn.Module):
    def __init__(self, cost_vertex: float = 1.0, cost_existence: float = 1.0):
        super().__init__()
        self.cost_vertex = cost_vertex
        self.cost_existence = cost_existence
        assert cost_vertex != 0 or cost_existence != 0, "all costs cant be 0"

    @torch.no_grad()
    def forward(self, outputs, targets):
        pv = ops._f32c(outputs["vertices"])
        pe = ops._f32c(outputs["existence_probabilities"])
        ops._need_cuda(pv)
        bs, nq = pv.shape[:2]
        dev = pv.device
        sizes = [len(v["vertices"]) for v in targets]
        tv = torch.cat([v["vertices"].reshape(-1, 3) for v in targets]).to(dev, torch.float32).contiguous()
        te = torch.cat([v["existence"].reshape(-1) for v in targets]).to(dev, torch.float32).contiguous()
        off = [0]
        for t in sizes:
            off.append(off[-1] + t)
        ld = max(max(sizes), 1)
        toff = torch.tensor(off, dtype=torch.int32, device=dev)
        cost = torch.zeros(bs, nq, ld, device=dev, dtype=torch.float32)
        if tv.numel() > 0:
            call("wf_wireframe_matcher_cost", ops._p(pv), ops._p(pe), ops._p(tv), ops._p(te), ops._p(toff), bs, nq,
                 float(self.cost_vertex), float(self.cost_existence), ops._p(cost), ld, ops._s())
            ops._count()
        nr = torch.full((bs,), nq, dtype=torch.int32, device=dev)
        nc = torch.tensor(sizes, dtype=torch.int32, device=dev)
        col, status = ops.lsap_batched(cost, nr, nc)
        _raise_status(status)
        return _to_pairs(col.cpu(), sizes, nq)


def build_wireframe_matcher(cost_vertex=1.0, cost_existence=1.0):
    return WireframeHungarianMatcher(cost_vertex=cost_vertex, cost_existence=cost_existence)
