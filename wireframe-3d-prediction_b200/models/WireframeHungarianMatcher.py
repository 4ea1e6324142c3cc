"""WireframeHungarianMatcher -- drop-in for the reference's models/WireframeHungarianMatcher.py.
Cost blocks (L1 cdist + |existence difference|) and the assignment are computed on the device; only the
final index arrays come back (the reference moves the whole (B*V, sum T) cost matrix to the CPU)."""
import torch
from torch import nn

from wf_b200 import ops
from wf_b200._lib import LSAP_INFEASIBLE, LSAP_INVALID, call


def _raise_status(status):
    st = status.tolist() if torch.is_tensor(status) else list(status)
    if any(s == LSAP_INFEASIBLE for s in st):
        raise ValueError("cost matrix is infeasible")
    if any(s == LSAP_INVALID for s in st):
        raise ValueError("matrix contains invalid numeric entries")


def _to_pairs(col, sizes, nq):
    """col_of_row (B, nq) host int array -> scipy-style (row_idx, col_idx) int64 tensors (rows ascending).  One vectorised
    pass over the whole batch: the matched (row, column) pairs of all samples come out of a single nonzero, split per sample."""
    import numpy as np
    col = np.asarray(col)[:, :nq]
    b_idx, r_idx = np.nonzero(col >= 0)                       # row-major: samples in order, rows ascending inside each
    c_idx = col[b_idx, r_idx].astype(np.int64)
    cuts = np.cumsum(np.bincount(b_idx, minlength=col.shape[0]))[:-1]
    rows = np.split(r_idx.astype(np.int64), cuts)
    cols = np.split(c_idx, cuts)
    return [(torch.from_numpy(np.ascontiguousarray(r)), torch.from_numpy(np.ascontiguousarray(c))) for r, c in zip(rows, cols)]


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
        # ONE device->host copy (assignments and solver status together) and no per-sample tensor ops on the way back
        host = torch.cat([col, status.view(bs, 1)], dim=1).cpu().numpy()
        _raise_status(host[:, -1])
        return _to_pairs(host[:, :-1], sizes, nq)


def build_wireframe_matcher(cost_vertex=1.0, cost_existence=1.0):
    return WireframeHungarianMatcher(cost_vertex=cost_vertex, cost_existence=cost_existence)
