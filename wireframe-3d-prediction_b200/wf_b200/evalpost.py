"""Ragged-batch fp64 primitives of the evaluation post-processing (include/wf_b200.h, "Evaluation
post-processing"): Hausdorff distance between sampled segments, Euclidean cdist and the assignment solver
for matrices of any shape.  Inputs and outputs are host numpy arrays, as `eval/ap_calculator.py` receives
and produces them; every list is packed into ONE device buffer with prefix offsets and costs ONE launch,
whatever the number of samples.  No CPU path: a missing CUDA device raises."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from . import ops

LSAP_ERRORS = {_lib.LSAP_INFEASIBLE: "cost matrix is infeasible",
               _lib.LSAP_INVALID: "matrix contains invalid numeric entries"}


def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.WfError("wf_b200.evalpost runs on a CUDA device only; there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def _offsets(counts: Sequence[int]) -> np.ndarray:
    off = np.zeros(len(counts) + 1, dtype=np.int64)
    np.cumsum(np.asarray(counts, dtype=np.int64), out=off[1:])
    return off


def _dev(a: np.ndarray, dev) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev, non_blocking=False)


def _pack_rows(parts: Sequence[np.ndarray], width: int) -> Tuple[np.ndarray, np.ndarray]:
    off = _offsets([len(p) for p in parts])
    flat = np.zeros((int(off[-1]), width), dtype=np.float64)
    for p, a in zip(parts, off[:-1]):
        if len(p):
            flat[a:a + len(p)] = np.asarray(p, dtype=np.float64).reshape(len(p), width)
    return flat, off


def _split_blocks(flat: np.ndarray, off: np.ndarray, shapes) -> List[np.ndarray]:
    return [flat[off[b]:off[b + 1]].reshape(shapes[b]) for b in range(len(shapes))]


def start_delta(lines: np.ndarray) -> np.ndarray:
    """(L,2,3) segments -> (L,2,3) float64 [start, end-start]; the difference is taken in the segments' own
    dtype before widening, exactly as numpy evaluates eval/ap_calculator.py:24-25 on float32 input."""
    lines = np.asarray(lines)
    out = np.empty(lines.shape, dtype=np.float64)
    out[:, 0, :] = lines[:, 0, :]
    out[:, 1, :] = lines[:, 1, :] - lines[:, 0, :]
    return out


def _hausdorff_device(p_lines, t_lines, samples, dev):
    """Pack, launch; returns (device matrix buffer or None, block offsets, shapes)."""
    B = len(p_lines)
    sd_p, sd_t = [], []
    for p, t in zip(p_lines, t_lines):
        p, t = np.asarray(p), np.asarray(t)
        common = np.result_type(p.dtype, t.dtype)
        sd_p.append(start_delta(p.astype(common, copy=False)))
        sd_t.append(start_delta(t.astype(common, copy=False)))
    pf, p_off = _pack_rows(sd_p, 6)
    tf, t_off = _pack_rows(sd_t, 6)
    shapes = [(len(p), len(t)) for p, t in zip(p_lines, t_lines)]
    o_off = _offsets([n * m for n, m in shapes])
    total = int(o_off[-1])
    if total == 0:
        return None, o_off, shapes
    w = np.linspace(0, 1, samples)
    d_p, d_t, d_w = _dev(pf, dev), _dev(tf, dev), _dev(w, dev)
    d_po, d_to, d_oo = _dev(p_off, dev), _dev(t_off, dev), _dev(o_off, dev)
    out = torch.empty(total, dtype=torch.float64, device=dev)
    _lib.call("wf_hausdorff_lines", ops._p(d_p), ops._p(d_po), ops._p(d_t), ops._p(d_to), ops._p(d_oo), B,
              max(n for n, _ in shapes), ops._p(d_w), int(samples), ops._p(out), ops._s())
    ops._count()
    return out, o_off, shapes


def hausdorff_lines_batched(p_lines: Sequence[np.ndarray], t_lines: Sequence[np.ndarray],
                            samples: int = 20) -> List[np.ndarray]:
    """Per sample b: eval/ap_calculator.py:8-36 on (N_b,2,3) x (M_b,2,3) -> (N_b,M_b) float64.  As in the
    reference both sets are first brought to one dtype (np.concatenate at :20); N_b == 0 gives np.array([])."""
    out, o_off, shapes = _hausdorff_device(p_lines, t_lines, samples, _device())
    if out is None:
        return [np.array([]) if n == 0 else np.zeros((n, m)) for n, m in shapes]
    host = out.cpu().numpy()
    return [np.array([]) if n == 0 else host[o_off[b]:o_off[b + 1]].reshape(n, m) for b, (n, m) in enumerate(shapes)]


def hausdorff_assign_batched(p_lines: Sequence[np.ndarray], t_lines: Sequence[np.ndarray], samples: int = 20):
    """eval/ap_calculator.py:159-163 for every sample at once: Hausdorff matrix -> linear_sum_assignment ->
    the matched distances.  The matrices (2016 x 90 doubles per sample at the model's maximum) stay on the device;
    only (row_ind, col_ind, matrix[row_ind, col_ind]) come back."""
    dev = _device()
    out, o_off, shapes = _hausdorff_device(p_lines, t_lines, samples, dev)
    if out is None:
        e = np.zeros(0, dtype=np.int64)
        return [(e, e, np.zeros(0)) for _ in shapes]
    return _solve_device(out, o_off, shapes, dev)


def _cdist_device(a_list, b_list, dev):
    """Pack, launch wf_cdist_f64; returns (device matrix buffer or None, block offsets, shapes)."""
    B = len(a_list)
    dim = 3
    for x in list(a_list) + list(b_list):
        x = np.asarray(x)
        if x.ndim == 2 and x.shape[0] > 0:
            dim = x.shape[1]
            break
    af, a_off = _pack_rows([np.asarray(a).reshape(-1, dim) for a in a_list], dim)
    bf, b_off = _pack_rows([np.asarray(b).reshape(-1, dim) for b in b_list], dim)
    shapes = [(int(a_off[i + 1] - a_off[i]), int(b_off[i + 1] - b_off[i])) for i in range(B)]
    o_off = _offsets([n * m for n, m in shapes])
    total = int(o_off[-1])
    if total == 0:
        return None, o_off, shapes
    d_a, d_b = _dev(af, dev), _dev(bf, dev)
    d_ao, d_bo, d_oo = _dev(a_off, dev), _dev(b_off, dev), _dev(o_off, dev)
    out = torch.empty(total, dtype=torch.float64, device=dev)
    _lib.call("wf_cdist_f64", ops._p(d_a), ops._p(d_ao), ops._p(d_b), ops._p(d_bo), ops._p(d_oo), B,
              max(n * m for n, m in shapes), dim, ops._p(out), ops._s())
    ops._count()
    return out, o_off, shapes


def cdist_batched(a_list: Sequence[np.ndarray], b_list: Sequence[np.ndarray]) -> List[np.ndarray]:
    """scipy.spatial.distance.cdist(a, b) (euclidean, float64) for every pair of the two lists."""
    out, o_off, shapes = _cdist_device(a_list, b_list, _device())
    if out is None:
        return [np.zeros(s) for s in shapes]
    return _split_blocks(out.cpu().numpy(), o_off, shapes)


def cdist_assign_batched(a_list: Sequence[np.ndarray], b_list: Sequence[np.ndarray]):
    """models/utils.py:45-49 for every pair of the two lists: fp64 Euclidean cdist -> linear_sum_assignment, both on the
    device (the matrices never leave it).  Returns [(row_ind, col_ind, matched distances)]; raises scipy's ValueError
    texts for NaN / infeasible matrices."""
    dev = _device()
    out, o_off, shapes = _cdist_device(a_list, b_list, dev)
    if out is None:
        e = np.zeros(0, dtype=np.int64)
        return [(e, e, np.zeros(0)) for _ in shapes]
    return _solve_device(out, o_off, shapes, dev)


def _solve_device(d_cost: torch.Tensor, c_off: np.ndarray, shapes, dev):
    """wf_lsap_f64 on matrices already resident on the device -> [(rows, cols, matched costs)]."""
    B = len(shapes)
    nr = np.array([s[0] for s in shapes], dtype=np.int32)
    nc = np.array([s[1] for s in shapes], dtype=np.int32)
    r_off = _offsets(nr)
    work = torch.empty_like(d_cost)
    d_nr, d_nc, d_co, d_ro = _dev(nr, dev), _dev(nc, dev), _dev(c_off, dev), _dev(r_off, dev)
    n_rows = max(int(r_off[-1]), 1)
    col = torch.empty(n_rows, dtype=torch.int32, device=dev)
    matched = torch.empty(n_rows, dtype=torch.float64, device=dev)
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    _lib.call("wf_lsap_f64", ops._p(d_cost), ops._p(d_co), ops._p(d_nr), ops._p(d_nc), ops._p(d_ro), B,
              int(nr.max()), int(nc.max()), ops._p(work), ops._p(col), ops._p(matched), ops._p(status), ops._s())
    ops._count()
    for code in status.cpu().numpy():
        if code != _lib.LSAP_OK:
            raise ValueError(LSAP_ERRORS.get(int(code), "linear_sum_assignment failed"))
    col_h, m_h = col.cpu().numpy(), matched.cpu().numpy()
    out = []
    for b in range(B):
        c = col_h[r_off[b]:r_off[b + 1]].astype(np.int64)
        rows = np.flatnonzero(c >= 0)
        out.append((rows, c[rows], m_h[r_off[b]:r_off[b + 1]][rows]))
    return out


def lsap_batched_f64(mats: Sequence[np.ndarray]) -> List[Tuple[np.ndarray, np.ndarray]]:
    """scipy.optimize.linear_sum_assignment for every matrix of the list (float64, any shapes), one launch.
    Raises scipy's ValueError texts for infeasible / invalid matrices."""
    dev = _device()
    mats = [np.asarray(m, dtype=np.float64) for m in mats]
    for m in mats:
        if m.ndim != 2:
            raise ValueError("expected a matrix (2-D array), got a %r array" % (m.shape,))
    shapes = [m.shape for m in mats]
    c_off = _offsets([a * b for a, b in shapes])
    if int(c_off[-1]) == 0:
        e = np.zeros(0, dtype=np.int64)
        return [(e, e) for _ in mats]
    d_c = _dev(np.concatenate([m.reshape(-1) for m in mats]), dev)
    return [(r, c) for r, c, _ in _solve_device(d_c, c_off, shapes, dev)]


def segment_endpoints(vertices: np.ndarray, edges: np.ndarray) -> np.ndarray:
    """evaluate.py:87-98: (E,2,3) endpoints of each edge, the endpoint with the larger z first (on equal z the
    edge's second vertex comes first: stable argsort, then flip)."""
    if len(edges) == 0:
        return np.empty((0, 2, 3))
    ends = np.stack((vertices[edges[:, 0]], vertices[edges[:, 1]]), axis=1)
    first_is_higher = ends[:, 0, 2] > ends[:, 1, 2]
    return np.where(first_is_higher[:, None, None], ends, ends[:, ::-1, :])


def make_ap_batch(predictions: dict, wf_vertices: Sequence, wf_edges: Sequence, threshold: float = 0.5) -> dict:
    """evaluate.py:74-107 for a whole batch: the dictionary `APCalculator.compute_metrics` consumes, built from the
    model's output dict and the loader's ground truth with ONE device->host copy of vertices and edge probabilities
    (the reference copies per sample).  Each sample's probability row is cut to its own number of candidate edges
    (the reference's boolean index at :80-81 needs equal lengths and fails on ragged batches)."""
    verts = predictions['vertices'].detach().cpu().numpy()
    probs = predictions['edge_probs'].detach().cpu().numpy()
    out = {k: [] for k in ('predicted_vertices', 'predicted_edges', 'pred_edges_vertices', 'wf_vertices', 'wf_edges',
                           'wf_edges_vertices')}
    for b in range(len(wf_vertices)):
        cand = np.array(predictions['edge_indices'][b], dtype=np.int64).reshape(-1, 2)
        kept = cand[probs[b, :len(cand)] > threshold]
        gt_v = np.asarray(wf_vertices[b].numpy() if hasattr(wf_vertices[b], 'numpy') else wf_vertices[b])
        gt_e = np.asarray(wf_edges[b].numpy() if hasattr(wf_edges[b], 'numpy') else wf_edges[b]).astype(np.int64)
        out['predicted_vertices'].append(verts[b])
        out['predicted_edges'].append(kept)
        out['pred_edges_vertices'].append(segment_endpoints(verts[b], kept))
        out['wf_vertices'].append(gt_v)
        out['wf_edges'].append(gt_e)
        out['wf_edges_vertices'].append(segment_endpoints(gt_v, gt_e))
    return out
