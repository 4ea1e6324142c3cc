"""CPU oracle for the evaluation post-processing that follows the hot path.

TEST INFRASTRUCTURE ONLY -- imported by tests/, never by the product package.

Restates, sample by sample and in plain numpy, what the reference's `eval/ap_calculator.py` computes
(`hausdorff_distance_line` :8-36, `graph_edit_distance` :39-84, `computer_edges` :87-101,
`remove_corners` :104-108, `APCalculator.compute_metrics` :121-272, `output_accuracy` :274-302), and
the per-sample batch construction of `evaluate.py:74-104`.  The two third-party pieces are restated too:
scipy's `cdist(..., 'euclidean')` (double, d0*d0 + d1*d1 + d2*d2 accumulated in order, then sqrt) and
`linear_sum_assignment` (oracle/lsap_oracle.c).  Pinned by tests/golden/ap_calculator.npz, made by
running the unmodified reference class (tests/golden/make_golden_ap.py), and against the installed
scipy for cdist.
"""
from typing import Dict, Tuple

import numpy as np

from . import wireframe_oracle as wo

SAMPLES = 20  # eval/ap_calculator.py:8 default `sample_points`


def cdist_euclid(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """scipy.spatial.distance.cdist(a, b) for (n,d),(m,d) -> (n,m), double, summed in index order."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    acc = np.zeros((a.shape[0], b.shape[0]))
    for k in range(a.shape[1]):
        diff = a[:, k][:, None] - b[:, k][None, :]
        acc = acc + diff * diff
    return np.sqrt(acc)


def line_samples(lines: np.ndarray, samples: int = SAMPLES) -> np.ndarray:
    """(L,2,3) segments -> (L,samples,3) evenly spaced points, in the dtype arithmetic numpy uses at
    eval/ap_calculator.py:21-25: end-start in the segments' own dtype, times float64 linspace weights."""
    w = np.linspace(0, 1, samples).reshape(1, samples, 1)
    first = lines[:, 0, :][:, None, :]
    return first + w * (lines[:, 1, :][:, None, :] - first)


def hausdorff_lines(p_line: np.ndarray, t_line: np.ndarray, samples: int = SAMPLES) -> np.ndarray:
    """eval/ap_calculator.py:8-36.  Symmetric Hausdorff distance between sampled segments, (N,M)."""
    n, m = len(p_line), len(t_line)
    if n == 0:
        return np.array([])
    both = np.concatenate((p_line, t_line), axis=0)          # one dtype for both sets (:20)
    pts = line_samples(both, samples)
    pp, tt = pts[:n], pts[n:]
    out = np.empty((n, m))
    for i in range(n):                                       # row blocks keep the oracle's memory small
        d = cdist_euclid(pp[i], tt.reshape(-1, 3)).reshape(samples, m, samples)
        fwd = d.min(axis=2).max(axis=0)                      # h(pred -> target)
        bwd = d.min(axis=0).max(axis=1)                      # h(target -> pred)
        out[i] = np.maximum(fwd, bwd)
    return out


def unique_rows(a: np.ndarray) -> np.ndarray:
    return np.unique(a, axis=0)


def rows_not_in(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """eval/ap_calculator.py:104-108: sorted unique rows of `a` that are not rows of `b` (same dtype)."""
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    if a.dtype != b.dtype:
        raise TypeError("remove_corners needs equal dtypes (numpy structured view)")
    rec = [("", a.dtype)] * a.shape[1]
    return np.setdiff1d(a.view(rec), b.view(rec)).view(a.dtype).reshape(-1, a.shape[1])


def index_edges(edge_pts: np.ndarray, vertices: np.ndarray) -> np.ndarray:
    """eval/ap_calculator.py:87-101: first index of each endpoint in `vertices` (-1 if absent), pair sorted."""
    idx = np.full((len(edge_pts), 2), -1, dtype=np.int64)
    for e in range(len(edge_pts)):
        for s in range(2):
            hit = np.flatnonzero((vertices == edge_pts[e, s]).all(axis=1))
            if hit.size:
                idx[e, s] = hit[0]
    return np.sort(idx, axis=-1)


def seg_len(a, b):
    """np.linalg.norm(a - b) in the vertices' OWN dtype (float32 from the loader): the reference's edge
    lengths and their running sums (`wed_e`, `sum_distance`, :70-80) are float32 scalars."""
    d = a - b
    return np.sqrt(d.dot(d))


def edit_distance(pd_v: np.ndarray, pd_e: np.ndarray, gt_v: np.ndarray, gt_e: np.ndarray, wed_v: float) -> float:
    """eval/ap_calculator.py:39-84 (the caller always passes at least one vertex; the empty case is kept)."""
    wed_e = 0
    remaining = gt_e.copy()
    if len(pd_v) > 0:
        d = cdist_euclid(pd_v, gt_v)
        near = 0
        for row_min in d.min(axis=1):                        # python sum(), then one add (:49)
            near = near + row_min
        wed_v = wed_v + near
        snapped = np.array([gt_v[j] for j in d.argmin(axis=1)], dtype=pd_v.dtype)
        uniq = unique_rows(snapped)
        relabel = pd_e.copy()
        for new_id, p in enumerate(uniq):
            for old_id in np.flatnonzero((snapped == p).all(axis=1)):
                relabel[pd_e == old_id] = new_id
        relabel = np.unique(relabel, axis=0)
        for e in relabel:
            i0 = np.flatnonzero((gt_v == uniq[e[0]]).all(axis=1))[0]
            i1 = np.flatnonzero((gt_v == uniq[e[1]]).all(axis=1))[0]
            key = np.array(sorted([i0, i1]))
            if ((gt_e == key).all(axis=1)).any():
                remaining = remaining[np.any(remaining != key, axis=1)]
            else:
                wed_e += seg_len(uniq[e[0]], uniq[e[1]])
    else:
        wed_v = 0
    for e in remaining:
        wed_e += seg_len(gt_v[e[0]], gt_v[e[1]])
    total = 0
    for e in gt_e:
        total += seg_len(gt_v[e[0]], gt_v[e[1]])
    return (wed_e + wed_v) / total


def sample_metrics(pred_corners, pred_edges, pred_edge_pts, gt_corners, gt_edges, gt_edge_pts,
                   thresh: float) -> Dict[str, float]:
    """One iteration of the loop at eval/ap_calculator.py:149-259."""
    out: Dict[str, float] = {}
    if len(pred_edges) != 0:
        dist = hausdorff_lines(pred_edge_pts, gt_edge_pts)
        pi, li = wo.lsap(np.ascontiguousarray(dist, dtype=np.float64))
        ok = dist[pi, li] <= thresh
        pr_c = pred_edge_pts[pi[ok]]
        gt_c = gt_edge_pts[li[ok]]
        pr_used = unique_rows(pr_c.reshape(-1, 3))
        gt_used = unique_rows(gt_c.reshape(-1, 3))
        free_pr = rows_not_in(pred_corners, pr_used)
        free_gt = rows_not_in(gt_corners, gt_used)
        dm = cdist_euclid(free_pr, free_gt)
        if dm.size:
            fi, fj = wo.lsap(np.ascontiguousarray(dm))
        else:
            fi = fj = np.zeros(0, dtype=np.int64)
        fok = dm[fi, fj] <= thresh
        distances = np.sum(dm[fi[fok], fj[fok]])
        out["tp_corners"] = len(pr_used) + int(fok.sum())
        out["tp_fp_corners"] = len(pred_corners)
        out["tp_fn_corners"] = len(gt_corners)
        out["tp_edges"] = int(ok.sum())
        out["tp_fp_edges"] = len(pred_edges)
        out["tp_fn_edges"] = len(gt_edges)
        # no matched edge -> (0,0) matrix -> numpy's "zero-size array to reduction" ValueError, as in the reference (:227)
        distances = distances + np.sum(np.min(cdist_euclid(pr_used, gt_used), axis=1))
        # wireframe edit distance: the reference rebuilds the submission from the LABEL edges (:232-237)
        sub_v = unique_rows(gt_edge_pts.reshape(-1, 3))
        sub_e = index_edges(gt_edge_pts, sub_v)
        out["wed"] = edit_distance(sub_v, sub_e.copy(), gt_corners.copy(), gt_edges.copy(), distances)
        out["distance"] = float(distances)
    else:
        dm = cdist_euclid(pred_corners, gt_corners)
        pi, li = wo.lsap(np.ascontiguousarray(dm))
        ok = dm[pi, li] <= thresh
        out["distance"] = float(np.sum(dm[pi[ok], li[ok]]))
        out["tp_corners"] = int(ok.sum())
        out["tp_fp_corners"] = len(pred_corners)
        out["tp_fn_corners"] = len(gt_corners)
        out["tp_edges"] = 0
        out["tp_fp_edges"] = 0
        out["tp_fn_edges"] = len(gt_edges)
        out["wed"] = 1
    return out


KEYS = ("tp_corners", "tp_fp_corners", "tp_fn_corners", "distance", "tp_edges", "wed", "tp_fp_edges", "tp_fn_edges")


def accumulate(samples, last_batch_size: int = 1) -> Dict[str, float]:
    """Sum per-sample metrics the way compute_metrics does (:262-272), then output_accuracy (:274-292).
    `average_wed` divides by the size of the LAST batch passed to compute_metrics (:143,276); evaluate.py
    feeds one sample per call."""
    tot = {k: 0 for k in KEYS}
    for s in samples:
        for k in KEYS:
            tot[k] += s[k]
    tot["average_corner_offset"] = tot["distance"] / tot["tp_corners"] if tot["tp_corners"] > 0 else 0.0
    tot["average_wed"] = tot["wed"] / last_batch_size if last_batch_size > 0 else 0.0
    for kind, key in (("corners", "corners"), ("edges", "edges")):
        p = tot[f"tp_{kind}"] / tot[f"tp_fp_{kind}"] if tot[f"tp_fp_{kind}"] > 0 else 0.0
        r = tot[f"tp_{kind}"] / tot[f"tp_fn_{kind}"] if tot[f"tp_fn_{kind}"] > 0 else 0.0
        tot[f"{key}_precision"], tot[f"{key}_recall"] = p, r
        tot[f"{key}_f1"] = 2 * p * r / (p + r) if p + r > 0 else 0.0
    return tot


def edge_points(vertices: np.ndarray, edges: np.ndarray) -> np.ndarray:
    """evaluate.py:87-98: endpoints of each edge, the endpoint with the larger z first (argsort+flip:
    on equal z the SECOND endpoint comes first)."""
    if len(edges) == 0:
        return np.empty((0, 2, 3))
    pts = np.stack((vertices[edges[:, 0]], vertices[edges[:, 1]]), axis=1)
    order = np.flip(np.argsort(pts[:, :, -1]), axis=1)
    return pts[np.arange(len(pts))[:, None], order]


def eval_sample_inputs(pred_vertices: np.ndarray, edge_indices, edge_probs: np.ndarray,
                       gt_vertices: np.ndarray, gt_edges: np.ndarray) -> Tuple[np.ndarray, ...]:
    """evaluate.py:74-98 for one sample: threshold the edge probabilities at 0.5 and attach endpoints."""
    keep = edge_probs > 0.5
    pd_edges = np.array(edge_indices)[keep]
    return (pred_vertices, pd_edges, edge_points(pred_vertices, pd_edges),
            gt_vertices, gt_edges, edge_points(gt_vertices, gt_edges))
