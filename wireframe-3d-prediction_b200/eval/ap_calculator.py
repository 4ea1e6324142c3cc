"""Drop-in for the reference's `eval/ap_calculator.py` (APCalculator :111-307 and the module functions
:8-108) -- the step right after the hot path in `evaluate.py:60,110,112`.

Same class, constructor, `compute_metrics(batch)` dictionary contract, `output_accuracy()` report and
`reset()`.  What changes is where the arithmetic runs and how it is batched: the reference loops over
the samples and, for each, calls scipy's cdist three or four times (one of them on a 40 320 x 1 800 point
matrix for 2 016 predicted edges) and linear_sum_assignment two or three times.  Here a call handles the
whole batch in two device phases through `wf_b200.evalpost` (fp64 kernels of libwf_b200.so):

  phase 1  Hausdorff matrices of all samples with predicted edges (1 launch), corner distance matrices of
           the samples without (1 launch), every assignment problem (1 launch);
  host     the reference's set logic on <= 64 corners per sample (unique rows, set difference);
  phase 2  the three remaining distance matrices of every sample (1 launch) and the free-corner
           assignment problems (1 launch).

Integer results (tp/fp/fn counts, assignments) are identical to the reference's, distances bit-equal
fp64 (tests/test_gpu_evalpost.py; the host logic alone, with scipy in place of the kernels: tests/test_ap_host_logic.py).  Kept on purpose: the wireframe edit distance is computed from the
LABEL edges (:232-237), `average_wed` divides by the size of the last batch (:143,276), a sample whose
predicted edges all miss the threshold raises numpy's zero-size-reduction ValueError (:227), matched
predicted segments are overwritten in the caller's array (:233-234).  Not kept: the debug prints of
:167-172 and :213-223.
"""
import numpy as np

from wf_b200 import evalpost

_COUNTERS = ('tp_corners', 'tp_fp_corners', 'tp_fn_corners', 'distance', 'tp_edges', 'wed', 'tp_fp_edges',
             'tp_fn_edges')


def hausdorff_distance_line(p_line, t_line, sample_points=20):
    """(N,2,3) x (M,2,3) -> N x M symmetric Hausdorff distance of `sample_points` samples per segment."""
    if p_line.shape[0] == 0:
        return np.array([])
    return evalpost.hausdorff_lines_batched([p_line], [t_line], sample_points)[0]


def _first_index_of_rows(table):
    """row bytes -> first index, for exact (bitwise) row look-ups in a small vertex table"""
    first = {}
    table = np.asarray(table) + 0.0                      # -0.0 -> +0.0: numpy's == treats them as equal
    for k in range(len(table) - 1, -1, -1):
        first[table[k].tobytes()] = k
    return first


def _key(row, dtype):
    return (np.asarray(row, dtype=dtype) + 0.0).tobytes()


def computer_edges(edges, vertices):
    """Index pair of every segment's endpoints in `vertices` (-1 when absent), each pair ascending."""
    vertices = np.asarray(vertices)
    where = _first_index_of_rows(vertices)
    idx = [[where.get(_key(pt, vertices.dtype), -1) for pt in seg] for seg in edges]
    return np.sort(np.array(idx), axis=-1)


def remove_corners(corner_a, corner_b):
    """Sorted unique rows of corner_a that do not occur in corner_b."""
    as_records = [('', corner_a.dtype)] * corner_a.shape[1]
    keep = np.setdiff1d(corner_a.view(as_records), corner_b.view(as_records))
    return keep.view(corner_a.dtype).reshape(-1, corner_a.shape[1])


def _running_sum(values):
    """sum of a 1-D array in ITS dtype, left to right (what `acc = 0; for v in values: acc += v` gives)"""
    return np.cumsum(values)[-1] if len(values) else 0


def _edge_lengths(vertices, edges):
    """np.linalg.norm(vertices[a] - vertices[b]) per edge, in the vertices' dtype (float32 from the loader)"""
    d = vertices[edges[:, 0]] - vertices[edges[:, 1]]
    return np.sqrt(np.einsum('ij,ij->i', d, d))


def _edit_distance_from(dist, pd_vertices, pd_edges, gt_vertices, gt_edges, wed_v):
    """graph_edit_distance (:39-84) given dist = cdist(pd_vertices, gt_vertices).  The reference walks the submitted
    edges one by one (a numpy row comparison against every label edge each time); the outcome only depends on which
    index pairs occur, so the walk is done on integer pair codes."""
    left = gt_edges
    wrong = np.zeros(0, dtype=gt_vertices.dtype)             # lengths of submitted edges that are not label edges
    if len(pd_vertices) > 0:
        wed_v += sum(np.min(dist, axis=1))
        snapped = pd_vertices.copy()
        snapped[:] = gt_vertices[np.argmin(dist, axis=1)]
        merged, new_id = np.unique(snapped, axis=0, return_inverse=True)
        new_id = np.asarray(new_id).reshape(-1)
        edges = np.unique(np.where(pd_edges >= 0, new_id[pd_edges], pd_edges), axis=0)
        where = _first_index_of_rows(gt_vertices)
        first = np.array([where[_key(row, gt_vertices.dtype)] for row in merged], dtype=np.int64)
        pairs = np.sort(first[edges], axis=1)                            # (:66) sorted([e1_index[0], e2_index[0]])
        span = int(max(len(gt_vertices), gt_edges.max(initial=0) + 1))
        code = pairs[:, 0] * span + pairs[:, 1]
        gt_code = gt_edges[:, 0].astype(np.int64) * span + gt_edges[:, 1].astype(np.int64)
        present = np.isin(code, gt_code)
        left = gt_edges[~np.isin(gt_code, code[present])]                # every label edge equal to a submitted pair goes
        wrong = _edge_lengths(merged, edges[~present])
    else:
        wed_v = 0
    # the reference adds the lengths one by one to an int 0: wrong submissions first, then the label edges left over
    wed_e = _running_sum(np.concatenate((wrong, _edge_lengths(gt_vertices, left))))
    whole = _running_sum(_edge_lengths(gt_vertices, gt_edges))
    return (wed_e + wed_v) / whole


def graph_edit_distance(pd_vertices, pd_edges, gt_vertices, gt_edges, wed_v):
    dist = evalpost.cdist_batched([pd_vertices], [gt_vertices])[0] if len(pd_vertices) > 0 else None
    return _edit_distance_from(dist, pd_vertices, pd_edges, gt_vertices, gt_edges, wed_v)


class APCalculator(object):
    def __init__(self, distance_thresh=0.1, confidence_thresh=0.7):
        self.distance_thresh = distance_thresh
        self.confidence_thresh = confidence_thresh
        self.batch_size = 0
        self.reset()

    def reset(self):
        self.ap_dict = {'tp_corners': 0, 'tp_fp_corners': 0, 'tp_fn_corners': 0, 'distance': 0, 'tp_edges': 0,
                        'wed': 0, 'tp_fp_edges': 0, 'tp_fn_edges': 0, 'average_corner_offset': 0,
                        'corners_precision': 0, 'corners_recall': 0, 'corner_f1': 0, 'edges_precision': 0,   # 'corner_f1': the reference's key (ap_calculator.py:119), never updated; 'corners_f1' appears in output_accuracy
                        'edges_recall': 0, 'edges_f1': 0}

    # ------------------------------------------------------------------------------------------------
    def compute_metrics(self, batch):
        n = len(batch['predicted_vertices'])
        self.batch_size = n
        corners, edges, seg = batch['predicted_vertices'], batch['predicted_edges'], batch['pred_edges_vertices']
        gt_corners, gt_edges, gt_seg = batch['wf_vertices'], batch['wf_edges'], batch['wf_edges_vertices']
        thr = self.distance_thresh
        has_edges = [b for b in range(n) if len(edges[b]) != 0]
        no_edges = [b for b in range(n) if len(edges[b]) == 0]

        # ---- phase 1: every first-level distance matrix and assignment of the batch; the Hausdorff matrices stay
        #      on the device, only the assignment and its distances come back
        first = dict(zip(has_edges, evalpost.hausdorff_assign_batched([seg[b] for b in has_edges],
                                                                      [gt_seg[b] for b in has_edges])))
        corner_d = evalpost.cdist_batched([corners[b] for b in no_edges], [gt_corners[b] for b in no_edges])
        for b, d, (pi, li) in zip(no_edges, corner_d, evalpost.lsap_batched_f64(corner_d)):
            first[b] = (pi, li, d[pi, li])

        # ---- host: matched segments -> used / free corners (per sample, tiny)
        state = {}
        for b in has_edges:
            pi, li, matched = first[b]
            hit = matched <= thr
            pr_pts = np.unique(seg[b][pi[hit]].reshape(-1, 3), axis=0)
            gt_pts = np.unique(gt_seg[b][li[hit]].reshape(-1, 3), axis=0)
            sub_v, sub_idx = np.unique(gt_seg[b].reshape(-1, 3), axis=0, return_inverse=True)
            sub_e = np.sort(np.asarray(sub_idx).reshape(-1, 2), axis=-1)     # = computer_edges(gt_seg[b], sub_v)
            state[b] = dict(hit=hit, pi=pi, li=li, pr_pts=pr_pts, gt_pts=gt_pts, sub_v=sub_v, sub_e=sub_e,
                            free_pr=remove_corners(corners[b], pr_pts), free_gt=remove_corners(gt_corners[b], gt_pts))

        # ---- phase 2: free-corner matrices (+ assignment), used-corner offsets, edit-distance snapping
        a_list, b_list = [], []
        for b in has_edges:
            s = state[b]
            a_list += [s['free_pr'], s['pr_pts'], s['sub_v']]
            b_list += [s['free_gt'], s['gt_pts'], gt_corners[b]]
        second = evalpost.cdist_batched(a_list, b_list)
        free_assign = evalpost.lsap_batched_f64([second[3 * k] for k in range(len(has_edges))])

        # ---- per-sample accounting in batch order (a failing sample raises after its predecessors counted)
        slot = {b: k for k, b in enumerate(has_edges)}
        for b in range(n):
            if b in slot:
                s, k = state[b], slot[b]
                free_d, used_d, snap_d = second[3 * k], second[3 * k + 1], second[3 * k + 2]
                fi, fj = free_assign[k]
                free_hit = free_d[fi, fj] <= thr
                distances = np.sum(free_d[fi[free_hit], fj[free_hit]])
                tp_corners = len(s['pr_pts']) + sum(free_hit)
                tp_fp_corners, tp_fn_corners = len(corners[b]), len(gt_corners[b])
                tp_edges, tp_fp_edges, tp_fn_edges = sum(s['hit']), len(edges[b]), len(gt_edges[b])
                distances += np.sum(np.min(used_d, axis=1))         # zero matched edges: numpy raises here
                for j, i in enumerate(s['pi'][s['hit']]):
                    seg[b][i] = gt_seg[b][s['li'][s['hit']][j]]
                wed = _edit_distance_from(snap_d, s['sub_v'], s['sub_e'].copy(), gt_corners[b].copy(),
                                          gt_edges[b].copy(), distances)
            else:
                pi, li, matched = first[b]
                hit = matched <= thr
                distances = np.sum(matched[hit])
                tp_corners, tp_fp_corners, tp_fn_corners = len(pi[hit]), len(corners[b]), len(gt_corners[b])
                tp_edges, tp_fp_edges, tp_fn_edges, wed = 0, 0, len(gt_edges[b]), 1
            for key, val in zip(_COUNTERS, (tp_corners, tp_fp_corners, tp_fn_corners, distances, tp_edges, wed,
                                            tp_fp_edges, tp_fn_edges)):
                self.ap_dict[key] += val

    # ------------------------------------------------------------------------------------------------
    def output_accuracy(self):
        d = self.ap_dict
        ratio = lambda a, b: a / b if b > 0 else 0.0                    # noqa: E731
        f1 = lambda p, r: 2 * p * r / (p + r) if (p + r) > 0 else 0.0   # noqa: E731
        d['average_corner_offset'] = ratio(d['distance'], d['tp_corners'])
        d['average_wed'] = ratio(d['wed'], self.batch_size)
        d['corners_precision'] = ratio(d['tp_corners'], d['tp_fp_corners'])
        d['corners_recall'] = ratio(d['tp_corners'], d['tp_fn_corners'])
        d['corners_f1'] = f1(d['corners_precision'], d['corners_recall'])
        d['edges_precision'] = ratio(d['tp_edges'], d['tp_fp_edges'])
        d['edges_recall'] = ratio(d['tp_edges'], d['tp_fn_edges'])
        d['edges_f1'] = f1(d['edges_precision'], d['edges_recall'])

        print('Wireframe Edit distance', d['average_wed'])
        print('Average Corner offset', d['average_corner_offset'])
        print('Corners Precision: ', d['corners_precision'])
        print('Corners Recall: ', d['corners_recall'])
        print('Corners F1：', d['corners_f1'])
        print('Edges Precision: ', d['edges_precision'])
        print('Edges Recall: ', d['edges_recall'])
        print('Edges F1: ', d['edges_f1'])
