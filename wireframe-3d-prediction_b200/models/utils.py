"""Host helpers that train.py imports from `models.utils` (reference models/utils.py:24-36).  These are
target-preparation code outside the hot path; kept importable with the same names and semantics."""
import numpy as np
import torch


def create_edge_labels_from_edge_set(edge_set, edge_indices):
    """0/1 label per candidate edge; (1, num_edges) float tensor like the reference."""
    edge_labels = torch.zeros(1, len(edge_indices))
    for k, (i, j) in enumerate(edge_indices):
        if (min(i, j), max(i, j)) in edge_set:
            edge_labels[0, k] = 1
    return edge_labels


def create_adjacency_matrix_from_predictions(edge_probs, edge_indices, num_vertices, threshold=0.5):
    batch_size = edge_probs.shape[0]
    adj = torch.zeros(batch_size, num_vertices, num_vertices)
    for b in range(batch_size):
        for k, (i, j) in enumerate(edge_indices):
            if edge_probs[b, k] > threshold:
                adj[b, i, j] = 1
                adj[b, j, i] = 1
    return adj


def hungarian_rmse(pred_vertices, true_vertices):
    """RMSE under the optimal L2 matching (reference models/utils.py:38-55): fp64 pairwise distances and the fp64
    assignment both run on the device (wf_cdist_f64 + wf_lsap_f64, the kernels of the evaluation post-processing), scipy's
    ValueError for invalid entries is raised like the reference's linear_sum_assignment call would."""
    if len(pred_vertices) == 0 and len(true_vertices) == 0:
        return 0.0
    if len(pred_vertices) == 0 or len(true_vertices) == 0:
        return float('inf')
    from wf_b200 import evalpost
    pv, tv = np.asarray(pred_vertices), np.asarray(true_vertices)
    # scipy's cdist works in float64 whatever the inputs are; the final expression keeps the inputs' own dtype (:52-55)
    rows, cols, _ = evalpost.cdist_assign_batched([pv.astype(np.float64)], [tv.astype(np.float64)])[0]
    return np.sqrt(np.mean((pv[rows] - tv[cols]) ** 2))
