"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's target preparation (train.py:48-88,112-115), the checker
for wf_b200.targets.prepare_targets.  Pure-Python loops like the reference's, small cases only.  Pinned to the reference by
tests/golden/targets.npz, which tests/golden/make_golden_targets.py produced with the reference's own
models.utils.create_edge_labels_from_edge_set (imported from /root/reference) inside the same loop structure."""
import torch


def edge_labels_from_edge_set(edge_set, edge_indices):
    """models/utils.py:24-36"""
    labels = torch.zeros(1, len(edge_indices))
    for k, (i, j) in enumerate(edge_indices):
        if (min(i, j), max(i, j)) in edge_set:
            labels[0, k] = 1
    return labels


def prepare_targets(wf_vertices, wf_edges, max_vertices, label_fn=edge_labels_from_edge_set):
    B = len(wf_vertices)
    existence = torch.zeros(B, max_vertices)                                   # train.py:52
    counts = []
    for i in range(B):
        c = len(wf_vertices[i]); counts.append(c)                              # train.py:56-57
        existence[i, :c] = 1.0                                                 # train.py:58
    counts_t = torch.tensor(counts, dtype=torch.long)                          # train.py:60
    label_list = []
    for i in range(B):
        c = counts[i]
        edge_set = set()
        for edge in wf_edges[i]:                                               # train.py:69-71
            v1, v2 = edge[0].item(), edge[1].item()
            edge_set.add((min(v1, v2), max(v1, v2)))
        pairs = [(j, k) for j in range(c) for k in range(j + 1, c)]            # train.py:74
        label_list.append(label_fn(edge_set, pairs).squeeze(0))                # train.py:77-78
    max_e = max([len(l) for l in label_list]) if label_list else 0             # train.py:81
    labels = torch.zeros(B, max_e)
    for i, l in enumerate(label_list):                                         # train.py:84-86
        if len(l) > 0:
            labels[i, :len(l)] = l
    tv = torch.zeros(B, max_vertices, 3)                                       # train.py:112
    for i in range(B):
        tv[i, :counts[i]] = wf_vertices[i][:counts[i]]                         # train.py:114-115
    return {"vertices": tv, "vertex_existence": existence, "edge_labels": labels, "vertex_counts": counts_t}


def make_case(seed, B, V, min_c=2):
    """Ragged ground truth shaped like collate_batch's lists (datasets/building3d.py:171-190): float32 edges, duplicates,
    reversed endpoints, a self loop and an out-of-range endpoint included on purpose."""
    g = torch.Generator().manual_seed(seed)
    verts, edges = [], []
    for b in range(B):
        c = int(torch.randint(min_c, V + 1, (1,), generator=g))
        verts.append(torch.rand(c, 3, generator=g) - 0.5)
        ne = int(torch.randint(0, 3 * c, (1,), generator=g))
        e = torch.randint(0, c, (ne, 2), generator=g)
        if ne >= 3:
            e[1] = e[0].flip(0)                       # reversed duplicate
            e[2, 1] = e[2, 0]                         # self loop
        edges.append(e.to(torch.float32))
    return verts, edges
