"""Target preparation: the oracle restatement of train.py:48-88,112-115 against the golden file made with
the reference's own label function (CPU), and the device kernel against the oracle and the golden (GPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import targets_oracle as to

KEYS = ("vertices", "vertex_existence", "edge_labels", "vertex_counts")


def _golden(golden_dir):
    return dict(np.load(os.path.join(golden_dir, "targets.npz")))


def test_oracle_matches_reference_golden(golden_dir):
    g = _golden(golden_dir)
    for seed, B, V in g["cases"].tolist():
        verts, edges = to.make_case(seed, B, V)
        t = to.prepare_targets(verts, edges, V)
        for k in KEYS:
            assert np.array_equal(t[k].numpy(), g[f"{seed}/{k}"]), (seed, k)


@pytest.mark.gpu
def test_prepare_targets_kernel_bit_exact(golden_dir):
    from wf_b200.targets import prepare_targets
    g = _golden(golden_dir)
    for seed, B, V in g["cases"].tolist() + [(21, 64, 64), (22, 1, 2), (23, 7, 33)]:
        verts, edges = to.make_case(seed, B, V)
        got = prepare_targets(verts, edges, V, "cuda")
        ref = to.prepare_targets(verts, edges, V)
        for k in KEYS:
            assert got[k].dtype == ref[k].dtype and got[k].shape == ref[k].shape, (seed, k)
            assert torch.equal(got[k].cpu(), ref[k]), (seed, k)
            if f"{seed}/{k}" in g:
                assert np.array_equal(got[k].cpu().numpy(), g[f"{seed}/{k}"]), (seed, k)
    # no edges at all / integer edge dtype
    verts = [torch.rand(3, 3), torch.rand(5, 3)]
    edges = [torch.zeros(0, 2), torch.tensor([[0, 4], [3, 1]], dtype=torch.int64)]
    got = prepare_targets(verts, edges, 8, "cuda"); ref = to.prepare_targets(verts, edges, 8)
    for k in KEYS:
        assert torch.equal(got[k].cpu(), ref[k]), k
    with pytest.raises(ValueError):
        prepare_targets([torch.rand(9, 3)], [torch.zeros(0, 2)], 8, "cuda")


@pytest.mark.gpu
def test_device_prefetcher_delivers_batches_in_order():
    from wf_b200.targets import DevicePrefetcher
    torch.manual_seed(0)
    host = [{"point_clouds": torch.rand(4, 300, 8), "vertex_counts": torch.randint(2, 9, (4,)), "tag": i} for i in range(5)]
    seen = []
    for i, b in enumerate(DevicePrefetcher(host, "cuda")):
        assert b["point_clouds"].is_cuda and b["tag"] == i
        assert torch.equal(b["point_clouds"].cpu(), host[i]["point_clouds"])
        assert torch.equal(b["vertex_counts"].cpu(), host[i]["vertex_counts"])
        seen.append(i)
    assert seen == list(range(5))
