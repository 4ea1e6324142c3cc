"""Pins oracle/wireframe_oracle.py against fixtures generated from the unmodified reference
(tests/golden/make_golden.py).  fp32 vs fp32 on the same CPU kernels, so the tolerance is tight;
matchings and counts must be identical."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import wireframe_oracle as wo

TRAIN = ["train_b2_n384_v12", "train_b3_n300_v20_pad", "train_b1_n256_v8_rawint", "train_b2_n10000_v64"]


def _load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name + ".npz")))


@pytest.mark.parametrize("name", TRAIN)
def test_train_step_matches_reference(golden_dir, name):
    g = _load(golden_dir, name)
    seed, B, N, V, pad, norm_i = [int(v) for v in g["meta"]]
    sd = {k: v.clone().requires_grad_(True) for k, v in wo.make_state_dict(seed, V).items()}
    cmin, cmax = [int(v) for v in g["count_range"]]
    x, tgt, counts = wo.make_inputs(seed, B, N, V, pad_frac=pad / 1000.0, norm_intensity=bool(norm_i), min_count=cmin,
                                    max_count=None if cmax < 0 else cmax)
    x = x.clone().requires_grad_(True)
    ld, pred = wo.train_step(sd, x, tgt, max_vertices=V)
    tol = dict(rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(pred["vertices"].detach().numpy(), g["vertices"], **tol)
    np.testing.assert_allclose(pred["existence_probabilities"].detach().numpy(), g["existence"], **tol)
    np.testing.assert_allclose(pred["edge_probs"].detach().numpy(), g["edge_probs"], **tol)
    np.testing.assert_allclose(pred["global_features"].detach().numpy(), g["global_features"], **tol)
    assert np.array_equal(pred["actual_vertex_counts"].numpy(), g["dyn_counts"])
    assert [len(e) for e in pred["edge_indices"]] == g["n_edges"].tolist()
    got = np.array([ld[k].item() for k in ("total_loss", "vertex_loss", "existence_loss", "edge_loss")])
    np.testing.assert_allclose(got, g["losses"], rtol=1e-5, atol=1e-6)
    for b, (pi, ti) in enumerate(wo.loss_matching(pred, tgt)):
        assert np.array_equal(pi, g[f"match_p/{b}"]) and np.array_equal(ti, g[f"match_t/{b}"])
    # gradients: every parameter's digest
    for k, p in sd.items():
        if "gnone/" + k in g:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        gr = p.grad.detach().double().reshape(-1)
        ref_norm, ref_sum = g["gnorm/" + k]
        assert abs(gr.norm().item() - ref_norm) <= 2e-3 * ref_norm + 1e-7, k
        scale = max(float(np.abs(g["ghead/" + k]).max()), ref_norm / np.sqrt(gr.numel()), 1e-8)
        np.testing.assert_allclose(gr[:16].float().numpy(), g["ghead/" + k], rtol=5e-3, atol=5e-3 * scale, err_msg=k)
        stride = max(1, -(-gr.numel() // 1024))
        np.testing.assert_allclose(gr[::stride].float().numpy(), g["gsamp/" + k], rtol=5e-3, atol=5e-3 * scale, err_msg=k)
    nrow = g["dx"].shape[1]
    np.testing.assert_allclose(x.grad.numpy()[:, :nrow], g["dx"], rtol=5e-3, atol=5e-3 * float(np.abs(g["dx"]).max()))
    # pooling internals incl. argmax (first maximal index)
    with torch.no_grad():
        pf = wo.encoder_point_features({k: v.detach() for k, v in sd.items()}, x.detach())
        mx, arg = pf.max(dim=1)
    np.testing.assert_allclose(mx.numpy(), g["pf_max"], **tol)
    decided = g["pf_top2_gap"] > 1e-5 * np.abs(g["pf_max"])
    assert (arg.numpy() == g["pf_argmax"])[decided].all() or "rawint" in name
    assert (arg.numpy() == g["pf_argmax"]).mean() > 0.999


@pytest.mark.parametrize("name", ["eval_b2_n256_v16", "eval_b2_n10000_v64"])
def test_eval_forward_matches_reference(golden_dir, name):
    g = _load(golden_dir, name)
    seed, B, N, V = [int(v) for v in g["meta"][:4]]
    sd = wo.make_state_dict(seed, V)
    if g["exist_bias"].size:
        sd["vertex_predictor.final_layer.bias"].view(V, 4)[:, 3] += torch.from_numpy(g["exist_bias"])
    x, tgt, counts = wo.make_inputs(seed, B, N, V, norm_intensity=True)
    with torch.no_grad():
        pred = wo.model_forward(sd, x, counts, training=False, max_vertices=V)
    tol = dict(rtol=2e-4, atol=2e-5)
    assert np.array_equal(pred["actual_vertex_counts"].numpy(), g["dyn_counts"])
    np.testing.assert_allclose(pred["vertices"].numpy(), g["vertices"], **tol)
    np.testing.assert_allclose(pred["edge_probs"].numpy(), g["edge_probs"], **tol)
    assert np.array_equal(np.asarray(pred["edge_indices"][0]), g["edge_indices0"])


def test_matchers_match_reference(golden_dir):
    g = _load(golden_dir, "matchers")
    seed, B, V = [int(v) for v in g["meta"]]
    rng = np.random.Generator(np.random.PCG64(seed))
    outputs = {"vertices": torch.from_numpy(rng.uniform(-1, 1, (B, V, 3)).astype(np.float32)),
               "existence_probabilities": torch.from_numpy(rng.uniform(0, 1, (B, V)).astype(np.float32))}
    sizes = g["sizes"].tolist()
    tg = [{"vertices": torch.from_numpy(rng.uniform(-1, 1, (t, 3)).astype(np.float32)),
           "existence": torch.ones(t)} for t in sizes]
    res = wo.wireframe_matcher(outputs, tg, cost_vertex=2.0, cost_existence=0.5)
    for b, (i, j) in enumerate(res):
        assert np.array_equal(i.numpy(), g[f"wf_i/{b}"]) and np.array_equal(j.numpy(), g[f"wf_j/{b}"])
    Q, K = 20, 7
    det = {"pred_logits": torch.from_numpy(rng.normal(size=(B, Q, K)).astype(np.float32)),
           "pred_boxes": torch.from_numpy(np.concatenate([rng.uniform(0.2, 0.8, (B, Q, 2)),
                                                          rng.uniform(0.05, 0.3, (B, Q, 2))], -1).astype(np.float32))}
    dt = [{"labels": torch.from_numpy(rng.integers(0, K, (t,)).astype(np.int64)),
           "boxes": torch.from_numpy(np.concatenate([rng.uniform(0.2, 0.8, (t, 2)),
                                                     rng.uniform(0.05, 0.3, (t, 2))], -1).astype(np.float32))}
          for t in g["detr_sizes"].tolist()]
    res = wo.detr_matcher(det, dt, cost_class=1.0, cost_bbox=5.0, cost_giou=2.0)
    for b, (i, j) in enumerate(res):
        assert np.array_equal(i.numpy(), g[f"detr_i/{b}"]) and np.array_equal(j.numpy(), g[f"detr_j/{b}"])


def test_loss_ties_match_reference(golden_dir):
    g = _load(golden_dir, "loss_ties")
    seed, B, V = [int(v) for v in g["meta"]]
    rng = np.random.Generator(np.random.PCG64(seed))
    pv = np.round(rng.uniform(-1, 1, (B, V, 3)) * 2) / 2
    pe = np.round(rng.uniform(0, 1, (B, V)) * 4) / 4
    tv = np.zeros((B, V, 3)); counts = g["counts"]
    for b in range(B):
        tv[b, :counts[b]] = np.round(rng.uniform(-1, 1, (counts[b], 3)) * 2) / 2
    pred = {"vertices": torch.from_numpy(pv.astype(np.float32)),
            "existence_probabilities": torch.from_numpy(pe.astype(np.float32))}
    tgt = {"vertices": torch.from_numpy(tv.astype(np.float32)), "vertex_counts": torch.from_numpy(counts)}
    for b, (i, j) in enumerate(wo.loss_matching(pred, tgt)):
        assert np.array_equal(i, g[f"p/{b}"]) and np.array_equal(j, g[f"t/{b}"])


def test_golden_files_present(golden_dir):
    assert len(glob.glob(os.path.join(golden_dir, "*.npz"))) >= 6


def test_real_building_matches_reference(golden_dir):
    """BASELINE.json configs[0]: a real building of the shipped dataset as the reference's loader delivers it (2560 points,
    un-normalised intensity), targets as train.py builds them, one training step -- oracle vs the unmodified reference."""
    from oracle import targets_oracle as to
    g = _load(golden_dir, "real_b1_n2560_v38")
    seed, B, N, V = [int(v) for v in g["meta"]]
    wf_v, wf_e = [torch.from_numpy(g["wf_vertices"])], [torch.from_numpy(g["wf_edges"])]
    tgt = to.prepare_targets(wf_v, wf_e, V)
    assert np.array_equal(tgt["edge_labels"].numpy(), g["tgt_edge_labels"]) and np.array_equal(tgt["vertices"].numpy(), g["tgt_vertices"])
    assert np.array_equal(tgt["vertex_counts"].numpy(), g["tgt_counts"])
    sd = {k: v.clone().requires_grad_(True) for k, v in wo.make_state_dict(seed, V).items()}
    ld, pred = wo.train_step(sd, torch.from_numpy(g["x"]), tgt, max_vertices=V)
    tol = dict(rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(pred["vertices"].detach().numpy(), g["vertices"], **tol)
    np.testing.assert_allclose(pred["edge_probs"].detach().numpy(), g["edge_probs"], **tol)
    got = np.array([ld[k].item() for k in ("total_loss", "vertex_loss", "existence_loss", "edge_loss")])
    np.testing.assert_allclose(got, g["losses"], rtol=1e-5, atol=1e-6)
    pi, ti = wo.loss_matching(pred, tgt)[0]
    assert np.array_equal(pi, g["match_p"]) and np.array_equal(ti, g["match_t"])
    for k, p in sd.items():
        if "gnone/" + k in g:
            continue
        ref_norm = g["gnorm/" + k][0]
        assert abs(p.grad.double().norm().item() - ref_norm) <= 5e-3 * ref_norm + 1e-7, k


def test_hungarian_rmse_restatement_matches_reference(golden_dir):
    """models/utils.py:38-55: fp64 cdist + fp64 assignment (scipy here, as the reference) -> the stored reference values."""
    from scipy.optimize import linear_sum_assignment
    from scipy.spatial.distance import cdist
    g = _load(golden_dir, "hungarian_rmse")
    for k, want in enumerate(g["rmse"]):
        p, t = g[f"p/{k}"], g[f"t/{k}"]
        r, c = linear_sum_assignment(cdist(p, t))
        assert np.sqrt(np.mean((p[r] - t[c]) ** 2)) == want
