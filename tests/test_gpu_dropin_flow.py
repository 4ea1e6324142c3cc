"""GPU: the call sequences of the reference's two callers, restated line by line against the drop-in modules
(the reference scripts themselves do not travel to the GPU box):

  train.py:25-142     first batch -> model -> host target loops -> WireframeLoss -> Adam built BEFORE the first forward
                      -> epochs of forward / loss / backward / clip_grad_norm_ / step on that same batch
  main.py:53          torch.save(model.state_dict())
  evaluate.py:46-112  max_vertices from the checkpoint's final layer, load_state_dict(strict=False), eval forward with
                      the label counts, per-sample numpy post-processing, APCalculator.compute_metrics / output_accuracy
"""
import contextlib
import io

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def fake_loader_batch(seed, B=3, N=2560, counts=(6, 9, 12)):
    """What collate_batch (datasets/building3d.py:171-190) hands to train.py: points (B,N,8) float32, ragged float32
    vertex and edge lists."""
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(-1, 1, (B, N, 3))
    xyz /= np.linalg.norm(xyz, axis=2).max(axis=1)[:, None, None]
    pts = np.concatenate((xyz, rng.integers(0, 256, (B, N, 4)) / 256.0, rng.uniform(2e4, 6e4, (B, N, 1)) / 65536.0), axis=2)
    counts = list(counts)[:B]
    verts = [torch.tensor(rng.uniform(-0.6, 0.6, (c, 3)).astype(np.float32)) for c in counts]
    edges = []
    for c in counts:
        ring = np.array([[i, (i + 1) % c] for i in range(c)])
        edges.append(torch.tensor(np.sort(ring, axis=1).astype(np.float32)))
    return {"point_clouds": torch.tensor(pts.astype(np.float32)), "wf_vertices": verts, "wf_edges": edges}


def test_train_py_then_evaluate_py_call_sequence(tmp_path):
    from models.PointCloudToWireframe import PointCloudToWireframe
    from models.utils import create_edge_labels_from_edge_set
    from losses.WireframeLoss import WireframeLoss
    from eval.ap_calculator import APCalculator
    from wf_b200 import ops

    ops.set_precision("bf16")
    torch.manual_seed(0)
    device = torch.device("cuda")
    first_batch = fake_loader_batch(0)
    point_clouds, wf_vertices, wf_edges = first_batch["point_clouds"], first_batch["wf_vertices"], first_batch["wf_edges"]
    batch_size = point_clouds.shape[0]
    input_dim = point_clouds.shape[2]
    max_vertices = max(len(v) for v in wf_vertices)                                   # train.py:37
    model = PointCloudToWireframe(input_dim=input_dim, max_vertices=max_vertices).to(device)
    n_before = sum(p.numel() for p in model.parameters())
    point_cloud_tensor = point_clouds.to(device)
    vertex_existence_batch = torch.zeros(batch_size, max_vertices).to(device)
    actual_vertex_counts = []
    for i in range(batch_size):
        actual_vertex_counts.append(len(wf_vertices[i]))
        vertex_existence_batch[i, :len(wf_vertices[i])] = 1.0
    actual_vertex_counts = torch.tensor(actual_vertex_counts, dtype=torch.long).to(device)
    edge_labels_list = []
    for i in range(batch_size):                                                       # train.py:62-78
        c = actual_vertex_counts[i].item()
        edge_set = set()
        for edge in wf_edges[i]:
            v1, v2 = edge[0].item(), edge[1].item()
            edge_set.add((min(v1, v2), max(v1, v2)))
        pairs = [(j, k) for j in range(c) for k in range(j + 1, c)]
        edge_labels_list.append(create_edge_labels_from_edge_set(edge_set, pairs).squeeze(0))
    max_edges = max(len(x) for x in edge_labels_list)
    edge_labels_batch = torch.zeros(batch_size, max_edges).to(device)
    for i, labels in enumerate(edge_labels_list):
        edge_labels_batch[i, :len(labels)] = labels
    criterion = WireframeLoss(vertex_weight=3.0, edge_weight=1.0, existence_weight=1.5)
    assert (criterion.vertex_weight, criterion.edge_weight) == (3.0, 1.0)            # read by train.py:106
    optimizer = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-6, eps=1e-8, betas=(0.9, 0.999))
    model.train()
    target_vertices = torch.zeros(batch_size, max_vertices, 3).to(device)
    for i in range(batch_size):
        c = actual_vertex_counts[i].item()
        target_vertices[i, :c] = wf_vertices[i][:c]

    # the device-side target preparation gives the same four tensors as the loops above
    from wf_b200.targets import prepare_targets
    fast = prepare_targets(wf_vertices, wf_edges, max_vertices, device)
    assert torch.equal(fast["vertices"], target_vertices) and torch.equal(fast["vertex_existence"], vertex_existence_batch)
    assert torch.equal(fast["edge_labels"], edge_labels_batch) and torch.equal(fast["vertex_counts"], actual_vertex_counts)

    history = []
    for epoch in range(60):                                                           # train.py:123-145
        optimizer.zero_grad()
        predictions = model(point_cloud_tensor, actual_vertex_counts)
        targets = {"vertices": target_vertices, "vertex_existence": vertex_existence_batch,
                   "edge_labels": edge_labels_batch, "vertex_counts": actual_vertex_counts}
        loss_dict = criterion(predictions, targets)
        total_loss = loss_dict["total_loss"]
        total_loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)
        optimizer.step()
        history.append(total_loss.item())
        with torch.no_grad():                                                         # train.py:148-151
            c0 = actual_vertex_counts[0].item()
            rmse = np.sqrt(np.mean((predictions["vertices"][0].cpu().numpy()[:c0] - target_vertices[0].cpu().numpy()[:c0]) ** 2))
        assert np.isfinite(rmse)
    assert all(np.isfinite(history))
    assert min(history[-10:]) < 0.7 * history[0], history[::10]                      # the overfit loop does overfit
    # Q1: the lazily created projection joined the module after the optimizer was built
    n_after = sum(p.numel() for p in model.parameters())
    assert n_after - n_before == 1024 * 512 + 512
    in_opt = {id(p) for g in optimizer.param_groups for p in g["params"]}
    assert id(model.vertex_predictor.point_pool_proj.weight) not in in_opt
    assert model.edge_predictor.spatial_proj[0].weight.grad is None                  # Q3
    path = tmp_path / "trained_model.pth"
    torch.save(model.state_dict(), path)                                              # main.py:53

    # ---- evaluate.py:46-112
    state_dict = torch.load(path, map_location=device)
    mv = state_dict["vertex_predictor.final_layer.weight"].shape[0] // 4
    assert mv == max_vertices
    model2 = PointCloudToWireframe(input_dim=input_dim, max_vertices=mv).to(device)
    model2.load_state_dict(state_dict, strict=False)
    model2.eval()
    ap_calculator = APCalculator(distance_thresh=1)
    for k in range(batch_size):                 # one building per batch: the per-sample probability rows of evaluate.py:80-81
        test_batch = {"point_clouds": point_clouds[k:k + 1], "wf_vertices": [wf_vertices[k]], "wf_edges": [wf_edges[k]]}
        with torch.no_grad():
            pcs, gts_v, gts_e = test_batch["point_clouds"], test_batch["wf_vertices"], test_batch["wf_edges"]
            vertex_counts = torch.tensor([len(v) for v in gts_v], dtype=torch.long).to(device)
            predictions = model2(pcs.to(device), vertex_counts)
            for i in range(len(gts_v)):
                pred_vertices = predictions["vertices"][i].cpu().numpy()
                edge_indices = predictions["edge_indices"][i]
                edge_probs = predictions["edge_probs"][i].cpu().numpy()
                mask = edge_probs > 0.5
                pd_edges = np.array(edge_indices)[mask]
                gt_vertices = gts_v[i].numpy()
                gt_edges = gts_e[i].numpy().astype(np.int64)
                if len(pd_edges) > 0:
                    pev = np.stack((pred_vertices[pd_edges[:, 0]], pred_vertices[pd_edges[:, 1]]), axis=1)
                    pev = pev[np.arange(pev.shape[0])[:, np.newaxis], np.flip(np.argsort(pev[:, :, -1]), axis=1)]
                else:
                    pev = np.empty((0, 2, 3))
                gev = np.stack((gt_vertices[gt_edges[:, 0]], gt_vertices[gt_edges[:, 1]]), axis=1)
                gev = gev[np.arange(gev.shape[0])[:, np.newaxis], np.flip(np.argsort(gev[:, :, -1]), axis=1)]
                batch = {"predicted_vertices": pred_vertices[np.newaxis, :], "predicted_edges": pd_edges[np.newaxis, :],
                         "pred_edges_vertices": pev.reshape((1, -1, 2, 3)), "wf_vertices": gt_vertices[np.newaxis, :],
                         "wf_edges": gt_edges[np.newaxis, :], "wf_edges_vertices": gev.reshape((1, -1, 2, 3))}
                # the batched helper builds the same dictionary (compared first: compute_metrics snaps matched segments
                # onto their labels inside the caller's array, eval/ap_calculator.py:233-234)
                from wf_b200.evalpost import make_ap_batch
                fastb = make_ap_batch(predictions, gts_v, gts_e)
                assert np.array_equal(fastb["predicted_edges"][i], pd_edges)
                assert np.array_equal(fastb["pred_edges_vertices"][i], pev) or len(pd_edges) == 0
                # ... and the oracle's restatement of the reference calculator, on the same arrays, must agree with the device path
                from oracle import ap_oracle as ao
                want = ao.sample_metrics(pred_vertices.copy(), pd_edges.copy(), pev.copy(), gt_vertices.copy(), gt_edges.copy(),
                                         gev.copy(), 1)
                before = dict(ap_calculator.ap_dict)
                ap_calculator.compute_metrics(batch)
                for k in ("tp_corners", "tp_fp_corners", "tp_fn_corners", "tp_edges", "tp_fp_edges", "tp_fn_edges"):
                    assert int(ap_calculator.ap_dict[k] - before[k]) == int(want[k]), k
                assert ap_calculator.ap_dict["distance"] - before["distance"] == pytest.approx(want["distance"], rel=1e-12, abs=1e-14)
                assert ap_calculator.ap_dict["wed"] - before["wed"] == pytest.approx(want["wed"], rel=1e-6, abs=1e-9)
    with contextlib.redirect_stdout(io.StringIO()) as text:
        ap_calculator.output_accuracy()
    assert "Corners Precision" in text.getvalue()
    d = ap_calculator.ap_dict
    assert d["tp_fn_corners"] == 6 + 9 + 12 and d["tp_fp_corners"] == 3 * max_vertices
    assert 0.0 <= d["corners_recall"] <= 1.0 and np.isfinite(d["average_wed"])
