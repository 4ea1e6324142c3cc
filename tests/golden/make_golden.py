"""
Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference,
which exists only in the authoring container) on deterministic inputs and weights.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Weights come from oracle.wireframe_oracle.make_state_dict(seed) (numpy PCG64, independent of the
torch RNG), loaded into the reference modules with load_state_dict(strict=True) after one
warm-up forward has materialised the lazy `vertex_predictor.point_pool_proj`.
The four dropout sites of the edge head are disabled.  Outputs are stored; weights
and inputs are not (both regenerate from their seeds), so the fixtures stay small.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)

from models.PointCloudToWireframe import PointCloudToWireframe  # noqa: E402  (reference)
from models.WireframeHungarianMatcher import WireframeHungarianMatcher  # noqa: E402
from models.HungarianMatcher import HungarianMatcher  # noqa: E402
from losses.WireframeLoss import WireframeLoss  # noqa: E402

from oracle import wireframe_oracle as wo  # noqa: E402

torch.set_num_threads(8)
GSAMP = 1024


def build_reference(seed, V, train_mode):
    torch.manual_seed(0)
    m = PointCloudToWireframe(input_dim=8, max_vertices=V)
    m.eval()
    with torch.no_grad():   # materialise the lazy layer; count may be <=1 on random init -> guard
        try:
            m(torch.rand(1, 16, 8))
        except IndexError:
            pass
    assert hasattr(m.vertex_predictor, "point_pool_proj")
    sd = wo.make_state_dict(seed, V)
    m.load_state_dict(sd, strict=True)
    m.train(train_mode)
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.eval()
    m.edge_predictor.attention.dropout = 0.0
    return m, sd


def grad_digest(named_params):
    """Per parameter: L2 norm, sum, and the first 16 entries (flattened)."""
    out = {}
    for k, p in named_params:
        if p.grad is None:
            out["gnone/" + k] = np.zeros(0, np.float32)
            continue
        g = p.grad.detach().double().reshape(-1)
        out["gnorm/" + k] = np.array([g.norm().item(), g.sum().item()], np.float64)
        out["ghead/" + k] = g[:16].float().numpy()
        # element-wise reference values on a fixed stride (<= GSAMP entries per parameter): the bf16 tests compare these
        # one by one with the reference's matching injected
        out["gsamp/" + k] = g[::max(1, -(-g.numel() // GSAMP))].float().numpy()
    return out


def case_train(name, seed, B, N, V, pad_frac=0.0, norm_intensity=True, min_count=2, max_count=None, store_dx=True):
    m, _ = build_reference(seed, V, True)
    x, tgt, counts = wo.make_inputs(seed, B, N, V, pad_frac=pad_frac, norm_intensity=norm_intensity, min_count=min_count,
                                    max_count=max_count)
    xr = x.clone().requires_grad_(True)
    crit = WireframeLoss(vertex_weight=3.0, edge_weight=1.0, existence_weight=1.5)
    pred = m(xr, counts)
    matched = crit._hungarian_matching(pred, tgt)
    ld = crit(pred, tgt)
    ld["total_loss"].backward()
    out = {
        "meta": np.array([seed, B, N, V, int(pad_frac * 1000), int(norm_intensity)], np.int64),
        "count_range": np.array([min_count, -1 if max_count is None else max_count], np.int64),
        "vertices": pred["vertices"].detach().numpy(),
        "existence": pred["existence_probabilities"].detach().numpy(),
        "edge_probs": pred["edge_probs"].detach().numpy(),
        "global_features": pred["global_features"].detach().numpy(),
        "dyn_counts": pred["actual_vertex_counts"].numpy(),
        "n_edges": np.array([len(e) for e in pred["edge_indices"]], np.int64),
        "losses": np.array([ld[k].item() for k in ("total_loss", "vertex_loss", "existence_loss",
                                                    "edge_loss")], np.float64),
        # large clouds: the first 64 points of every cloud only (the fixture stays small)
        "dx": xr.grad.numpy() if store_dx else xr.grad[:, :64].numpy(),
    }
    for b, (pi, ti) in enumerate(matched):
        out[f"match_p/{b}"] = np.asarray(pi, np.int64)
        out[f"match_t/{b}"] = np.asarray(ti, np.int64)
    # encoder internals for the argmax check
    with torch.no_grad():
        g, pf = m.encoder(x)
        out["pf_max"] = pf.max(dim=1).values.numpy()
        out["pf_argmax"] = pf.max(dim=1).indices.numpy()
        out["pf_mean"] = pf.mean(dim=1).numpy()
        # gap between the largest and second-largest value per (cloud, channel): the argmax is only well defined where it
        # exceeds fp32 rounding noise
        top2 = pf.topk(2, dim=1).values
        out["pf_top2_gap"] = (top2[:, 0] - top2[:, 1]).numpy()
    out.update(grad_digest(m.named_parameters()))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "losses", out["losses"], "counts", counts.tolist())


def case_eval(name, seed, B, N, V, exist_bias=None):
    m, sd = build_reference(seed, V, False)
    if exist_bias is not None:
        # random-init existence logits hover around 0; push slot k's logit by exist_bias[k] so that the eval-mode counts
        # (#slots with p > 0.5, models/VertexPredictor.py:124-127) are neither 0/1 (IndexError) nor all V
        sd = {k: v.clone() for k, v in sd.items()}
        sd["vertex_predictor.final_layer.bias"].view(V, 4)[:, 3] += torch.as_tensor(exist_bias, dtype=torch.float32)
        m.load_state_dict(sd, strict=True)
    x, tgt, counts = wo.make_inputs(seed, B, N, V, norm_intensity=True)
    with torch.no_grad():
        pred = m(x, counts)
    out = {
        "meta": np.array([seed, B, N, V, 0, 1], np.int64),
        "exist_bias": np.zeros(0, np.float32) if exist_bias is None else np.asarray(exist_bias, np.float32),
        "vertices": pred["vertices"].numpy(),
        "existence": pred["existence_probabilities"].numpy(),
        "edge_probs": pred["edge_probs"].numpy(),
        "global_features": pred["global_features"].numpy(),
        "dyn_counts": pred["actual_vertex_counts"].numpy(),
        "n_edges": np.array([len(e) for e in pred["edge_indices"]], np.int64),
        "edge_indices0": np.asarray(pred["edge_indices"][0], np.int64),
    }
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "dyn counts", out["dyn_counts"].tolist())


def case_matchers(name, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    B, V = 5, 24
    outputs = {"vertices": torch.from_numpy(rng.uniform(-1, 1, (B, V, 3)).astype(np.float32)),
               "existence_probabilities": torch.from_numpy(rng.uniform(0, 1, (B, V)).astype(np.float32))}
    sizes = [3, 24, 30, 1, 11]          # includes T > V (tall after split) and T == V
    tg = [{"vertices": torch.from_numpy(rng.uniform(-1, 1, (t, 3)).astype(np.float32)),
           "existence": torch.ones(t)} for t in sizes]
    res = WireframeHungarianMatcher(cost_vertex=2.0, cost_existence=0.5)(outputs, tg)
    out = {"meta": np.array([seed, B, V], np.int64), "sizes": np.array(sizes, np.int64)}
    for b, (i, j) in enumerate(res):
        out[f"wf_i/{b}"] = i.numpy(); out[f"wf_j/{b}"] = j.numpy()
    # DETR matcher
    Q, K = 20, 7
    det = {"pred_logits": torch.from_numpy(rng.normal(size=(B, Q, K)).astype(np.float32)),
           "pred_boxes": torch.from_numpy(np.concatenate([rng.uniform(0.2, 0.8, (B, Q, 2)),
                                                          rng.uniform(0.05, 0.3, (B, Q, 2))], -1).astype(np.float32))}
    tsz = [4, 0, 9, 20, 2]
    dt = [{"labels": torch.from_numpy(rng.integers(0, K, (t,)).astype(np.int64)),
           "boxes": torch.from_numpy(np.concatenate([rng.uniform(0.2, 0.8, (t, 2)),
                                                     rng.uniform(0.05, 0.3, (t, 2))], -1).astype(np.float32))}
          for t in tsz]
    res = HungarianMatcher(cost_class=1.0, cost_bbox=5.0, cost_giou=2.0)(det, dt)
    out["detr_sizes"] = np.array(tsz, np.int64)
    for b, (i, j) in enumerate(res):
        out[f"detr_i/{b}"] = i.numpy(); out[f"detr_j/{b}"] = j.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "ok")


def case_loss_ties(name, seed):
    """Loss-style matching where many predictions are identical (constant dummy columns and
    duplicated rows -> heavy ties); pins the tie-breaking of the LSAP restatement."""
    rng = np.random.Generator(np.random.PCG64(seed))
    B, V = 6, 16
    pv = np.round(rng.uniform(-1, 1, (B, V, 3)) * 2) / 2            # coarse grid -> ties
    pe = np.round(rng.uniform(0, 1, (B, V)) * 4) / 4
    tv = np.zeros((B, V, 3)); counts = np.array([1, 2, 8, 16, 5, 12])
    for b in range(B):
        tv[b, :counts[b]] = np.round(rng.uniform(-1, 1, (counts[b], 3)) * 2) / 2
    pred = {"vertices": torch.from_numpy(pv.astype(np.float32)),
            "existence_probabilities": torch.from_numpy(pe.astype(np.float32))}
    tgt = {"vertices": torch.from_numpy(tv.astype(np.float32)),
           "vertex_counts": torch.from_numpy(counts.astype(np.int64))}
    res = WireframeLoss()._hungarian_matching(pred, tgt)
    out = {"meta": np.array([seed, B, V], np.int64), "counts": counts.astype(np.int64)}
    for b, (i, j) in enumerate(res):
        out[f"p/{b}"] = np.asarray(i, np.int64); out[f"t/{b}"] = np.asarray(j, np.int64)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "ok")


def case_hungarian_rmse(name, seed):
    """models/utils.py:38-55 (fp64 cdist + fp64 LSAP): random, rectangular both ways, and grid-snapped (tied) vertex sets."""
    from models.utils import hungarian_rmse
    rng = np.random.Generator(np.random.PCG64(seed))
    out = {"meta": np.array([seed], np.int64)}
    shapes = [(5, 5), (12, 7), (7, 12), (1, 4), (38, 38), (64, 20), (16, 16), (9, 9)]
    vals = []
    for k, (a, b) in enumerate(shapes):
        p = rng.uniform(-1, 1, (a, 3)); t = rng.uniform(-1, 1, (b, 3))
        if k >= 6:                                       # heavy ties: coordinates on a coarse grid
            p = np.round(p * 2) / 2; t = np.round(t * 2) / 2
        if k == 7:
            p = p.astype(np.float32); t = t.astype(np.float32)
        out[f"p/{k}"] = p; out[f"t/{k}"] = t
        vals.append(hungarian_rmse(p, t))
    out["rmse"] = np.asarray(vals, np.float64)
    out["empty"] = np.array([hungarian_rmse(np.zeros((0, 3)), np.zeros((0, 3))), hungarian_rmse(np.zeros((0, 3)), np.ones((2, 3)))])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, out["rmse"], out["empty"])


if __name__ == "__main__":
    case_train("train_b2_n384_v12", seed=3, B=2, N=384, V=12)
    case_train("train_b3_n300_v20_pad", seed=5, B=3, N=300, V=20, pad_frac=0.1)
    case_train("train_b1_n256_v8_rawint", seed=7, B=1, N=256, V=8, norm_intensity=False)
    case_eval("eval_b2_n256_v16", seed=11, B=2, N=256, V=16)
    case_matchers("matchers", seed=13)
    case_loss_ties("loss_ties", seed=17)
    # BASELINE.json's real shapes (configs[1]/[2]): 10,000-point clouds, 64 vertex slots, counts ~ U{16..64}
    case_train("train_b2_n10000_v64", seed=23, B=2, N=10000, V=64, min_count=16, max_count=64, store_dx=False)
    case_eval("eval_b2_n10000_v64", seed=29, B=2, N=10000, V=64, exist_bias=np.linspace(3.0, -3.0, 64))
    case_hungarian_rmse("hungarian_rmse", seed=31)


def case_real_building(name, seed, index=0):
    """BASELINE.json configs[0] / SURVEY 8d config 1: a REAL building from the shipped dataset, read by the reference's own
    loader (datasets/building3d.py: np.loadtxt, RGBA/256, centring + unit max-norm, 2560-point resampling; intensity stays
    un-normalised), targets built as train.py:48-88,112-115 does, one training step of the unmodified reference with
    the deterministic weights.  The loader's output (points, vertices, edges) is stored too -- it cannot regenerate from a seed."""
    import yaml
    from datasets import build_dataset
    from oracle import targets_oracle as to

    class Cfg(dict):
        __getattr__ = dict.__getitem__
    cfg = Cfg(yaml.safe_load(open(os.path.join(REF, "datasets", "dataset_config.yaml")))["Building3D"])
    cfg["root_dir"] = os.path.join(REF, "datasets")
    cfg["augment"] = False
    np.random.seed(seed)                                   # random_sampling draws with the global numpy RNG
    ds = build_dataset(cfg)["train"]
    ds.pc_files.sort(); ds.wireframe_files.sort()
    batch = ds.collate_batch([ds[index]])
    x = batch["point_clouds"].float()                      # (1, 2560, 8), train.py:48
    wf_v, wf_e = batch["wf_vertices"], batch["wf_edges"]
    V = 38                                                 # the dataset's largest vertex count
    from models.utils import create_edge_labels_from_edge_set
    tgt = to.prepare_targets(wf_v, wf_e, V, label_fn=create_edge_labels_from_edge_set)
    counts = tgt["vertex_counts"]
    m, _ = build_reference(seed, V, True)
    crit = WireframeLoss(vertex_weight=3.0, edge_weight=1.0, existence_weight=1.5)
    pred = m(x, counts)
    matched = crit._hungarian_matching(pred, tgt)
    ld = crit(pred, tgt)
    ld["total_loss"].backward()
    out = {
        "meta": np.array([seed, 1, x.shape[1], V], np.int64),
        "x": x.numpy(), "wf_vertices": wf_v[0].numpy(), "wf_edges": wf_e[0].numpy(),
        "tgt_vertices": tgt["vertices"].numpy(), "tgt_existence": tgt["vertex_existence"].numpy(),
        "tgt_edge_labels": tgt["edge_labels"].numpy(), "tgt_counts": counts.numpy(),
        "vertices": pred["vertices"].detach().numpy(), "existence": pred["existence_probabilities"].detach().numpy(),
        "edge_probs": pred["edge_probs"].detach().numpy(), "global_features": pred["global_features"].detach().numpy(),
        "losses": np.array([ld[k].item() for k in ("total_loss", "vertex_loss", "existence_loss", "edge_loss")], np.float64),
        "match_p": np.asarray(matched[0][0], np.int64), "match_t": np.asarray(matched[0][1], np.int64),
    }
    out.update(grad_digest(m.named_parameters()))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "points", tuple(x.shape), "vertices", int(counts[0]), "edges", len(wf_e[0]), "losses", out["losses"])


if __name__ == "__main__" and os.environ.get("WF_GOLDEN_REAL", "1") == "1":
    case_real_building("real_b1_n2560_v38", seed=19)
