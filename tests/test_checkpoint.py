"""Checkpoint fidelity / resume -- host logic only, runs on CPU (no kernels are launched)."""
import os

import torch

from models.PointCloudToWireframe import PointCloudToWireframe
from wf_b200.checkpoint import export_reference_checkpoint, load_model_state, load_training_state, save_training_state


def _with_proj(seed):
    torch.manual_seed(seed)
    m = PointCloudToWireframe(input_dim=8, max_vertices=8)
    m.vertex_predictor.point_pool_proj = torch.nn.Linear(1024, 512)      # what the first forward creates (VertexPredictor.py:94-97)
    return m


def test_reference_checkpoint_roundtrip_keeps_lazy_projection(tmp_path):
    src = _with_proj(1)
    p = os.path.join(tmp_path, "trained_model.pth")
    export_reference_checkpoint(p, src)
    sd = torch.load(p)
    assert len(sd) == 80 and sd["vertex_predictor.final_layer.weight"].shape[0] // 4 == 8      # evaluate.py:50-52
    fresh = PointCloudToWireframe(input_dim=8, max_vertices=8)
    # the reference's way: strict=False on a fresh model drops the projection ...
    res = fresh.load_state_dict(sd, strict=False)
    assert "vertex_predictor.point_pool_proj.weight" in res.unexpected_keys
    # ... the helper keeps it
    fresh2 = PointCloudToWireframe(input_dim=8, max_vertices=8)
    res2 = load_model_state(fresh2, p)
    assert not res2.unexpected_keys and not res2.missing_keys
    for (k, a), (_, b) in zip(src.state_dict().items(), fresh2.state_dict().items()):
        assert torch.equal(a, b), k


def test_training_state_resume(tmp_path):
    m = _with_proj(2)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3, weight_decay=1e-6)
    for p_ in m.parameters():
        p_.grad = torch.randn_like(p_) * 1e-3
    opt.step()
    path = os.path.join(tmp_path, "state.pt")
    save_training_state(path, m, opt, step=7)
    m2 = PointCloudToWireframe(input_dim=8, max_vertices=8)
    step = None
    load_model_state(m2, path)                                           # creates the projection, then the optimizer can see it
    opt2 = torch.optim.Adam(m2.parameters(), lr=1e-3, weight_decay=1e-6)
    step = load_training_state(path, m2, opt2)
    assert step == 7
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k
    s1, s2 = opt.state_dict()["state"], opt2.state_dict()["state"]
    assert s1.keys() == s2.keys()
    for k in s1:
        assert torch.equal(s1[k]["exp_avg"], s2[k]["exp_avg"]) and torch.equal(s1[k]["exp_avg_sq"], s2[k]["exp_avg_sq"])
