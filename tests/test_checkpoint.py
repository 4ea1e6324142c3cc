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


def _ckpt_worker(rank, world, port, path, q):
    """Two ranks whose replicas differ (they should not -- this makes the direction of the copy visible): rank 0 writes, both
    wait at the barrier inside save_training_state, both load, and end up with rank 0's weights, Adam moments and step."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [os.path.join(root, "wireframe-3d-prediction_b200"), root]
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = _with_proj(10 + rank)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3, weight_decay=1e-6)
    g = torch.Generator().manual_seed(rank)
    for p_ in m.parameters():
        p_.grad = torch.randn(p_.shape, generator=g) * 1e-3
    opt.step()
    save_training_state(path, m, opt, step=3 + rank)               # rank 1's step / weights must NOT end up in the file
    wrote = os.path.exists(path) and not os.path.exists(path + ".tmp")
    m2 = PointCloudToWireframe(input_dim=8, max_vertices=8)
    load_model_state(m2, path)
    opt2 = torch.optim.Adam(m2.parameters(), lr=1e-3, weight_decay=1e-6)
    step = load_training_state(path, m2, opt2)
    flat = torch.cat([v.reshape(-1).float() for v in m2.state_dict().values()])
    mom = torch.cat([s["exp_avg"].reshape(-1) for s in opt2.state_dict()["state"].values()])
    both = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(both, flat)
    moms = [torch.zeros_like(mom) for _ in range(world)]
    dist.all_gather(moms, mom)
    ref0 = torch.cat([v.reshape(-1).float() for v in _with_proj(10).state_dict().values()])
    ok = wrote and step == 3 and torch.equal(both[0], both[1]) and torch.equal(moms[0], moms[1]) and not torch.equal(both[0], ref0)
    if rank == 0:                                                   # (weights moved one Adam step away from the seed-10 initial values)
        q.put(bool(ok))
    dist.destroy_process_group()


def test_training_state_is_written_by_rank0_and_restored_identically_on_all_ranks(tmp_path):
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    path = os.path.join(tmp_path, "dp_state.pt")
    procs = [ctx.Process(target=_ckpt_worker, args=(r, 2, port, path, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True
