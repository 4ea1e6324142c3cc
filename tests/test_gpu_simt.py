"""GPU parity of the non-tensor-core kernels, each against torch fp32/fp64 on the same inputs, called
through the C ABI (wf_b200.ops -> libwf_b200.so).  Tolerances are scale-relative max errors."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from gpu_util import assert_close, col_to_pairs, scipy_pairs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    """These tests check the fp32 SIMT kernels: run them in the library's fp32 mode (in production precision big
    products are routed to TF32 tensor cores, which test_gpu_tc.py covers)."""
    from wf_b200 import ops as o
    o.set_precision("fp32")
    yield o
    o.set_precision("bf16")


def test_lsap_batched_matches_scipy(ops):
    rng = np.random.default_rng(0)
    cases = []
    for _ in range(1500):
        nr, nc = rng.integers(1, 65, size=2)
        kind = rng.integers(0, 4)
        if kind == 0:
            c = rng.uniform(0, 1, (nr, nc))
        elif kind == 1:
            c = rng.integers(0, 3, (nr, nc)).astype(np.float64)          # heavy ties
        elif kind == 2:
            c = np.ones((nr, nc)) * rng.integers(0, 2)                     # constant
        else:
            c = rng.normal(size=(nr, nc))
        cases.append(c.astype(np.float32))
    B = len(cases)
    cost = np.zeros((B, 64, 64), np.float32)
    for b, c in enumerate(cases):
        cost[b, :c.shape[0], :c.shape[1]] = c
    nr = torch.tensor([c.shape[0] for c in cases], dtype=torch.int32, device="cuda")
    nc = torch.tensor([c.shape[1] for c in cases], dtype=torch.int32, device="cuda")
    col, status = ops.lsap_batched(torch.from_numpy(cost).cuda(), nr, nc)
    col = col.cpu().numpy(); status = status.cpu().numpy()
    assert (status == 0).all()
    bad = 0
    for b, c in enumerate(cases):
        r0, c0 = scipy_pairs(c)
        r1, c1 = col_to_pairs(col[b], c.shape[0], c.shape[1])
        if not (np.array_equal(r0, r1) and np.array_equal(c0, c1)):
            bad += 1
    assert bad == 0, f"{bad}/{B} assignments differ from scipy"


def test_lsap_status_codes(ops):
    cost = torch.zeros(3, 4, 4)
    cost[0, :, 2:] = float("inf"); cost[0, :, :2] = 1.0          # infeasible: 4 rows, 2 finite columns
    cost[1, 1, 1] = float("nan")
    cost[2] = torch.arange(16.).reshape(4, 4)
    n = torch.full((3,), 4, dtype=torch.int32, device="cuda")
    col, status = ops.lsap_batched(cost.cuda(), n, n)
    assert status.tolist() == [1, 2, 0]
    r, c = scipy_pairs(cost[2].numpy())
    assert col[2].tolist() == c.tolist()


def test_loss_match_matches_oracle(ops):
    from oracle import wireframe_oracle as wo
    rng = np.random.default_rng(1)
    for V, quant in ((64, False), (38, False), (16, True), (5, True)):
        B = 96
        pv = rng.uniform(-1, 1, (B, V, 3)); pe = rng.uniform(0, 1, (B, V)); tv = rng.uniform(-1, 1, (B, V, 3))
        if quant:
            pv = np.round(pv * 2) / 2; pe = np.round(pe * 4) / 4; tv = np.round(tv * 2) / 2
        counts = rng.integers(0, V + 1, (B,))
        pvt, pet, tvt = (torch.from_numpy(a.astype(np.float32)) for a in (pv, pe, tv))
        ct = torch.from_numpy(counts.astype(np.int64))
        col, status, cost = ops.loss_match(pvt.cuda(), pet.cuda(), tvt.cuda(), ct.cuda(), want_cost=True)
        assert (status == 0).all()
        col = col.cpu().numpy(); cost = cost.cpu().numpy()
        for b in range(B):
            c = int(counts[b])
            ref_cost = wo.loss_cost_matrix(pvt[b], pet[b], tvt[b], c).numpy()
            assert np.array_equal(ref_cost, cost[b]), f"cost matrix differs V={V} b={b}"
            r0, c0 = scipy_pairs(ref_cost)
            assert np.array_equal(c0, col[b]), f"assignment differs V={V} b={b}"
    # count > V -> infeasible, as the reference's inf rows make scipy raise
    _, status, _ = ops.loss_match(pvt.cuda(), pet.cuda(), tvt.cuda(), torch.full((B,), V + 1).cuda())
    assert (status == 1).all()


@pytest.mark.parametrize("tA,tB", [(False, True), (False, False), (True, False), (True, True)])
def test_gemm_f32(ops, tA, tB):
    torch.manual_seed(0)
    for (M, N, K) in ((64, 4096, 512), (37, 129, 1031), (300, 8, 3), (1, 1, 128), (515, 64, 70)):
        A = torch.randn((K, M) if tA else (M, K), device="cuda")
        B = torch.randn((N, K) if tB else (K, N), device="cuda")
        bias = torch.randn(N, device="cuda")
        ref = (A.t() if tA else A).double() @ (B.t() if tB else B).double() + bias.double()
        got = ops.gemm_f32(A, B, transA=tA, transB=tB, bias=bias)
        assert_close(got, ref, 2e-6 * math.sqrt(K) + 1e-6, f"gemm_f32 {M}x{N}x{K}")
    # strided views + accumulate
    W = torch.randn(512, 1031, device="cuda"); x = torch.randn(77, 512, device="cuda")
    out = torch.randn(77, 512, device="cuda"); ref = out.double() + x.double() @ W[:, 512:1024].double().t() * 0.5
    ops.gemm_f32(x, W[:, 512:1024], transB=True, out=out, beta=1.0, alpha=0.5)
    assert_close(out, ref, 1e-5, "gemm_f32 strided accumulate")


@pytest.mark.parametrize("act", [0, 1, 2])
@pytest.mark.parametrize("C", [256, 512, 1024, 2048, 4096])
def test_linear_ln_act_fwd_bwd(ops, act, C):
    torch.manual_seed(C + act)
    M, K = 45, 96
    x = torch.randn(M, K, device="cuda", dtype=torch.float64, requires_grad=True)
    W = (torch.randn(C, K, device="cuda", dtype=torch.float64) / math.sqrt(K)).requires_grad_(True)
    b = torch.randn(C, device="cuda", dtype=torch.float64, requires_grad=True)
    g = (1 + 0.1 * torch.randn(C, device="cuda", dtype=torch.float64)).requires_grad_(True)
    be = (0.1 * torch.randn(C, device="cuda", dtype=torch.float64)).requires_grad_(True)
    res = torch.randn(M, C, device="cuda", dtype=torch.float64, requires_grad=True)
    keep = (torch.rand(M, C, device="cuda") > 0.1).to(torch.uint8)
    actf = [lambda t: t, torch.relu, F.gelu][act]
    ref = actf(F.layer_norm(F.linear(x, W, b), (C,), g, be, 1e-5)) * keep.double() / 0.9 + res
    go = torch.randn_like(ref)
    ref.backward(go)
    leaves = [x, W, b, g, be, res]
    f32 = [t.detach().float().requires_grad_(True) for t in leaves]
    out = ops.linear_ln_act(f32[0], f32[1], f32[2], f32[3], f32[4], act, f32[5], keep, 1.0 / 0.9)
    out.backward(go.float())
    assert_close(out, ref, 2e-5, "fwd")
    for n, a, r in zip("x W b gamma beta res".split(), f32, leaves):
        assert_close(a.grad, r.grad, 5e-5, f"grad {n} act={act} C={C}")


def test_plain_linear_and_noln_gelu(ops):
    torch.manual_seed(3)
    x = torch.randn(33, 256, device="cuda", dtype=torch.float64, requires_grad=True)
    W = (torch.randn(128, 256, device="cuda", dtype=torch.float64) / 16).requires_grad_(True)
    b = torch.randn(128, device="cuda", dtype=torch.float64, requires_grad=True)
    for act, fn in ((0, lambda t: t), (2, F.gelu)):
        for t in (x, W, b):
            t.grad = None
        ref = fn(F.linear(x, W, b)); go = torch.randn_like(ref); ref.backward(go)
        f = [t.detach().float().requires_grad_(True) for t in (x, W, b)]
        out = ops.linear_ln_act(f[0], f[1], f[2], None, None, act)
        out.backward(go.float())
        assert_close(out, ref, 1e-5, "fwd")
        for a, r in zip(f, (x, W, b)):
            assert_close(a.grad, r.grad, 2e-5, "grad")


def test_pool_fwd_bwd(ops):
    torch.manual_seed(4)
    B, N, C = 3, 517, 96
    pf = torch.randn(B, N, C, device="cuda")
    pf[0, 5] = pf[0, 2]                       # exact duplicates -> first index must win
    pf[1, :, 7] = 1.25                        # constant column -> argmax 0
    mask = (torch.rand(B, N, device="cuda") > 0.2)
    mask[2] = False                           # fully masked sample -> max 0, no gradient through max
    valid = mask.sum(1).clamp(min=1).float()
    pfd = pf.double().requires_grad_(True)
    avg = (pfd * mask.unsqueeze(-1)).sum(1) / valid.double().unsqueeze(1)
    mm, am = pfd.masked_fill(~mask.unsqueeze(-1), float("-inf")).max(1)
    mm = torch.where(torch.isfinite(mm), mm, torch.zeros_like(mm))
    mu, au = pfd.max(1)
    mean = pfd.mean(1)
    gs = [torch.randn(B, C, device="cuda", dtype=torch.float64) for _ in range(4)]
    (mm * gs[0] + avg * gs[1] + mu * gs[2] + mean * gs[3]).sum().backward()
    p32 = pf.clone().requires_grad_(True)
    r = ops.PoolPoints.apply(p32, mask.to(torch.uint8), valid)
    assert_close(r[0], mm, 1e-6, "max_m"); assert_close(r[1], avg, 1e-5, "avg_m")
    assert_close(r[2], mu, 1e-6, "max_u"); assert_close(r[3], mean, 1e-5, "mean_u")
    assert torch.equal(r[5].long(), au), "unmasked argmax"
    assert torch.equal(r[4][:2].long(), am[:2]), "masked argmax"
    assert (r[4][2] == -1).all()
    (r[0] * gs[0].float() + r[1] * gs[1].float() + r[2] * gs[2].float() + r[3] * gs[3].float()).sum().backward()
    assert_close(p32.grad, pfd.grad, 1e-5, "d_pf")


def test_attention_matches_torch_mha(ops):
    torch.manual_seed(5)
    counts = [2, 64, 17, 33, 5]
    rg = ops.Ragged(counts, "cuda")
    mha = torch.nn.MultiheadAttention(512, 8, dropout=0.0, batch_first=True).cuda().double()
    f = torch.randn(rg.T, 512, device="cuda", dtype=torch.float64, requires_grad=True)
    outs = []
    off = 0
    for c in counts:
        o, _ = mha(f[off:off + c][None], f[off:off + c][None], f[off:off + c][None])
        outs.append(o[0]); off += c
    ref = torch.cat(outs); go = torch.randn_like(ref); ref.backward(go)
    f32 = f.detach().float().requires_grad_(True)
    Wi = mha.in_proj_weight.detach().float().requires_grad_(True); bi = mha.in_proj_bias.detach().float().requires_grad_(True)
    Wo = mha.out_proj.weight.detach().float().requires_grad_(True); bo = mha.out_proj.bias.detach().float().requires_grad_(True)
    qkv = ops.linear_ln_act(f32, Wi, bi)
    o = ops.AttentionCore.apply(qkv, rg, None, 1.0)
    out = ops.linear_ln_act(o, Wo, bo)
    out.backward(go.float())
    assert_close(out, ref, 2e-5, "mha fwd")
    assert_close(f32.grad, f.grad, 5e-5, "mha d_input")
    assert_close(Wi.grad, mha.in_proj_weight.grad, 5e-5, "mha d_in_proj")
    assert_close(Wo.grad, mha.out_proj.weight.grad, 5e-5, "mha d_out_proj")


def test_edge_pair_and_out(ops):
    torch.manual_seed(6)
    counts = [3, 9, 2, 30]
    rg = ops.Ragged(counts, "cuda")
    C = 512
    mk = lambda *s: torch.randn(*s, device="cuda", dtype=torch.float64, requires_grad=True)
    P, Q, verts, wd, bias = mk(rg.T, C), mk(rg.T, C), mk(rg.T, 3), mk(C), mk(C)
    w_out, b_out = mk(1, 128), mk(1)
    Wm = mk(128, C)
    zs = []
    off = 0
    for c in counts:
        iu = torch.triu_indices(c, c, offset=1, device="cuda")
        i, j = iu[0] + off, iu[1] + off
        d = torch.norm(verts[i] - verts[j], dim=-1, keepdim=True)
        zs.append(P[i] + Q[j] + d * wd + bias); off += c
    z_ref = torch.cat(zs)
    h_ref = F.gelu(z_ref @ Wm.t())
    logits = (h_ref @ w_out.t() + b_out).reshape(-1)
    pr = torch.sigmoid(logits)
    padded = torch.zeros(len(counts), rg.max_e, device="cuda", dtype=torch.float64)
    eo = 0
    rows = []
    for b, c in enumerate(counts):
        e = c * (c - 1) // 2
        rows.append(F.pad(pr[eo:eo + e], (0, rg.max_e - e))); eo += e
    padded = torch.stack(rows)
    go = torch.randn_like(padded); padded.backward(go)
    leaves = [P, Q, verts, wd, bias, Wm, w_out, b_out]
    f = [t.detach().float().requires_grad_(True) for t in leaves]
    pq = torch.cat([f[0], f[1]], dim=1)                     # the layer takes the stacked [P | Q] product
    z = ops.EdgePairLayer.apply(pq, f[2], f[3], f[4], rg)
    assert_close(z, z_ref, 1e-5, "pair fwd")
    h = ops.linear_ln_act(z, f[5], None, None, None, 2)
    out = ops.EdgeOut.apply(h, f[6], f[7], rg)
    assert_close(out, padded, 1e-4, "edge out fwd")
    out.backward(go.float())
    for n, a, r in zip("P Q verts wd bias Wm w_out b_out".split(), f, leaves):
        assert_close(a.grad, r.grad, 3e-4, f"edge grad {n}")


@pytest.mark.parametrize("hidden,heads", [(256, 4), (128, 8), (1024, 8), (512, 4)])
def test_edge_predictor_other_widths(ops, hidden, heads):
    """models/EdgePredictor.py:19 accepts hidden_dim / num_heads; the drop-in runs the same kernels instantiated for other widths.
    Held to a float64 restatement of the reference forward (models/EdgePredictor.py:91-140) on the same parameters, forward
    and every parameter gradient."""
    from models.EdgePredictor import EdgePredictor
    torch.manual_seed(8)
    m = EdgePredictor(vertex_dim=3, hidden_dim=hidden, num_heads=heads).cuda().eval()   # eval: the three dropout sites off
    B, V = 3, 11
    verts = torch.rand(B, V, 3, device="cuda")
    vg = verts.clone().requires_grad_(True)
    probs, idx = m(vg)
    go = torch.randn_like(probs)
    probs.backward(go)
    got = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    m.zero_grad()
    md = EdgePredictor.__new__(EdgePredictor)
    torch.nn.Module.__init__(md)
    import copy
    for name in ("vertex_proj", "attention", "spatial_proj", "edge_mlp"):
        setattr(md, name, copy.deepcopy(getattr(m, name)).double())
    vd = verts.double().requires_grad_(True)
    f = md.vertex_proj(vd)
    a, _ = md.attention(f, f, f)
    f = f + a
    iu = torch.triu_indices(V, V, offset=1, device="cuda")
    v1, v2 = vd[:, iu[0]], vd[:, iu[1]]
    feat = torch.cat([f[:, iu[0]], f[:, iu[1]], v1, v2, torch.norm(v1 - v2, dim=-1, keepdim=True)], dim=-1)
    ref = torch.sigmoid(md.edge_mlp(feat.view(-1, feat.shape[-1]))).view(B, -1)
    ref.backward(go.double())
    assert idx == iu.t().tolist()
    assert_close(probs, ref, 2e-5, "edge probs")
    assert_close(vg.grad, vd.grad, 2e-4, "d vertices")
    for k, p in md.named_parameters():
        if k.startswith("spatial_proj"):
            assert k not in got and p.grad is None          # unused in the reference as well
            continue
        assert_close(got[k], p.grad, 3e-4, f"grad {k}")


def test_edge_predictor_unsupported_dims(ops):
    from models.EdgePredictor import EdgePredictor
    with pytest.raises(AssertionError):                      # nn.MultiheadAttention: embed_dim must be divisible by num_heads
        EdgePredictor(hidden_dim=512, num_heads=7)
    with pytest.raises(ops._lib.WfError):
        EdgePredictor(vertex_dim=2)


def test_gather_prefix_and_vertex_split(ops):
    torch.manual_seed(7)
    B, V = 4, 10
    vf = torch.randn(B, 4 * V, device="cuda", requires_grad=True)
    coords, prob, count = ops.VertexSplit.apply(vf, V)
    ref = vf.detach().double().reshape(B, V, 4)
    assert_close(coords, ref[:, :, :3], 1e-7, "coords"); assert_close(prob, torch.sigmoid(ref[:, :, 3]), 1e-6, "prob")
    assert torch.equal(count, (torch.sigmoid(ref[:, :, 3]).float() > 0.5).sum(1))
    rg = ops.Ragged([2, 10, 5, 7], "cuda")
    packed = ops.GatherPrefix.apply(coords, rg)
    refp = torch.cat([coords[b, :c] for b, c in enumerate(rg.counts)])
    assert torch.equal(packed, refp)
    gp = torch.randn_like(packed); gq = torch.randn_like(prob)
    (packed * gp).sum().backward(retain_graph=True)
    vd = vf.detach().double().requires_grad_(True)
    r4 = vd.reshape(B, V, 4)
    (torch.cat([r4[b, :c, :3] for b, c in enumerate(rg.counts)]) * gp.double()).sum().backward()
    assert_close(vf.grad, vd.grad, 1e-6, "gather/split bwd")


def test_loss_kernel_matches_oracle(ops):
    from oracle import wireframe_oracle as wo
    rng = np.random.default_rng(8)
    B, V = 5, 14
    counts = torch.tensor([2, 14, 7, 3, 9])
    Ep, El = 91, 80
    pv = torch.from_numpy(rng.uniform(-2, 2, (B, V, 3)).astype(np.float32))
    pe = torch.from_numpy(rng.uniform(0.01, 0.99, (B, V)).astype(np.float32))
    ep = torch.from_numpy(rng.uniform(0.0, 1.0, (B, Ep)).astype(np.float32)); ep[0, :3] = 0.0; ep[1, :2] = 1.0
    tv = torch.from_numpy(rng.uniform(-1, 1, (B, V, 3)).astype(np.float32))
    te = (torch.arange(V)[None] < counts[:, None]).float()
    el = torch.from_numpy((rng.uniform(size=(B, El)) < 0.3).astype(np.float32))
    pred = {"vertices": pv.clone().requires_grad_(True), "existence_probabilities": pe.clone().requires_grad_(True),
            "edge_probs": ep.clone().requires_grad_(True)}
    tgt = {"vertices": tv, "vertex_existence": te, "edge_labels": el, "vertex_counts": counts}
    ref = wo.loss_forward(pred, tgt, 3.0, 1.0, 1.5)
    ref["total_loss"].backward()
    from losses.WireframeLoss import WireframeLoss
    crit = WireframeLoss(3.0, 1.0, 1.5)
    predg = {k: v.detach().cuda().requires_grad_(True) for k, v in pred.items()}
    tgtg = {k: v.cuda() for k, v in tgt.items()}
    out = crit(predg, tgtg)
    for k in ("total_loss", "vertex_loss", "existence_loss", "edge_loss"):
        assert abs(out[k].item() - ref[k].item()) <= 2e-5 * max(1.0, abs(ref[k].item())), k
    out["total_loss"].backward()
    for k in pred:
        assert_close(predg[k].grad, pred[k].grad, 2e-5, f"loss grad {k}")
    m_ref = wo.loss_matching(pred, tgt); m_got = crit._hungarian_matching(predg, tgtg)
    for (a, b), (c, d) in zip(m_ref, m_got):
        assert np.array_equal(a, c) and np.array_equal(b, d)
