"""End-to-end GPU parity through the drop-in module API (models.PointCloudToWireframe, losses.WireframeLoss,
matchers) against the golden fixtures generated from the unmodified reference and against the oracle.

fp32 mode: BASELINE.json's 1e-5-relative class -- outputs and losses 2e-5, gradient norms 5e-5, gradient entries 2e-4
(scale-relative; the reference's own run-to-run reproducibility on its CPU path is 1.4e-5 on a gradient entry); matchings,
counts and decided argmax indices identical.  bf16 mode (production): outputs 1e-2, gradients element-wise 5e-2 with the
reference's assignment injected.  Observed values: profiles/r02_parity_observed.jsonl."""
import os

import numpy as np
import pytest
import torch

from gpu_util import assert_close, rel_err

pytestmark = pytest.mark.gpu

TRAIN = ["train_b2_n384_v12", "train_b3_n300_v20_pad", "train_b1_n256_v8_rawint", "train_b2_n10000_v64"]


def _model(seed, V, train):
    from oracle import wireframe_oracle as wo
    from models.PointCloudToWireframe import PointCloudToWireframe
    os.environ["WF_B200_EAGER_POOL_PROJ"] = "1"
    m = PointCloudToWireframe(input_dim=8, max_vertices=V)
    os.environ["WF_B200_EAGER_POOL_PROJ"] = "0"
    m.load_state_dict(wo.make_state_dict(seed, V), strict=True)
    m = m.cuda()
    m.train(train)
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.eval()
    m.edge_predictor.attention.dropout = 0.0
    return m


# Tolerances (scale-relative max error unless noted).  Each is <= 10x the error observed on a B200 (profiles/r02_parity_observed.jsonl,
# written by gpu_util.record) and the fp32 gradient ones cannot go below the reference's OWN reproducibility: two runs of the
# unmodified reference on this container's 8 threads differ by up to 1.4e-5 on a gradient entry and 1e-5 on a gradient norm
# (MKL summation order; measured while regenerating the fixtures).
# bf16 mode against the fp32 REFERENCE: outputs 1e-2 (observed 5e-3).  Its gradients are compared by norm (3e-2, observed
# 2.4e-2) and by the direction of the whole gradient (cosine >= 0.99): entry by entry they differ from the fp32 reference by
# up to ~0.4 of the largest entry EVEN THOUGH every kernel is exact to rounding, because a forward perturbation of 3e-3 flips
# the ReLU mask of the few units that sit within 3e-3 of zero, and at batch size 2 a head's weight-gradient row is a sum over
# TWO samples (tools/diag_bf16_grads.py: same level with the reference's matching AND argmax injected, TF32 heads on or off;
# fp32 mode: 2e-6).  The entry-by-entry check of the bf16 mode is therefore made against the oracle evaluated WITH the bf16
# mode's roundings inserted (test_bf16_mode_vs_bf16_emulating_oracle below), where masks, argmax and matching coincide.
TOL = {
    "fp32": dict(out=2e-5, loss=2e-5, gnorm=5e-5, gelem=2e-4, dx=2e-4, cost=1e-5),
    "bf16": dict(out=1e-2, loss=5e-3, gnorm=3e-2, gelem=None, dx=None, cost=5e-3, cosine=0.99),
}


def _ref_col(g, B, V):
    """The reference's assignment as the loss kernels take it: col_of_row[b, pred slot] = target index, -1 = unmatched."""
    col = torch.full((B, V), -1, dtype=torch.int32)
    for b in range(B):
        col[b, torch.from_numpy(g[f"match_p/{b}"])] = torch.from_numpy(g[f"match_t/{b}"]).to(torch.int32)
    return col


def _assignment_cost(cost_b, pi, ti, count):
    """Total cost of a filtered assignment on one sample's (V, V) loss cost matrix: matched real columns + the constant
    dummy-column cost of every unmatched prediction row (losses/WireframeLoss.py:216-219)."""
    V = cost_b.shape[0]
    tot = float(cost_b[pi, ti].double().sum())
    un = np.setdiff1d(np.arange(V), pi)
    if len(un) and count < V:
        tot += float(cost_b[un, count].double().sum())
    return tot


def _check_own_matching(crit, pred, tg, g, B, tol_cost):
    """Our assignment equals the reference's, or -- where near-tied costs make the optimum itself ill-defined -- has the
    same total cost on OUR cost matrix to within rounding.  Returns the number of samples with identical pairs."""
    from wf_b200 import ops
    ours = crit._hungarian_matching(pred, tg)
    _, _, cost = ops.loss_match(pred["vertices"], pred["existence_probabilities"], tg["vertices"], tg["vertex_counts"],
                                want_cost=True)
    cost = cost.cpu()
    same = 0
    for b, (pi, ti) in enumerate(ours):
        rp, rt = g[f"match_p/{b}"], g[f"match_t/{b}"]
        if np.array_equal(pi, rp) and np.array_equal(ti, rt):
            same += 1
            continue
        cnt = int(tg["vertex_counts"][b])
        c_own, c_ref = _assignment_cost(cost[b], pi, ti, cnt), _assignment_cost(cost[b], rp, rt, cnt)
        assert c_own <= c_ref * (1 + 1e-6) + 1e-6, f"sample {b}: our assignment costs more than the reference's on our own matrix"
        assert c_ref - c_own <= tol_cost * abs(c_ref), f"sample {b}: assignments differ and are not tied ({c_own} vs {c_ref})"
    return same


@pytest.mark.parametrize("name", TRAIN)
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_train_step_vs_reference_golden(golden_dir, name, prec):
    """One training step through the drop-in modules against the UNMODIFIED reference's outputs, losses, matching, argmax and
    all 80 parameter gradients (norm + 16 leading + <=1024 strided entries each) -- incl. BASELINE.json's real shape
    (2 clouds x 10,000 points, 64 vertex slots, counts ~ U{16..64}).
    Matching: identical pairs, or an assignment of equal cost (to rounding) where the reference's optimum is tied; the loss
    and the gradients are then taken with the REFERENCE's assignment injected, so that a flipped assignment cannot hide (or
    fake) a gradient error.  Argmax (fp32): identical wherever the reference's top-2 gap exceeds 1e-5*|max|."""
    from oracle import wireframe_oracle as wo
    from wf_b200 import ops
    from losses.WireframeLoss import WireframeLoss
    from gpu_util import record
    g = dict(np.load(os.path.join(golden_dir, name + ".npz")))
    seed, B, N, V, pad, norm_i = [int(v) for v in g["meta"]]
    cmin, cmax = [int(v) for v in g["count_range"]]
    tol = TOL[prec]
    rawint = "rawint" in name
    ops.set_precision(prec)
    try:
        m = _model(seed, V, True)
        x, tgt, counts = wo.make_inputs(seed, B, N, V, pad_frac=pad / 1000.0, norm_intensity=bool(norm_i), min_count=cmin,
                                        max_count=None if cmax < 0 else cmax)
        xg = x.cuda().requires_grad_(True)
        tg = {k: v.cuda() for k, v in tgt.items()}
        pred = m(xg, counts.cuda())
        crit = WireframeLoss(vertex_weight=3.0, edge_weight=1.0, existence_weight=1.5)
        same_pairs = _check_own_matching(crit, pred, tg, g, B, tol["cost"])
        ref_col = _ref_col(g, B, V).cuda()
        crit._match_device = lambda predictions, targets, sync=None: ref_col          # the reference's assignment, injected
        ld = crit(pred, tg)
        ld["total_loss"].backward()
        obs = {"matching_identical_samples": same_pairs, "samples": B}
        obs["vertices"] = assert_close(pred["vertices"], torch.from_numpy(g["vertices"]), tol["out"], "vertices")
        obs["existence"] = assert_close(pred["existence_probabilities"], torch.from_numpy(g["existence"]), tol["out"], "existence")
        obs["edge_probs"] = assert_close(pred["edge_probs"], torch.from_numpy(g["edge_probs"]), tol["out"], "edge_probs")
        obs["global_features"] = assert_close(pred["global_features"], torch.from_numpy(g["global_features"]), tol["out"],
                                              "global_features")
        assert [len(e) for e in pred["edge_indices"]] == g["n_edges"].tolist()
        got = np.array([ld[k].item() for k in ("total_loss", "vertex_loss", "existence_loss", "edge_loss")])
        obs["loss"] = float(np.max(np.abs(got - g["losses"]) / np.abs(g["losses"])))
        # raw intensity (~5e4): the bf16 rounding of the first layer's output is amplified by its ill-conditioned LayerNorm
        np.testing.assert_allclose(got, g["losses"], rtol=tol["loss"] * (4 if rawint and prec == "bf16" else 1), atol=0)
        if prec == "fp32":
            assert np.array_equal(pred["actual_vertex_counts"].cpu().numpy(), g["dyn_counts"])
        worst_n, worst_e = ("", 0.0), ("", 0.0)
        fails = []
        dot = n_our = n_ref = 0.0
        for k, p in m.named_parameters():
            if "gnone/" + k in g:
                assert p.grad is None, k
                continue
            gr = p.grad.detach().double().reshape(-1).cpu()
            ref_norm = float(g["gnorm/" + k][0])
            e = abs(float(gr.norm()) - ref_norm) / max(ref_norm, 1e-12)
            samp = torch.from_numpy(g["gsamp/" + k]).double()
            stride = max(1, -(-gr.numel() // 1024))
            # element-wise on the strided sample; the scale is the parameter's RMS gradient entry or the sample's largest
            scale = max(float(samp.abs().max()), ref_norm / np.sqrt(gr.numel()))
            es = float((gr[::stride] - samp).abs().max()) / scale
            l1 = k.startswith("encoder.mlp.0.") or k.startswith("encoder.mlp.1.")
            if not (rawint and l1):
                dot += float(gr[::stride] @ samp) * stride; n_our += float(gr[::stride].norm() ** 2) * stride; n_ref += float(samp.norm() ** 2) * stride
            if e > worst_n[1]:
                worst_n = (k, e)
            if es > worst_e[1]:
                worst_e = (k, es)
            # un-normalised intensity (~5e4) makes the first layer's LayerNorm backward ill-conditioned
            # in fp32 (rstd ~ 1e-4, heavy cancellation): the reference's own fp32 gradient is only ~1e-2 accurate there
            loose = (600.0 if l1 else 20.0) if (rawint and prec == "fp32") else (4.0 if rawint and l1 else 1.0)
            if e > tol["gnorm"] * loose or (tol["gelem"] is not None and es > tol["gelem"] * loose):
                fails.append((k, f"norm {e:.2e}", f"elem {es:.2e}"))
        obs["grad_norm_worst"], obs["grad_norm_worst_param"] = worst_n[1], worst_n[0]
        obs["grad_elem_worst"], obs["grad_elem_worst_param"] = worst_e[1], worst_e[0]
        obs["grad_cosine_sampled"] = dot / max(1e-300, np.sqrt(n_our * n_ref))
        if prec == "bf16":
            # un-normalised intensity, ONE cloud of 256 points: the intensity column dominates every normalised row, the rows are
            # nearly alike and NO pooled maximum has a decided top-2 gap (argmax_decided_fraction 0.0 -- even the fp32 mode
            # agrees with the reference on only 93 % of the argmax indices there), so bf16-sized noise re-routes the pooled
            # gradient to other points: observed cosine 0.9889 (0.996 on the other cases), asserted >= 0.98
            cos_tol = 0.98 if rawint else tol["cosine"]
            assert obs["grad_cosine_sampled"] >= cos_tol, f"gradient direction: cosine {obs['grad_cosine_sampled']:.4f}"
        # d/d(input) is a 512-term sum with LayerNorm cancellation: not meaningful under bf16 noise, nor in fp32 with
        # un-normalised intensity (the reference's own fp32 value is noise-dominated there)
        if prec == "fp32" and not rawint:
            # zero-padded points are exact duplicates: which duplicate the max-pool gradient lands on depends on last-bit
            # rounding inside the reference's MKL GEMM (all parameter gradients are unaffected) -> compare real points only
            dx_ref = torch.from_numpy(g["dx"])
            nrow = dx_ref.shape[1]                                 # large clouds store the first 64 points of each cloud
            real = (x[:, :nrow].abs().sum(-1) > 0)
            obs["dx"] = assert_close(xg.grad.cpu()[:, :nrow][real], dx_ref[real], tol["dx"], "dx")
        if prec == "fp32":        # argmax parity: identical wherever the reference's top-2 gap is above rounding noise
            r = m.encoder.pooled(xg.detach())
            # raw intensity (~5e4) leaves ~1e-3 relative fp32 noise on the point features (in the reference too)
            thr = (2e-3 if rawint else 1e-5) * np.abs(g["pf_max"])
            decided = g["pf_top2_gap"] > thr
            same = r[5].cpu().numpy() == g["pf_argmax"]
            obs["argmax_decided_fraction"] = float(decided.mean())
            obs["argmax_agree_undecided"] = float(same[~decided].mean()) if (~decided).any() else 1.0
            assert same[decided].all(), f"argmax differs on {int((~same[decided]).sum())} decided (cloud, channel) pairs"
            obs["pf_max"] = assert_close(r[2], torch.from_numpy(g["pf_max"]), tol["out"] * (100 if rawint else 1), "pooled max")
        record(f"train_step/{name}/{prec}", **obs)
        print(prec, name, obs)
        assert not fails, f"{prec} gradient mismatches: {fails[:6]} (+{max(0, len(fails) - 6)} more)"
    finally:
        ops.set_precision("bf16")


@pytest.mark.parametrize("B,N,V,cmin,cmax,seed", [(2, 384, 12, 2, 12, 3), (4, 1500, 32, 8, 32, 41), (2, 10000, 64, 16, 64, 23)])
def test_bf16_mode_vs_bf16_emulating_oracle(B, N, V, cmin, cmax, seed):
    """The production (bf16) mode entry by entry.  The oracle is evaluated WITH the bf16 mode's roundings inserted where the
    kernels round (oracle.encoder_point_features(emulate_bf16=True): bf16 h1..h4 / z2..z4 / weights of layers 2-5, fp32
    statistics and accumulation, mean pools through the affine map; straight-through gradients).  It then predicts the bf16
    mode's forward to ~1e-5 -- ReLU masks, argmax and matching coincide -- and what remains in the gradients is the bf16
    rounding of the stored gradient tensors (dz, dh: 2^-9 per stage), which is smooth.  The heads' products run in fp32
    here (the TF32 tensor-core path is compared with the fp32 path in tests/test_gpu_tc.py); the assignment and the argmax
    indices of the oracle are injected, so that the comparison cannot be spoilt by a tie."""
    from oracle import wireframe_oracle as wo
    from wf_b200 import ops
    from losses.WireframeLoss import WireframeLoss
    from gpu_util import record
    # The oracle's roundings and the kernels' cannot coincide exactly: the fp32 values that get rounded differ by ~1e-6
    # (summation order), which moves ~5e-4 of them across a bf16 rounding boundary -> a residual of ~8 % of the plain bf16
    # noise: outputs agree to ~3e-4 instead of ~3e-3 (asserted: 2e-3).
    OUT_EMU = 2e-3
    ops.set_precision("bf16")
    tf32 = ops.USE_TF32_HEADS
    ops.USE_TF32_HEADS = False
    try:
        m = _model(seed, V, True)
        x, tgt, counts = wo.make_inputs(seed, B, N, V, norm_intensity=True, min_count=cmin, max_count=cmax)
        sd = wo.make_state_dict(seed, V)
        sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        pred_o = wo.model_forward(sdr, x, counts, training=True, max_vertices=V, emulate_bf16=True)
        matches = wo.loss_matching(pred_o, tgt)
        ld_o = wo.loss_forward(pred_o, tgt, 3.0, 1.0, 1.5, matches=matches)
        ld_o["total_loss"].backward()
        with torch.no_grad():
            pf = wo.encoder_point_features(sd, x, True)
            _, _, _, arg_m = wo.encoder_pools(x, pf)
            arg_u = pf.max(dim=1).indices
        col = torch.full((B, V), -1, dtype=torch.int32)
        for b, (pi, ti) in enumerate(matches):
            col[b, torch.as_tensor(pi)] = torch.as_tensor(ti).to(torch.int32)
        col = col.cuda()
        ops.ARGMAX_OVERRIDE = (arg_m.to(torch.int32).cuda(), arg_u.to(torch.int32).cuda())
        pred = m(x.cuda(), counts.cuda())
        ops.ARGMAX_OVERRIDE = None
        crit = WireframeLoss(vertex_weight=3.0, edge_weight=1.0, existence_weight=1.5)
        own = crit._hungarian_matching(pred, {k: v.cuda() for k, v in tgt.items()})
        same = sum(int(np.array_equal(a, c) and np.array_equal(b_, d)) for (a, b_), (c, d) in zip(own, matches))
        crit._match_device = lambda predictions, targets, sync=None: col
        ld = crit(pred, {k: v.cuda() for k, v in tgt.items()})
        ld["total_loss"].backward()
        r = m.encoder.pooled(x.cuda())
        obs = {"matching_identical_samples": same, "samples": B,
               "argmax_agreement": float((r[5].cpu() == arg_u).float().mean()),
               "vertices": assert_close(pred["vertices"], pred_o["vertices"], OUT_EMU, "vertices"),
               "existence": assert_close(pred["existence_probabilities"], pred_o["existence_probabilities"], OUT_EMU, "existence"),
               "edge_probs": assert_close(pred["edge_probs"], pred_o["edge_probs"], OUT_EMU, "edge_probs"),
               "global_features": assert_close(pred["global_features"], pred_o["global_features"], OUT_EMU, "global_features"),
               "loss": abs(ld["total_loss"].item() - ld_o["total_loss"].item()) / abs(ld_o["total_loss"].item())}
        assert obs["loss"] < 5e-4
        worst_f, worst_e, fails = ("", 0.0), ("", 0.0), []
        dot = n_a = n_b = 0.0
        for k, p in m.named_parameters():
            go = sdr[k].grad
            if go is None or float(go.abs().max()) == 0.0:
                assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
                continue
            a, b_ = p.grad.detach().double().cpu().reshape(-1), go.double().reshape(-1)
            fro = float((a - b_).norm() / b_.norm())
            scale = max(float(b_.abs().max()), float(b_.norm()) / np.sqrt(b_.numel()))
            el = float((a - b_).abs().max()) / scale
            if fro > worst_f[1]:
                worst_f = (k, fro)
            if el > worst_e[1]:
                worst_e = (k, el)
            dot += float(a @ b_); n_a += float(a @ a); n_b += float(b_ @ b_)
            # Frobenius-relative per parameter: observed <= 3.6e-2 (heads: the few ReLU units that still flip at batch 2-4;
            # encoder: inherits the pooled gradient's error uniformly).  A single flipped unit moves single entries of a
            # head's bias gradient by up to 0.3 of the largest entry, so entries are not asserted one by one here -- the
            # encoder's backward is (test_bf16_encoder_backward_vs_bf16_emulating_oracle).
            if fro > 1e-1:
                fails.append((k, f"fro {fro:.2e}", f"elem {el:.2e}"))
        obs.update(grad_fro_worst=worst_f[1], grad_fro_worst_param=worst_f[0], grad_elem_worst=worst_e[1],
                   grad_elem_worst_param=worst_e[0], grad_cosine=dot / np.sqrt(n_a * n_b))
        assert obs["grad_cosine"] >= 0.999, f"direction of the whole gradient: cosine {obs['grad_cosine']:.5f}"
        record(f"bf16_emulating_oracle/B{B}_N{N}_V{V}", **obs)
        print(obs)
        assert not fails, f"bf16-mode gradients vs the bf16-emulating oracle: {fails[:6]} (+{max(0, len(fails) - 6)} more)"
    finally:
        ops.USE_TF32_HEADS = tf32
        ops.ARGMAX_OVERRIDE = None
        ops.set_precision("bf16")


@pytest.mark.parametrize("B,N", [(2, 700), (3, 4000)])
def test_bf16_encoder_backward_vs_bf16_emulating_oracle(B, N):
    """The tensor-core encoder's backward ENTRY BY ENTRY: all 18 parameter gradients of the per-point MLP for fixed upstream
    gradients on the four pooled outputs, against autograd through the bf16-emulating oracle (same roundings in the
    forward, straight-through), with the oracle's argmax injected.  No head in the path: what remains is the bf16 storage of
    dz / dh and the few ReLU units within ~3e-4 of zero, averaged over thousands of points."""
    from oracle import wireframe_oracle as wo
    from models.PointNetEncoder import PointNetEncoder
    from wf_b200 import ops
    from gpu_util import record
    ops.set_precision("bf16")
    try:
        torch.manual_seed(0)
        enc = PointNetEncoder().cuda()
        sd = {k: v for k, v in wo.make_state_dict(21, 16).items() if k.startswith("encoder.")}
        enc.load_state_dict({k[len("encoder."):]: v for k, v in sd.items()})
        x, _, _ = wo.make_inputs(7, B, N, 16, pad_frac=0.1, norm_intensity=True)
        gen = torch.Generator().manual_seed(1)
        gs = [torch.randn(B, 512, generator=gen) for _ in range(4)]
        sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        pf, h4 = wo.encoder_point_features(sdr, x, True, return_hidden=True)
        mask, _, mx_m, arg_m = wo.encoder_pools(x, pf)
        cnt = mask.sum(dim=1, keepdim=True).clamp(min=1).float()
        h3 = h4.reshape(B, N, -1)
        avg_m = torch.nn.functional.linear((h3 * mask.unsqueeze(-1)).sum(1) / cnt, sdr["encoder.mlp.16.weight"], sdr["encoder.mlp.16.bias"])
        mean_u = torch.nn.functional.linear(h3.mean(1), sdr["encoder.mlp.16.weight"], sdr["encoder.mlp.16.bias"])
        mx_u, arg_u = pf.max(dim=1)
        (mx_m * gs[0] + avg_m * gs[1] + mx_u * gs[2] + mean_u * gs[3]).sum().backward()
        ops.ARGMAX_OVERRIDE = (arg_m.to(torch.int32).cuda(), arg_u.to(torch.int32).cuda())
        r = enc.pooled(x.cuda())
        ops.ARGMAX_OVERRIDE = None
        sum(t * g_.cuda() for t, g_ in zip(r[:4], gs)).sum().backward()
        obs = {"pooled": max(rel_err(a, b_) for a, b_ in zip(r[:4], (mx_m, avg_m, mx_u, mean_u))),
               "argmax_agreement": float((r[5].cpu() == arg_u).float().mean())}
        assert obs["pooled"] < 2e-3
        worst_f, worst_e = ("", 0.0), ("", 0.0)
        for k, p in enc.named_parameters():
            if not k.startswith("mlp."):
                continue
            a, b_ = p.grad.double().cpu().reshape(-1), sdr["encoder." + k].grad.double().reshape(-1)
            fro = float((a - b_).norm() / b_.norm())
            el = float((a - b_).abs().max()) / max(float(b_.abs().max()), float(b_.norm()) / np.sqrt(b_.numel()))
            worst_f = max(worst_f, (k, fro), key=lambda t: t[1]); worst_e = max(worst_e, (k, el), key=lambda t: t[1])
        obs.update(grad_fro_worst=worst_f[1], grad_fro_worst_param=worst_f[0], grad_elem_worst=worst_e[1], grad_elem_worst_param=worst_e[0])
        record(f"bf16_encoder_bwd_vs_emulating_oracle/B{B}_N{N}", **obs)
        print(obs)
        assert worst_f[1] < 3e-2 and worst_e[1] < 6e-2, obs
    finally:
        ops.ARGMAX_OVERRIDE = None
        ops.set_precision("bf16")


def test_counts_of_fresh_tensors_are_never_stale():
    """a fresh counts tensor per batch gets the previous batch's recycled device address with version 0;
    the edge head must still run with ITS counts (edge_probs width, edge_indices) -- through freshly allocated tensors."""
    from oracle import wireframe_oracle as wo
    from wf_b200 import ops
    ops.set_precision("fp32")
    try:
        V = 12
        m = _model(3, V, True)
        x, _, _ = wo.make_inputs(3, 2, 256, V, norm_intensity=True)
        xg = x.cuda()
        ptrs = set()
        for step, cs in enumerate(([5, 7], [9, 3], [2, 12], [6, 6], [11, 4])):
            c = torch.tensor(cs).cuda()                       # fresh tensor each step
            ptrs.add(c.data_ptr())
            out = m(xg, c)
            assert [len(e) for e in out["edge_indices"]] == [k * (k - 1) // 2 for k in cs], (step, cs)
            assert out["edge_probs"].shape[1] == max(k * (k - 1) // 2 for k in cs)
            del c, out
        # in-place update of ONE tensor (version bump) is seen too
        c = torch.tensor([5, 7]).cuda()
        m(xg, c)
        c.copy_(torch.tensor([4, 8]))
        assert [len(e) for e in m(xg, c)["edge_indices"]] == [6, 28]
        print("distinct device addresses over 5 fresh count tensors:", len(ptrs))
    finally:
        ops.set_precision("bf16")


@pytest.mark.parametrize("name", ["eval_b2_n256_v16", "eval_b2_n10000_v64"])
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_eval_forward_vs_reference_golden(golden_dir, name, prec):
    """evaluate.py:71 (eval mode, no_grad, counts from the existence probabilities) against the unmodified reference, incl.
    BASELINE.json configs[1]'s shape (10,000-point clouds, 64 vertex slots; the existence bias of the fixture spreads the
    probabilities so that the data-dependent counts are 32 of 64).  bf16 takes the chunked inference encoder."""
    from oracle import wireframe_oracle as wo
    from wf_b200 import ops
    from gpu_util import record
    g = dict(np.load(os.path.join(golden_dir, name + ".npz")))
    seed, B, N, V = [int(v) for v in g["meta"][:4]]
    ops.set_precision(prec)
    try:
        m = _model(seed, V, False)
        if g["exist_bias"].size:
            with torch.no_grad():
                m.vertex_predictor.final_layer.bias.view(V, 4)[:, 3] += torch.from_numpy(g["exist_bias"]).cuda()
        x, tgt, counts = wo.make_inputs(seed, B, N, V, norm_intensity=True)
        with torch.no_grad():
            pred = m(x.cuda(), counts.cuda())
        otol = TOL[prec]["out"]
        # counts are a threshold on the existence probabilities: identical in fp32; in bf16 wherever no probability sits
        # within the tolerance of 0.5
        near = np.abs(g["existence"] - 0.5) < (1e-5 if prec == "fp32" else 2e-2)
        if not near.any():
            assert np.array_equal(pred["actual_vertex_counts"].cpu().numpy(), g["dyn_counts"])
        obs = {"vertices": assert_close(pred["vertices"], torch.from_numpy(g["vertices"]), otol, "vertices"),
               "existence": assert_close(pred["existence_probabilities"], torch.from_numpy(g["existence"]), otol, "existence")}
        if np.array_equal(pred["actual_vertex_counts"].cpu().numpy(), g["dyn_counts"]):
            obs["edge_probs"] = assert_close(pred["edge_probs"], torch.from_numpy(g["edge_probs"]), otol, "edge_probs")
            assert np.array_equal(np.asarray(pred["edge_indices"][0]), g["edge_indices0"])
            assert pred["edge_probs"].shape == tuple(g["edge_probs"].shape)
        record(f"eval/{name}/{prec}", **obs)
    finally:
        ops.set_precision("bf16")


def test_public_submodule_api():
    """PointNetEncoder.forward / VertexPredictor.forward / EdgePredictor.forward keep the reference signatures."""
    from oracle import wireframe_oracle as wo
    from wf_b200 import ops
    ops.set_precision("fp32")
    try:
        m = _model(9, 10, False)
        sd = wo.make_state_dict(9, 10)
        x, _, _ = wo.make_inputs(9, 2, 200, 10, norm_intensity=True)
        with torch.no_grad():
            gf, pf = m.encoder(x.cuda())
            gf_ref, pf_ref = wo.encoder_forward(sd, x)
            assert pf.shape == (2, 200, 512)
            assert_close(pf, pf_ref, 2e-5, "point_features"); assert_close(gf, gf_ref, 2e-5, "global_features")
            vo = m.vertex_predictor(gf, pf, None)
            v_ref, p_ref, c_ref = wo.vertex_forward(sd, gf_ref, pf_ref, 10)
            assert_close(vo["vertices"], v_ref, 2e-5, "vertices")
            assert torch.equal(vo["actual_vertex_counts"].cpu(), c_ref)
            probs, idx = m.edge_predictor(vo["vertices"][:, :6, :].contiguous())
            for b in range(2):
                pr, pairs = wo.edge_forward(sd, v_ref[b, :6])
                assert_close(probs[b], pr, 2e-5, "edge probs"); assert idx == pairs.tolist()
            with pytest.raises(IndexError):
                m.edge_predictor(vo["vertices"][:, :1, :].contiguous())
        ops.set_precision("bf16")
        with torch.no_grad():
            gf2, pf2 = m.encoder(x.cuda())
        assert_close(pf2, pf_ref, 1e-2, "bf16 point_features")
    finally:
        ops.set_precision("bf16")


def test_matchers_vs_reference_golden(golden_dir):
    from models.WireframeHungarianMatcher import WireframeHungarianMatcher
    from models.HungarianMatcher import HungarianMatcher
    g = dict(np.load(os.path.join(golden_dir, "matchers.npz")))
    seed, B, V = [int(v) for v in g["meta"]]
    rng = np.random.Generator(np.random.PCG64(seed))
    outputs = {"vertices": torch.from_numpy(rng.uniform(-1, 1, (B, V, 3)).astype(np.float32)).cuda(),
               "existence_probabilities": torch.from_numpy(rng.uniform(0, 1, (B, V)).astype(np.float32)).cuda()}
    tg = [{"vertices": torch.from_numpy(rng.uniform(-1, 1, (t, 3)).astype(np.float32)), "existence": torch.ones(t)}
          for t in g["sizes"].tolist()]
    res = WireframeHungarianMatcher(cost_vertex=2.0, cost_existence=0.5)(outputs, tg)
    for b, (i, j) in enumerate(res):
        assert i.dtype == torch.int64 and not i.is_cuda
        assert np.array_equal(i.numpy(), g[f"wf_i/{b}"]) and np.array_equal(j.numpy(), g[f"wf_j/{b}"]), b
    Q, K = 20, 7
    det = {"pred_logits": torch.from_numpy(rng.normal(size=(B, Q, K)).astype(np.float32)).cuda(),
           "pred_boxes": torch.from_numpy(np.concatenate([rng.uniform(0.2, 0.8, (B, Q, 2)),
                                                          rng.uniform(0.05, 0.3, (B, Q, 2))], -1).astype(np.float32)).cuda()}
    dt = [{"labels": torch.from_numpy(rng.integers(0, K, (t,)).astype(np.int64)),
           "boxes": torch.from_numpy(np.concatenate([rng.uniform(0.2, 0.8, (t, 2)),
                                                     rng.uniform(0.05, 0.3, (t, 2))], -1).astype(np.float32))}
          for t in g["detr_sizes"].tolist()]
    res = HungarianMatcher(cost_class=1.0, cost_bbox=5.0, cost_giou=2.0)(det, dt)
    for b, (i, j) in enumerate(res):
        assert np.array_equal(i.numpy(), g[f"detr_i/{b}"]) and np.array_equal(j.numpy(), g[f"detr_j/{b}"]), b


def test_loss_ties_vs_reference_golden(golden_dir):
    from losses.WireframeLoss import WireframeLoss
    g = dict(np.load(os.path.join(golden_dir, "loss_ties.npz")))
    seed, B, V = [int(v) for v in g["meta"]]
    rng = np.random.Generator(np.random.PCG64(seed))
    pv = np.round(rng.uniform(-1, 1, (B, V, 3)) * 2) / 2
    pe = np.round(rng.uniform(0, 1, (B, V)) * 4) / 4
    tv = np.zeros((B, V, 3)); counts = g["counts"]
    for b in range(B):
        tv[b, :counts[b]] = np.round(rng.uniform(-1, 1, (counts[b], 3)) * 2) / 2
    pred = {"vertices": torch.from_numpy(pv.astype(np.float32)).cuda(),
            "existence_probabilities": torch.from_numpy(pe.astype(np.float32)).cuda()}
    tgt = {"vertices": torch.from_numpy(tv.astype(np.float32)).cuda(), "vertex_counts": torch.from_numpy(counts).cuda()}
    for b, (i, j) in enumerate(WireframeLoss()._hungarian_matching(pred, tgt)):
        assert np.array_equal(i, g[f"p/{b}"]) and np.array_equal(j, g[f"t/{b}"]), b


def test_lazy_point_pool_proj_like_reference():
    """the projection appears on the first forward; an optimizer built before does not see it."""
    from models.PointCloudToWireframe import PointCloudToWireframe
    torch.manual_seed(0)
    m = PointCloudToWireframe(max_vertices=8).cuda().train()
    n0 = sum(p.numel() for p in m.parameters())
    assert "vertex_predictor.point_pool_proj.weight" not in m.state_dict()
    x = torch.rand(2, 64, 8, device="cuda")
    m(x, torch.tensor([3, 4], device="cuda"))
    assert sum(p.numel() for p in m.parameters()) == n0 + 524800
    assert m.vertex_predictor.point_pool_proj.weight.is_cuda


def test_loss_raises_scipy_errors_immediately_or_deferred():
    """NaN predictions make scipy raise inside the reference's forward (losses/WireframeLoss.py:236).  check_status=True
    reproduces that; the default "deferred" raises the same ValueError at the next call / check_pending() without a
    device->host sync inside the step."""
    from losses.WireframeLoss import WireframeLoss
    B, V = 3, 8
    torch.manual_seed(0)
    good = {"vertices": torch.rand(B, V, 3, device="cuda"), "existence_probabilities": torch.rand(B, V, device="cuda"),
            "edge_probs": torch.rand(B, 6, device="cuda")}
    bad = {k: v.clone() for k, v in good.items()}
    bad["vertices"][1, 2, 0] = float("nan")
    tgt = {"vertices": torch.rand(B, V, 3, device="cuda"), "vertex_existence": torch.ones(B, V, device="cuda"),
           "edge_labels": torch.ones(B, 6, device="cuda"), "vertex_counts": torch.tensor([4, 4, 4], device="cuda")}
    crit = WireframeLoss()
    crit.check_status = True
    with pytest.raises(ValueError, match="invalid numeric"):
        crit(bad, tgt)
    crit.check_status = "deferred"
    crit(bad, tgt)                                   # no exception yet
    with pytest.raises(ValueError, match="invalid numeric"):
        crit(good, tgt)                              # surfaces at the next call
    crit(bad, tgt)
    with pytest.raises(ValueError, match="invalid numeric"):
        crit.check_pending()
    crit(good, tgt); crit.check_pending()            # clean


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_real_building_vs_reference_golden(golden_dir, prec):
    """BASELINE.json configs[0] / SURVEY 8d config 1: a REAL building as the reference's loader delivers it (2560 points,
    RGBA/256, un-normalised intensity ~5e4), targets prepared ON THE DEVICE (wf_pack_targets), one training step through the
    drop-in modules -- against the unmodified reference's outputs, losses, matching and all gradient norms."""
    from wf_b200 import ops
    from wf_b200.targets import prepare_targets
    from losses.WireframeLoss import WireframeLoss
    g = dict(np.load(os.path.join(golden_dir, "real_b1_n2560_v38.npz")))
    seed, B, N, V = [int(v) for v in g["meta"]]
    ops.set_precision(prec)
    try:
        m = _model(seed, V, True)
        tgt = prepare_targets([torch.from_numpy(g["wf_vertices"])], [torch.from_numpy(g["wf_edges"])], V, "cuda")
        assert np.array_equal(tgt["edge_labels"].cpu().numpy(), g["tgt_edge_labels"])
        assert np.array_equal(tgt["vertices"].cpu().numpy(), g["tgt_vertices"])
        assert np.array_equal(tgt["vertex_existence"].cpu().numpy(), g["tgt_existence"])
        assert np.array_equal(tgt["vertex_counts"].cpu().numpy(), g["tgt_counts"])
        x = torch.from_numpy(g["x"]).cuda()
        pred = m(x, tgt["vertex_counts"])
        crit = WireframeLoss(vertex_weight=3.0, edge_weight=1.0, existence_weight=1.5)
        ld = crit(pred, tgt)
        ld["total_loss"].backward()
        otol, gtol = (5e-4, 5e-3) if prec == "fp32" else (5e-2, 2.5e-1)
        assert_close(pred["vertices"], torch.from_numpy(g["vertices"]), otol, "vertices")
        assert_close(pred["existence_probabilities"], torch.from_numpy(g["existence"]), otol, "existence")
        assert_close(pred["edge_probs"], torch.from_numpy(g["edge_probs"]), otol, "edge_probs")
        assert_close(pred["global_features"], torch.from_numpy(g["global_features"]), otol, "global_features")
        got = np.array([ld[k].item() for k in ("total_loss", "vertex_loss", "existence_loss", "edge_loss")])
        np.testing.assert_allclose(got[2:], g["losses"][2:], rtol=otol * 5, atol=otol)
        # With random-init weights the 38 predicted vertices of this building are nearly coincident, so many assignment
        # costs are tied to ~1e-4 and the optimal matching is not stable under fp32 summation-order noise (raw intensity,
        # SURVEY D6): the vertex term is compared loosely, and the matching against the oracle's solver on OUR predictions
        # (same input -> must be identical, like tests/test_gpu_simt.py holds the solver to scipy).
        np.testing.assert_allclose(got[:2], g["losses"][:2], rtol=5e-2)
        from oracle import wireframe_oracle as wo
        pred_cpu = {k: (v.detach().cpu() if torch.is_tensor(v) else v) for k, v in pred.items()}
        ref_match = wo.loss_matching(pred_cpu, {k: v.cpu() for k, v in tgt.items()})[0]
        pi, ti = crit._hungarian_matching(pred, tgt)[0]
        assert np.array_equal(pi, ref_match[0]) and np.array_equal(ti, ref_match[1])
        fails = []
        from gpu_util import record
        obs = {"vertices": rel_err(pred["vertices"], torch.from_numpy(g["vertices"])),
               "edge_probs": rel_err(pred["edge_probs"], torch.from_numpy(g["edge_probs"])),
               "global_features": rel_err(pred["global_features"], torch.from_numpy(g["global_features"])),
               "loss_exist_edge": float(np.max(np.abs(got[2:] - g["losses"][2:]) / g["losses"][2:])),
               "loss_total_vertex": float(np.max(np.abs(got[:2] - g["losses"][:2]) / g["losses"][:2])), "gnorm": {}}
        for k, p in m.named_parameters():
            if "gnone/" + k in g:
                assert p.grad is None, k
                continue
            ref_norm = float(g["gnorm/" + k][0])
            e = abs(float(p.grad.double().norm()) - ref_norm) / max(ref_norm, 1e-12)
            obs["gnorm"][k] = round(e, 7)
            # raw intensity: the first LayerNorm's backward is ill-conditioned in fp32 (in the reference too); a different
            # (equally optimal within noise) matching changes the vertex-term gradients of the heads
            loose = 30.0 if k.startswith("encoder.mlp.0.") else (10.0 if prec == "fp32" else 1.0)
            if e > gtol * loose:
                fails.append((k, e))
        obs["gnorm_worst"] = max(obs["gnorm"].values()); obs["gnorm_worst_excl_l1"] = max(v for k, v in obs["gnorm"].items() if not k.startswith("encoder.mlp.0."))
        del obs["gnorm"]
        record(f"real_building/{prec}", **obs)
        assert not fails, f"{prec} gradient-norm mismatches: {fails[:6]}"
    finally:
        ops.set_precision("bf16")


def test_matching_beside_the_edge_head_is_identical_and_falls_back_safely():
    """WireframeLoss runs its assignment on a side stream when the vertex head's outputs carry a ready-event and the
    targets are known complete (tagged by wf_b200.targets, or the same tensors as in the previous call -- train.py's
    loop).  Results must not depend on where it ran; unknown fresh targets must take the main-stream path."""
    from oracle import wireframe_oracle as wo
    from losses.WireframeLoss import WireframeLoss
    from wf_b200 import ops
    ops.set_precision("fp32")
    seed, B, N, V = 5, 4, 512, 16
    model = _model(seed, V, True)
    x, tgt, counts = wo.make_inputs(seed, B, N, V, norm_intensity=True)
    x = x.cuda(); tgt = {k: v.cuda() for k, v in tgt.items()}
    runs = {}
    for overlap in (False, "repeat", True):
        crit = WireframeLoss(vertex_weight=3.0, edge_weight=1.0, existence_weight=1.5)
        crit.overlap_matching = overlap is True
        launched_on = []
        real = crit._launch_match
        crit._launch_match = lambda p, t, m, real=real: (launched_on.append(torch.cuda.current_stream().cuda_stream),
                                                        real(p, t, m))[1]
        vals = []
        for it in range(3):
            model.zero_grad()
            pred = model(x, tgt["vertex_counts"])
            assert hasattr(pred["vertices"], "_wf_ready")
            ld = crit(pred, tgt)
            ld["total_loss"].backward()
            vals.append((ld["total_loss"].item(), model.encoder.mlp[0].weight.grad.clone()))
        runs[overlap] = (vals, launched_on)
        crit.check_pending()
    main = torch.cuda.current_stream().cuda_stream
    assert all(s == main for s in runs[False][1])
    # first call: targets not yet seen and not tagged -> main stream; afterwards the side stream
    assert runs[True][1][0] == main and all(s != main for s in runs[True][1][1:])
    # weight gradients of split-K products are accumulated atomically: demand bit-equality of the gradient only if a plain
    # repeat of the main-stream run has it
    repeatable = all(torch.equal(a[1], b[1]) for a, b in zip(runs[False][0], runs["repeat"][0]))
    for (l0, g0), (l1, g1) in zip(runs[False][0], runs[True][0]):
        assert l0 == l1
        assert torch.equal(g0, g1) if repeatable else rel_err(g1, g0) < 1e-5
    # tagged targets (prepare_targets / DevicePrefetcher) overlap from the first call; in-place changed ones do not
    crit = WireframeLoss()
    seen = []
    real = crit._launch_match
    crit._launch_match = lambda p, t, m: (seen.append(torch.cuda.current_stream().cuda_stream), real(p, t, m))[1]
    tagged = {k: ops.mark_ready(v.clone()) for k, v in tgt.items()}
    pred = model(x, tgt["vertex_counts"])
    crit(pred, tagged)
    assert seen[-1] != main
    crit(pred, tgt); crit(pred, tgt)
    assert seen[-1] != main
    tgt["vertices"].mul_(1.0)                          # version bump: completion unknown again
    crit(pred, tgt)
    assert seen[-1] == main
    ops.set_precision("bf16")


@pytest.mark.parametrize("B,N,V,max_count", [(1, 100, 5, 5), (2, 333, 128, 100), (3, 257, 200, 130), (1, 2560, 64, 2)])
def test_unusual_shapes_vs_oracle(B, N, V, max_count):
    """`max_vertices` is whatever the first batch holds (train.py:37) and clouds need not fill GEMM tiles:
    tiny and large vertex-slot counts (the matcher drops to fewer samples per CTA above ~110 slots), odd point counts,
    the two-vertex minimum -- full fp32 training step against the oracle; assignments identical."""
    from oracle import wireframe_oracle as wo
    from losses.WireframeLoss import WireframeLoss
    from wf_b200 import ops
    ops.set_precision("fp32")
    try:
        seed = 11 + V
        model = _model(seed, V, True)
        x, tgt, counts = wo.make_inputs(seed, B, N, V, norm_intensity=True, max_count=max_count)
        sd = wo.make_state_dict(seed, V)
        sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        ld_ref, pred_ref = wo.train_step(sdr, x, tgt, max_vertices=V)
        crit = WireframeLoss(vertex_weight=3.0, edge_weight=1.0, existence_weight=1.5)
        pred = model(x.cuda(), counts.cuda())
        tg = {k: v.cuda() for k, v in tgt.items()}
        ld = crit(pred, tg)
        ld["total_loss"].backward()
        crit.check_pending()
        assert_close(pred["vertices"], pred_ref["vertices"], 2e-5, "vertices")
        assert_close(pred["edge_probs"], pred_ref["edge_probs"], 2e-5, "edge_probs")
        assert abs(ld["total_loss"].item() - ld_ref["total_loss"].item()) <= 2e-5 * abs(ld_ref["total_loss"].item())
        ours = crit._hungarian_matching(pred, tg)
        # the oracle's matching of OUR predictions: cost matrices bit-equal, assignment index-identical
        ref = wo.loss_matching({k: pred[k].detach().cpu() for k in ("vertices", "existence_probabilities")}, tgt)
        for (a, b), (c, d) in zip(ours, ref):
            assert np.array_equal(a, c) and np.array_equal(b, d)
        # The oracle matched ITS OWN predictions.  With ~100 constant dummy columns per row and fp32 noise of 1e-7 between the
        # CPU and GPU forward, two near-equal assignments can swap (a discontinuity of the loss, same value to 1e-5): the
        # gradients that flow through the matched vertices are then compared only when both sides chose the same matching.
        own = wo.loss_matching(pred_ref, tgt)
        same = all(np.array_equal(a, c) and np.array_equal(b, d) for (a, b), (c, d) in zip(ours, own))
        names = ["edge_predictor.edge_mlp.0.weight", "edge_predictor.attention.in_proj_weight"]
        if same:
            names += ["encoder.mlp.0.weight", "vertex_predictor.final_layer.weight"]
        else:
            assert V >= 200, "matching flipped between oracle and GPU predictions on a small problem"
        for name in names:
            g = dict(model.named_parameters())[name].grad
            assert_close(g, sdr[name].grad, 2e-4, name)
    finally:
        ops.set_precision("bf16")


@pytest.mark.parametrize("input_dim", [3, 7])
def test_fewer_input_features_use_the_tensor_core_encoder(input_dim):
    """datasets/building3d.py:103-111: use_color / use_intensity off -> 3, 4 or 7 features.  The bf16 tensor-core encoder
    takes them zero-padded to 8; results equal the generic fp32 path of the same module within the bf16 tolerance and the
    oracle within it; the first layer's weight gradient has the module's own shape."""
    from oracle import wireframe_oracle as wo
    from models.PointNetEncoder import PointNetEncoder
    from wf_b200 import ops
    seed, B, N = 21, 2, 640
    torch.manual_seed(0)
    enc = PointNetEncoder(input_dim=input_dim).cuda()
    assert enc._tc_shape
    x, _, _ = wo.make_inputs(seed, B, N, 8, norm_intensity=True)
    x = x[:, :, :input_dim].contiguous().cuda()
    g = [torch.randn(B, 512, device="cuda") for _ in range(4)]
    outs, grads = {}, {}
    for prec in ("fp32", "bf16"):
        ops.set_precision(prec)
        enc.zero_grad()
        r = enc.pooled(x)
        # upstream gradients on the two MEAN pools only: max-pool routing is discontinuous under bf16-sized perturbations
        # (tests/test_gpu_tc.py::test_encoder_tc_vs_fp32_path covers that case by norm)
        (r[1] * g[1] + r[3] * g[3]).sum().backward()
        outs[prec] = [t.detach().clone() for t in r[:4]]
        grads[prec] = {k: p.grad.detach().clone() for k, p in enc.named_parameters() if p.grad is not None}
    ops.set_precision("bf16")
    assert grads["bf16"]["mlp.0.weight"].shape == (512, input_dim)
    for a, b in zip(outs["bf16"], outs["fp32"]):
        assert_close(a, b, 1e-2, "pooled features bf16 vs fp32 path")
    for k in ("mlp.0.weight", "mlp.4.weight", "mlp.16.weight"):
        ga, gb = grads["bf16"][k].double(), grads["fp32"][k].double()
        assert float((ga - gb).norm() / gb.norm()) < 8e-2, k
    # fp32 path vs the oracle (state dict of the same weights)
    sd = {f"encoder.{k}": v.detach().cpu() for k, v in enc.state_dict().items()}
    pf = wo.encoder_point_features(sd, x.cpu())
    ref_max = pf.max(dim=1).values
    assert_close(outs["fp32"][2], ref_max, 2e-5, "unmasked max vs oracle")
    # inference (chunked, no grad) takes the same route
    with torch.no_grad():
        ri = enc.pooled(x)
    for a, b in zip(ri[:4], outs["bf16"]):
        assert torch.equal(a, b) or rel_err(a, b) < 1e-6


def test_hungarian_rmse_vs_reference_golden(golden_dir):
    """models/utils.py:38-55 on wf_cdist_f64 + wf_lsap_f64: identical to the unmodified reference (fp64 cdist, fp64
    assignment) on random, rectangular and heavily tied vertex sets; float32 inputs keep the reference's float32 final
    expression; NaN input raises scipy's ValueError like the reference's linear_sum_assignment call."""
    from models.utils import hungarian_rmse
    g = dict(np.load(os.path.join(golden_dir, "hungarian_rmse.npz")))
    for k, want in enumerate(g["rmse"]):
        got = hungarian_rmse(g[f"p/{k}"], g[f"t/{k}"])
        assert abs(float(got) - float(want)) <= 1e-12 * max(1.0, abs(float(want))), (k, got, want)
        assert np.asarray(got).dtype == np.result_type(g[f"p/{k}"].dtype, g[f"t/{k}"].dtype)
    assert hungarian_rmse(np.zeros((0, 3)), np.zeros((0, 3))) == 0.0
    assert hungarian_rmse(np.zeros((0, 3)), np.ones((2, 3))) == float("inf")
    bad = g["p/0"].copy(); bad[1, 2] = np.nan
    with pytest.raises(ValueError, match="invalid numeric"):
        hungarian_rmse(bad, g["t/0"])
