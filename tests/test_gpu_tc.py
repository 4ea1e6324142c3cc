"""GPU parity of the tcgen05 GEMM (wf_gemm_bf16) against torch fp32 matmul of the same bf16 operands,
and of the bf16 encoder building blocks.  bf16 operands are exact in both; the only difference is
fp32 accumulation order, so the tolerance is tight (scale-relative 1e-4 before the output rounding)."""
import math

import pytest
import torch

from gpu_util import assert_close, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from wf_b200 import ops as o
    return o


SHAPES_K = [(128, 256, 64), (128, 256, 512), (256, 512, 1024), (1000, 1024, 512), (130, 264, 72), (77, 40, 8),
            (4096, 2048, 1024), (20000, 512, 1024)]


@pytest.mark.parametrize("M,N,K", SHAPES_K)
def test_gemm_bf16_kmajor_fp32_out(ops, M, N, K):
    torch.manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    B = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    out = torch.full((M, N), float("nan"), device="cuda")
    from wf_b200._lib import call
    parts = call("wf_gemm_rowstats_parts", N)
    stats = torch.full((parts, M, 2), float("nan"), device="cuda")
    ops.gemm_bf16(A, B, M=M, N=N, K=K, bias=bias, out=out, rowstats=stats)
    ref = A.float().double() @ B.float().double().t() + bias.double()
    assert_close(out, ref, 1e-4, f"kmajor f32 {M}x{N}x{K}")
    assert_close(stats[:, :, 0].sum(0), ref.sum(1), 1e-4, "rowstats sum")
    assert_close(stats[:, :, 1].sum(0), (ref * ref).sum(1), 1e-4, "rowstats sumsq")
    mean = torch.empty(M, device="cuda"); rstd = torch.empty(M, device="cuda")
    call("wf_stats_finalize", ops._p(stats), M, N, parts, 1e-5, ops._p(mean), ops._p(rstd), ops._s())
    assert_close(mean, ref.mean(1), 1e-4, "row mean")
    assert_close(rstd, (ref.var(1, unbiased=False) + 1e-5).rsqrt(), 1e-3, "row rstd")
    # deterministic: a second run gives the same bits (no floating-point atomics in the forward path)
    out2 = torch.empty_like(out); stats2 = torch.empty_like(stats)
    ops.gemm_bf16(A, B, M=M, N=N, K=K, bias=bias, out=out2, rowstats=stats2)
    assert torch.equal(out, out2) and torch.equal(stats, stats2)


@pytest.mark.parametrize("M,N,K", [(256, 512, 256), (1000, 1024, 512), (130, 264, 72)])
def test_gemm_bf16_kmajor_bf16_out(ops, M, N, K):
    torch.manual_seed(1)
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    B = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(torch.bfloat16)
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm_bf16(A, B, M=M, N=N, K=K, out=out)
    ref = A.float() @ B.float().t()
    assert_close(out.float(), ref, 6e-3, "kmajor bf16 out")          # one bf16 rounding of the result


@pytest.mark.parametrize("Mo,No,Kr,split", [(128, 256, 64, 1), (512, 1024, 4096, 4), (1024, 512, 10000, 7),
                                              (200, 72, 333, 3), (2048, 1024, 50000, 5)])
def test_gemm_bf16_mnmajor_splitk(ops, Mo, No, Kr, split):
    """D[Mo,No] = A^T B with A [Kr,Mo], B [Kr,No] (the dW = dZ^T H shape), split-K atomic accumulate."""
    torch.manual_seed(Mo + Kr)
    A = torch.randn(Kr, Mo, device="cuda").to(torch.bfloat16)
    B = (torch.randn(Kr, No, device="cuda") / math.sqrt(Kr)).to(torch.bfloat16)
    out = torch.zeros(Mo, No, device="cuda")
    ops.gemm_bf16(A, B, M=Mo, N=No, K=Kr, kmajor=False, out=out, accumulate=True, split_k=split)
    ref = A.float().double().t() @ B.float().double()
    assert_close(out, ref, 1e-4, f"mnmajor {Mo}x{No}x{Kr} split {split}")


def test_cast_and_stats(ops):
    torch.manual_seed(2)
    W = torch.randn(300, 520, device="cuda")
    assert torch.equal(ops.cast_bf16(W), W.to(torch.bfloat16))
    assert torch.equal(ops.cast_bf16(W, transpose=True), W.t().contiguous().to(torch.bfloat16))


def _fro(a, b):
    a = a.double(); b = b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def test_encoder_tc_vs_fp32_path(ops):
    """bf16 tensor-core encoder (four pooled outputs + all 18 parameter gradients) against the fp32 SIMT path
    of this library on the same weights and the same upstream gradients.

    (a) upstream gradients on the two MEAN pools only: no argmax in the path, so the only difference is
        bf16 rounding of operands/activations -> Frobenius-relative error <= 3e-2 on every gradient.
    (b) all four pools: max-pool routing is discontinuous (a bf16-sized perturbation can move the argmax to
        another point), so gradients are compared by norm and the argmax agreement is reported."""
    from oracle import wireframe_oracle as wo
    from models.PointNetEncoder import PointNetEncoder
    torch.manual_seed(0)
    enc = PointNetEncoder().cuda()
    sd = {k[len("encoder."):]: v for k, v in wo.make_state_dict(21, 16).items() if k.startswith("encoder.")}
    enc.load_state_dict(sd)
    x, _, _ = wo.make_inputs(3, 2, 700, 16, pad_frac=0.1, norm_intensity=True)
    x = x.cuda()
    gs = [torch.randn(2, 512, device="cuda") for _ in range(4)]
    for use_max in (False, True):
        res = {}
        for prec in ("fp32", "bf16"):
            ops.set_precision(prec)
            enc.zero_grad()
            r = enc.pooled(x)
            obj = (r[1] * gs[1] + r[3] * gs[3]).sum()
            if use_max:
                obj = obj + (r[0] * gs[0] + r[2] * gs[2]).sum()
            obj.backward()
            res[prec] = ([t.detach().clone() for t in r[:6]],
                         {k: p.grad.detach().clone() for k, p in enc.named_parameters() if p.grad is not None})
        ops.set_precision("bf16")
        for a, b, n in zip(res["bf16"][0][:4], res["fp32"][0][:4], ("max_m", "avg_m", "max_u", "mean_u")):
            assert_close(a, b, 3e-2, f"pooled {n}")
        agree = float((res["bf16"][0][5] == res["fp32"][0][5]).float().mean())
        worst = ("", 0.0)
        for k, g in res["fp32"][1].items():
            if not k.startswith("mlp."):
                continue
            gb = res["bf16"][1][k]
            e = _fro(gb, g) if not use_max else abs(float(gb.norm() / g.norm()) - 1.0)
            if e > worst[1]:
                worst = (k, e)
            print(f"   use_max={use_max} {k:16s} err {e:.3e}")
            # 8e-2: the first layer's bias/gain gradients are sums with LayerNorm cancellation, the most noise-sensitive
            assert e < (8e-2 if not use_max else 2.5e-1), f"use_max={use_max} grad {k}: {e:.3e}"
        print(f"encoder bf16 vs fp32 (max pools in path: {use_max}): worst grad error {worst}, argmax agreement {agree:.4f}")
        from gpu_util import record
        record(f"encoder_tc_vs_fp32/use_max={use_max}", worst_param=worst[0], worst=worst[1], argmax_agreement=agree,
               pooled=max(rel_err(a, b) for a, b in zip(res["bf16"][0][:4], res["fp32"][0][:4])))


def test_enc_l1_kernels_fp32_dtype(ops):
    """The fused first-layer kernels (wf_enc_l1_fwd / wf_enc_l1_bwd) with fp32 output/gradient dtype against torch
    autograd in fp64 -- isolates the kernel logic from bf16 rounding."""
    import torch.nn.functional as F
    from wf_b200._lib import call, F32
    torch.manual_seed(5)
    M = 1000
    x = torch.randn(M, 8, device="cuda"); x[:, 7] = x[:, 7] * 3 + 5
    W = torch.randn(512, 8, device="cuda") / 8 ** 0.5; b = 0.1 * torch.randn(512, device="cuda")
    g = 1 + 0.1 * torch.randn(512, device="cuda"); be = 0.1 * torch.randn(512, device="cuda")
    d = [t.double().requires_grad_(True) for t in (x, W, b, g, be)]
    ref = torch.relu(F.layer_norm(F.linear(d[0], d[1], d[2]), (512,), d[3], d[4], 1e-5))
    go = torch.randn_like(ref); ref.backward(go)
    h = torch.empty(M, 512, device="cuda")
    call("wf_enc_l1_fwd", ops._p(x), ops._p(W), ops._p(b), ops._p(g), ops._p(be), ops._p(h), F32, M, 8, 512, 1e-5, ops._s())
    assert_close(h, ref, 1e-5, "l1 fwd")
    go32 = go.float().contiguous()
    dW = torch.zeros(512, 8, device="cuda"); db = torch.zeros(512, device="cuda"); dg = torch.zeros(512, device="cuda")
    dbe = torch.zeros(512, device="cuda"); dx = torch.empty(M, 8, device="cuda")
    call("wf_enc_l1_bwd", ops._p(x), ops._p(W), ops._p(b), ops._p(g), ops._p(be), ops._p(go32), F32, ops._p(dW), ops._p(db),
         ops._p(dg), ops._p(dbe), ops._p(dx), M, 8, 512, 1e-5, ops._s())
    for n, a, r in (("dx", dx, d[0].grad), ("dW", dW, d[1].grad), ("db", db, d[2].grad), ("dgamma", dg, d[3].grad),
                    ("dbeta", dbe, d[4].grad)):
        assert_close(a, r, 5e-5, f"l1 bwd {n}")
    dW2 = torch.zeros(512, 8, device="cuda"); db2 = torch.zeros(512, device="cuda"); dg2 = torch.zeros(512, device="cuda")
    dbe2 = torch.zeros(512, device="cuda")
    call("wf_enc_l1_bwd", ops._p(x), ops._p(W), ops._p(b), ops._p(g), ops._p(be), ops._p(go32), F32, ops._p(dW2), ops._p(db2),
         ops._p(dg2), ops._p(dbe2), None, M, 8, 512, 1e-5, ops._s())
    for n, a, r in (("dW", dW2, d[1].grad), ("db", db2, d[2].grad), ("dgamma", dg2, d[3].grad), ("dbeta", dbe2, d[4].grad)):
        assert_close(a, r, 5e-5, f"l1 bwd {n} (no dx: channel-stationary kernel)")


@pytest.mark.parametrize("variant", ["mma_ntl1", "mma_ntl2", "simt"])
@pytest.mark.parametrize("M,raw_intensity", [(1, False), (255, False), (257, True), (5003, True), (70001, False)])
def test_enc_l1_channel_stationary_kernels(ops, M, raw_intensity, variant, monkeypatch):
    """The production first-layer kernels (LayerNorm statistics from the 9x9 Cholesky factor of the centred layer instead of a
    reduction over channels) against torch autograd in fp64, incl. un-normalised intensity (~5e4) and point counts
    that are not multiples of the staging block / the 8-point group / one block per CTA.  Variants of the bf16-gradient
    backward: the SIMT kernel (l1c::bwd_kernel, the default) and the 3xTF32 mma.sync kernel with 8 or 16 points per barrier
    (l1m::bwd_kernel<1|2>, WF_B200_L1_BWD=mma1|mma2; correct but slower on a B200, kept as the measured alternative)."""
    import torch.nn.functional as F
    from wf_b200._lib import call, F32, BF16
    monkeypatch.setenv("WF_B200_L1_BWD", {"mma_ntl1": "mma1", "mma_ntl2": "mma2", "simt": "simt"}[variant])
    torch.manual_seed(M)
    x = torch.randn(M, 8, device="cuda")
    if raw_intensity:
        x[:, 7] = torch.rand(M, device="cuda") * 4e4 + 2e4
    W = torch.randn(512, 8, device="cuda") / 8 ** 0.5; b = 0.1 * torch.randn(512, device="cuda")
    g = 1 + 0.1 * torch.randn(512, device="cuda"); be = 0.1 * torch.randn(512, device="cuda")
    d = [t.double().requires_grad_(True) for t in (x, W, b, g, be)]
    ref = torch.relu(F.layer_norm(F.linear(d[0], d[1], d[2]), (512,), d[3], d[4], 1e-5))
    go = torch.randn_like(ref).to(torch.bfloat16).double()
    ref.backward(go)
    h = torch.empty(M, 512, device="cuda")
    call("wf_enc_l1_fwd", ops._p(x), ops._p(W), ops._p(b), ops._p(g), ops._p(be), ops._p(h), F32, M, 8, 512, 1e-5, ops._s())
    assert_close(h, ref, 2e-5, "l1 fwd (fp32 out)")
    hb = torch.empty(M, 512, device="cuda", dtype=torch.bfloat16)
    call("wf_enc_l1_fwd", ops._p(x), ops._p(W), ops._p(b), ops._p(g), ops._p(be), ops._p(hb), BF16, M, 8, 512, 1e-5, ops._s())
    assert_close(hb.float(), ref, 5e-3, "l1 fwd (bf16 out)")
    for dt, gten in ((F32, go.float().contiguous()), (BF16, go.to(torch.bfloat16).contiguous())):
        dW = torch.zeros(512, 8, device="cuda"); db = torch.zeros(512, device="cuda"); dg = torch.zeros(512, device="cuda")
        dbe = torch.zeros(512, device="cuda")
        call("wf_enc_l1_bwd", ops._p(x), ops._p(W), ops._p(b), ops._p(g), ops._p(be), ops._p(gten), dt, ops._p(dW), ops._p(db),
             ops._p(dg), ops._p(dbe), None, M, 8, 512, 1e-5, ops._s())
        # raw intensity: the LayerNorm backward is ill-conditioned in fp32 (in the reference too)
        tol = 6e-3 if raw_intensity else 1e-4
        for n, a, r in (("dW", dW, d[1].grad), ("db", db, d[2].grad), ("dgamma", dg, d[3].grad), ("dbeta", dbe, d[4].grad)):
            assert_close(a, r, tol, f"l1 bwd {n} dtype {dt} M={M}")


def _tf32_rna(x):
    """cvt.rna.tf32.f32: round to nearest (ties away from zero) to 10 mantissa bits."""
    return ((x.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


@pytest.mark.parametrize("M,N,K", [(64, 2048, 4096), (64, 4096, 512), (64, 256, 1024), (3, 512, 1024), (128, 1024, 512),
                                   (100, 48, 1024), (64, 2048, 520), (1, 1024, 2048), (130, 512, 1024)])
def test_rowmlp_kernels(ops, M, N, K):
    """csrc/rowmlp.cu (the 64-row heads' weight-streaming products): forward Y = X W^T + b, dX = dZ W (W read in place) and
    dW = dZ^T X (+ db) against float64 products of the TF32-rounded operands (the kernels round once with cvt.rna and
    accumulate in fp32: 2e-5), against the unrounded float64 product at TF32 level (2e-3), and bit-reproducible."""
    from wf_b200._lib import call
    torch.manual_seed(M * 7 + N + K)
    X = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / K ** 0.5; b = torch.randn(N, device="cuda")
    dZ = torch.randn(M, N, device="cuda")
    Xr, Wr, dZr = _tf32_rna(X).double(), _tf32_rna(W).double(), _tf32_rna(dZ).double()
    p, s = ops._p, ops._s
    Y = torch.full((M, N), float("nan"), device="cuda")
    call("wf_rowmlp_linear", p(X), K, p(W), K, 0, p(b), M, N, K, p(Y), N, s())
    assert_close(Y, Xr @ Wr.t() + b.double(), 2e-5, "forward vs rounded operands")
    assert_close(Y, X.double() @ W.double().t() + b.double(), 2e-3, "forward vs exact")
    Y2 = torch.empty_like(Y)
    call("wf_rowmlp_linear", p(X), K, p(W), K, 0, p(b), M, N, K, p(Y2), N, s())
    assert torch.equal(Y, Y2), "forward not reproducible"
    dX = torch.full((M, K), float("nan"), device="cuda")
    call("wf_rowmlp_linear", p(dZ), N, p(W), K, 1, None, M, K, N, p(dX), K, s())
    assert_close(dX, dZr @ Wr, 2e-5, "dX vs rounded operands")
    assert_close(dX, dZ.double() @ W.double(), 2e-3, "dX vs exact")
    if M <= 128:
        dW = torch.full((N, K), float("nan"), device="cuda"); db = torch.full((N,), float("nan"), device="cuda")
        call("wf_rowmlp_dw", p(dZ), N, p(X), K, M, N, K, p(dW), K, p(db), s())
        assert_close(dW, dZr.t() @ Xr, 2e-5, "dW vs rounded operands")
        assert_close(dW, dZ.double().t() @ X.double(), 2e-3, "dW vs exact")
        assert_close(db, dZ.double().sum(0), 1e-6, "db")
        dW2 = torch.empty_like(dW)
        call("wf_rowmlp_dw", p(dZ), N, p(X), K, M, N, K, p(dW2), K, None, s())
        assert torch.equal(dW, dW2), "dW not reproducible"


def test_heads_route_through_rowmlp_and_match_the_tf32_path(ops):
    """ops.linear_ln_act at 64 rows in the production precision: the rowmlp kernels (default) against the split-K kind::tf32
    launches they replace (WF_B200_ROWMLP=0 behaviour, ops.USE_ROWMLP) -- same precision class, forward and all gradients."""
    torch.manual_seed(11)
    x = torch.randn(64, 1024, device="cuda"); W = torch.randn(2048, 1024, device="cuda") / 32; b = 0.1 * torch.randn(2048, device="cuda")
    g = 1 + 0.1 * torch.randn(2048, device="cuda"); be = 0.1 * torch.randn(2048, device="cuda"); go = torch.randn(64, 2048, device="cuda")
    res = {}
    old = ops.USE_ROWMLP
    try:
        for flag in (True, False):
            ops.USE_ROWMLP = flag
            leaves = [t.clone().requires_grad_(True) for t in (x, W, b, g, be)]
            n0 = ops.LAUNCHES
            out = ops.linear_ln_act(leaves[0], leaves[1], leaves[2], leaves[3], leaves[4], 1)
            out.backward(go)
            res[flag] = (out.detach(), [t.grad for t in leaves], ops.LAUNCHES - n0)
    finally:
        ops.USE_ROWMLP = old
    assert res[True][2] < res[False][2], "the rowmlp route should need fewer launches"
    assert_close(res[True][0], res[False][0], 2e-3, "forward")
    for n, a, r in zip("x W b gamma beta".split(), res[True][1], res[False][1]):
        assert_close(a, r, 3e-3, f"grad {n}")


@pytest.mark.parametrize("C", [512, 1024, 2048])
def test_ln_relu_bf16_kernels(ops, C):
    """wf_ln_relu_bf16_fwd/bwd against torch autograd (fp64) on the same bf16-rounded inputs and the same row statistics."""
    import torch.nn.functional as F
    from wf_b200._lib import call
    torch.manual_seed(C)
    M = 1003
    z = (torch.randn(M, C, device="cuda") * 1.5 + 0.3).to(torch.bfloat16)
    dh = torch.randn(M, C, device="cuda").to(torch.bfloat16)
    g = (1 + 0.1 * torch.randn(C, device="cuda")); be = 0.1 * torch.randn(C, device="cuda")
    zd = z.double().requires_grad_(True); gd = g.double().requires_grad_(True); bd = be.double().requires_grad_(True)
    mean = zd.detach().mean(1); var = zd.detach().var(1, unbiased=False); rstd = (var + 1e-5).rsqrt()
    ref = torch.relu(F.layer_norm(zd, (C,), gd, bd, 1e-5))
    ref.backward(dh.double())
    m32, r32 = mean.float().contiguous(), rstd.float().contiguous()
    h = torch.empty_like(z)
    call("wf_ln_relu_bf16_fwd", ops._p(z), ops._p(m32), ops._p(r32), ops._p(g), ops._p(be), ops._p(h), M, C, ops._s())
    assert_close(h.float(), ref, 6e-3, "ln fwd bf16")
    dz = torch.empty_like(z); dg = torch.zeros(C, device="cuda"); dbe = torch.zeros(C, device="cuda"); db = torch.zeros(C, device="cuda")
    call("wf_ln_relu_bf16_bwd", ops._p(dh), ops._p(z), ops._p(m32), ops._p(r32), ops._p(g), ops._p(be), ops._p(dz), ops._p(dg),
         ops._p(dbe), ops._p(db), M, C, ops._s())
    assert_close(dz.float(), zd.grad, 8e-3, "ln bwd dz")
    assert_close(dg, gd.grad, 1e-4, "ln bwd dgamma")
    assert_close(dbe, bd.grad, 1e-4, "ln bwd dbeta")
    assert_close(db, zd.grad.sum(0), 2e-3, "ln bwd colsum(dz)")


@pytest.mark.parametrize("tA,tB", [(False, True), (False, False), (True, False), (True, True)])
def test_gemm_tf32_all_layouts(ops, tA, tB):
    """wf_gemm_tf32 (tcgen05 kind::tf32 on fp32 storage) in the four operand layouts the heads use, against fp64.
    TF32 keeps 10 mantissa bits per operand: scale-relative tolerance 2e-3."""
    from wf_b200._lib import call
    torch.manual_seed(11)
    for (M, N, K) in ((64, 4096, 512), (300, 520, 136), (2750, 1536, 512), (61000, 256, 512), (129, 260, 36)):
        A = torch.randn((K, M) if tA else (M, K), device="cuda")
        B = torch.randn((N, K) if tB else (K, N), device="cuda") / math.sqrt(K)
        if A.stride(0) % 4 or B.stride(0) % 4:
            continue
        bias = torch.randn(N, device="cuda")
        ref = (A.t() if tA else A).double() @ (B.t() if tB else B).double() + bias.double()
        out = torch.full((M, N), float("nan"), device="cuda")
        call("wf_gemm_tf32", ops._p(A), A.stride(0), int(not tA), ops._p(B), B.stride(0), int(tB), M, N, K, ops._p(bias),
             ops._p(out), out.stride(0), 0, 1, ops._s())
        assert_close(out, ref, 2e-3, f"tf32 tA={tA} tB={tB} {M}x{N}x{K}")
    # split-K accumulate (weight-gradient shape)
    if tA and not tB:
        A = torch.randn(61000, 512, device="cuda"); B = torch.randn(61000, 256, device="cuda") / 250
        out = torch.zeros(512, 256, device="cuda")
        call("wf_gemm_tf32", ops._p(A), 512, 0, ops._p(B), 256, 0, 512, 256, 61000, None, ops._p(out), 256, 1, 37, ops._s())
        assert_close(out, A.double().t() @ B.double(), 2e-3, "tf32 split-K dW")


def test_heads_tf32_dispatch(ops):
    """ops.gemm_f32 routes big aligned products to TF32 tensor cores in production precision and to the fp32 SIMT
    kernel in fp32 mode; both agree within the TF32 tolerance."""
    torch.manual_seed(12)
    x = torch.randn(64, 512, device="cuda"); W = torch.randn(4096, 512, device="cuda") / 22
    ops.set_precision("fp32"); a = ops.gemm_f32(x, W, transB=True)
    ops.set_precision("bf16"); b = ops.gemm_f32(x, W, transB=True)
    assert_close(a, x.double() @ W.double().t(), 1e-5, "simt")
    assert_close(b, a, 2e-3, "tf32 vs simt")
    assert not torch.equal(a, b)              # i.e. the tensor-core path really ran


# ----------------------------------------------------------------------------------------------
# fused pooling: max pools in the GEMM epilogue, mean pools through the affine map, analytic backward
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("masking", ["random", "all_valid", "padded_tail"])
@pytest.mark.parametrize("B,N,K,C", [(3, 300, 256, 512), (2, 128, 64, 96), (5, 1000, 1024, 512), (1, 4099, 128, 200)])
def test_gemm_bf16_pool_epilogue_bit_exact(ops, B, N, K, C, masking):
    """wf_gemm_bf16_pool: max and FIRST argmax per (cloud, channel), all rows and valid rows, must be bit-identical to
    reducing the stored fp32 output of wf_gemm_bf16 on the same operands (tile rows straddle cloud boundaries here:
    N is not a multiple of 128/32).  Also in row chunks (row_offset) as the inference path calls it.  Masking: random
    (the two-kind walk), all points valid and a zero-padded tail per cloud (the one-chain paths of fully valid / fully
    invalid 32-row groups)."""
    torch.manual_seed(B * 1000 + N)
    M = B * N
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    W = (torch.randn(C, K, device="cuda") / math.sqrt(K)).to(torch.bfloat16)
    # force ties: duplicate rows (same value in every channel) so that "first index" is exercised
    A[N // 2] = A[3]; A[N - 1] = A[3]
    bias = torch.randn(C, device="cuda")
    if masking == "random":
        mask = (torch.rand(B, N, device="cuda") > 0.3)
        mask[:, 3] = False
        if B > 1:
            mask[1] = False                               # a cloud without valid points
    elif masking == "all_valid":
        mask = torch.ones(B, N, device="cuda", dtype=torch.bool)
    else:
        mask = torch.ones(B, N, device="cuda", dtype=torch.bool)
        mask[:, N - max(1, (2 * N) // 5):] = False        # the last 40 % of every cloud is padding
    mask_u8 = mask.to(torch.uint8).contiguous()
    pf = torch.empty(M, C, device="cuda")
    ops.gemm_bf16(A, W, M=M, N=C, K=K, bias=bias, out=pf)
    pf3 = pf.view(B, N, C)
    ref_u, ref_au = pf3.max(dim=1)
    pm = pf3.masked_fill(~mask.unsqueeze(-1), float("-inf"))
    ref_m, ref_am = pm.max(dim=1)
    # torch.max returns *an* index of the maximum; compute the first one explicitly
    ar = torch.arange(N, device="cuda").view(1, N, 1)
    first = lambda t, mx: torch.where(t == mx.unsqueeze(1), ar, N).min(dim=1).values
    ref_au, ref_am = first(pf3, ref_u), first(pm, ref_m)
    for chunks in (1, 3):
        packed = torch.zeros(2, B, C, device="cuda", dtype=torch.int64)
        step = ((M + chunks - 1) // chunks + 127) // 128 * 128
        for r0 in range(0, M, step):
            r1 = min(M, r0 + step)
            ops.gemm_bf16_pool(A[r0:r1], W, M=r1 - r0, N=C, K=K, bias=bias, points_per_cloud=N, row_offset=r0,
                               mask=mask_u8.view(-1)[r0:r1], packed=packed)
        outs = [torch.empty(B, C, device="cuda") for _ in range(4)]
        args = [torch.empty(B, C, device="cuda", dtype=torch.int32) for _ in range(2)]
        lin = torch.zeros(2 * B, C, device="cuda")
        from wf_b200._lib import call
        call("wf_pool_finalize", ops._p(packed[0]), ops._p(packed[1]), ops._p(lin), None, B, C, ops._p(outs[0]), ops._p(args[0]),
             ops._p(outs[1]), ops._p(outs[2]), ops._p(args[1]), ops._p(outs[3]), ops._s())
        assert torch.equal(outs[2], ref_u), "unmasked max not bit-identical"
        assert torch.equal(args[1].long(), ref_au), "unmasked argmax differs"
        has = mask.any(dim=1)
        exp_m = torch.where(has.unsqueeze(1), ref_m, torch.zeros_like(ref_m))
        exp_am = torch.where(has.unsqueeze(1), ref_am, torch.full_like(ref_am, -1))
        assert torch.equal(outs[0], exp_m), "masked max not bit-identical"
        assert torch.equal(args[0].long(), exp_am), "masked argmax differs"


@pytest.mark.parametrize("B,N", [(3, 333), (2, 20011)])
def test_ln_colsum_and_seg_mean(ops, B, N):
    """wf_ln_relu_bf16_fwd_colsum writes the same h as wf_ln_relu_bf16_fwd and wf_seg_mean returns the per-cloud means of
    that h (all rows / valid rows) -- also when called in row chunks.  The 20 011-point clouds have 157 row blocks each: the
    interleaved ranges of wf_seg_mean take several blocks per thread, and the cloud boundary falls inside a block."""
    from wf_b200._lib import call
    torch.manual_seed(11)
    C = 1024
    M = B * N
    z = (torch.randn(M, C, device="cuda") * 1.5).to(torch.bfloat16)
    mean = z.float().mean(1).contiguous(); rstd = (z.float().var(1, unbiased=False) + 1e-5).rsqrt().contiguous()
    g = 1 + 0.1 * torch.randn(C, device="cuda"); be = 0.1 * torch.randn(C, device="cuda")
    mask = (torch.rand(M, device="cuda") > 0.25).to(torch.uint8)
    valid = mask.view(B, N).sum(1).clamp_min(1).float()
    h_ref = torch.empty_like(z)
    call("wf_ln_relu_bf16_fwd", ops._p(z), ops._p(mean), ops._p(rstd), ops._p(g), ops._p(be), ops._p(h_ref), M, C, ops._s())
    for step in (M, 384):
        h = torch.zeros_like(z)
        part = torch.empty(call("wf_seg_part_floats", M, C), device="cuda")
        for r0 in range(0, M, step):
            n = min(step, M - r0)
            call("wf_ln_relu_bf16_fwd_colsum", ops._p(z[r0:]), ops._p(mean[r0:]), ops._p(rstd[r0:]), ops._p(g), ops._p(be),
                 ops._p(h[r0:]), ops._p(mask[r0:]), n, C, N, r0, ops._p(part), ops._s())
        assert torch.equal(h, h_ref)
        hbar = torch.empty(2 * B, C, device="cuda")
        call("wf_seg_mean", ops._p(part), ops._p(valid), B, N, C, ops._p(hbar), ops._s())
        h3 = h_ref.float().view(B, N, C)
        assert_close(hbar[:B], h3.mean(1), 1e-5, "mean of h")
        assert_close(hbar[B:], (h3 * mask.view(B, N, 1)).sum(1) / valid.view(B, 1), 1e-5, "masked mean of h")


def test_encoder_fused_pool_vs_unfused(ops):
    """The whole bf16 encoder with the fused pooling against the same tensor-core path that materialises the fp32 point
    features (WF_B200_FUSED_POOL=0 behaviour): max pools and argmax bit-identical, mean pools to fp32 rounding, and
    all 18 parameter gradients close (the analytic backward keeps the pooled gradient in fp32 where the dense path rounds
    it to bf16, so agreement is at bf16 level, not bitwise)."""
    from oracle import wireframe_oracle as wo
    from models.PointNetEncoder import PointNetEncoder
    torch.manual_seed(0)
    enc = PointNetEncoder().cuda()
    sd = {k[len("encoder."):]: v for k, v in wo.make_state_dict(5, 16).items() if k.startswith("encoder.")}
    enc.load_state_dict(sd)
    x, _, _ = wo.make_inputs(7, 3, 900, 16, pad_frac=0.15, norm_intensity=True)
    x = x.cuda()
    gs = [torch.randn(3, 512, device="cuda") for _ in range(4)]
    ops.set_precision("bf16")
    res = {}
    for fused in (False, True):
        ops.FUSED_POOL = fused
        enc.zero_grad()
        r = enc.pooled(x)
        (r[0] * gs[0] + r[1] * gs[1] + r[2] * gs[2] + r[3] * gs[3]).sum().backward()
        res[fused] = ([t.detach().clone() for t in r[:6]], {k: p.grad.detach().clone() for k, p in enc.named_parameters()
                                                            if p.grad is not None})
    ops.FUSED_POOL = True
    a, b = res[True][0], res[False][0]
    assert torch.equal(a[0], b[0]) and torch.equal(a[4], b[4]), "masked max / argmax"
    assert torch.equal(a[2], b[2]) and torch.equal(a[5], b[5]), "unmasked max / argmax"
    assert_close(a[1], b[1], 2e-3, "masked mean pool")
    assert_close(a[3], b[3], 2e-3, "unmasked mean pool")
    for k, g in res[False][1].items():
        if k.startswith("mlp."):
            e = _fro(res[True][1][k], g)
            print(f"   fused vs unfused {k:16s} err {e:.3e}")
            assert e < 3e-2, f"grad {k}: {e:.3e}"


def test_pool_fused_bwd_matches_dense(ops):
    """wf_pool_fused_bwd (analytic) against the dense definition in fp64: dh = d_pf W, dW = d_pf^T h, db = colsum(d_pf)
    with d_pf built from the four pooled gradients exactly as autograd would (reference train.py:140)."""
    from wf_b200._lib import call
    torch.manual_seed(3)
    B, N, C, K = 3, 257, 512, 1024
    M = B * N
    h = torch.relu(torch.randn(M, K, device="cuda")).to(torch.bfloat16)
    W = torch.randn(C, K, device="cuda") / math.sqrt(K)
    mask = (torch.rand(B, N, device="cuda") > 0.2)
    mask[2] = False
    valid = mask.sum(1).clamp_min(1).float()
    arg_m = torch.randint(0, N, (B, C), device="cuda", dtype=torch.int32)
    arg_m[:, :40] = 7                                      # many channels share one row
    arg_m[2] = -1
    arg_u = torch.randint(0, N, (B, C), device="cuda", dtype=torch.int32)
    arg_u[:, :20] = 7
    g = [torch.randn(B, C, device="cuda") for _ in range(4)]       # g_max_m, g_avg_m, g_max_u, g_mean_u
    # dense reference
    d_pf = torch.zeros(B, N, C, device="cuda", dtype=torch.float64)
    d_pf += (g[3].double() / N).unsqueeze(1)
    d_pf += mask.unsqueeze(-1).double() * (g[1].double() / valid.double().unsqueeze(1)).unsqueeze(1)
    bi = torch.arange(B, device="cuda").unsqueeze(1).expand(B, C); ci = torch.arange(C, device="cuda").unsqueeze(0).expand(B, C)
    okm = arg_m >= 0
    d_pf.index_put_((bi[okm], arg_m[okm].long(), ci[okm]), g[0].double()[okm], accumulate=True)
    d_pf.index_put_((bi, arg_u.long(), ci), g[2].double(), accumulate=True)
    d2 = d_pf.view(M, C)
    dh_ref = d2 @ W.double(); dW_ref = d2.t() @ h.double(); db_ref = d2.sum(0)
    # library
    hbar = torch.empty(2 * B, K, device="cuda")
    h3 = h.float().view(B, N, K)
    hbar[:B] = h3.mean(1); hbar[B:] = (h3 * mask.unsqueeze(-1)).sum(1) / valid.unsqueeze(1)
    G = torch.cat([g[3], g[1]], 0).contiguous()
    dbar = ops.gemm_f32(G, W, tc=False)
    dW = ops.gemm_f32(G, hbar, transA=True, tc=False)
    db = torch.empty(C, device="cuda")
    work = torch.empty(call("wf_pool_fused_bwd_work_ints", B, C), device="cuda", dtype=torch.int32)
    dh = torch.empty(M, K, device="cuda", dtype=torch.bfloat16)
    mask_u8 = mask.to(torch.uint8).contiguous()
    call("wf_pool_fused_bwd", ops._p(g[0]), ops._p(g[1]), ops._p(g[2]), ops._p(g[3]), ops._p(arg_m), ops._p(arg_u), ops._p(mask_u8),
         ops._p(valid), ops._p(dbar), ops._p(W), ops._p(h), B, N, C, K, ops._p(work), ops._p(dh), ops._p(dW), ops._p(db), ops._s())
    assert_close(dh.float(), dh_ref, 5e-3, "dh (bf16 storage)")
    assert_close(dW, dW_ref, 2e-5, "dW")
    # masked-avg term of db: the library's rule is "cloud has a valid point" (arg_m >= 0), the dense one sums the mask
    assert_close(db, db_ref, 2e-5, "db")


def _enc_and_input(seed, B, N, pad):
    from oracle import wireframe_oracle as wo
    from models.PointNetEncoder import PointNetEncoder
    torch.manual_seed(0)
    enc = PointNetEncoder().cuda()
    sd = {k[len("encoder."):]: v for k, v in wo.make_state_dict(seed, 16).items() if k.startswith("encoder.")}
    enc.load_state_dict(sd)
    x, _, _ = wo.make_inputs(seed + 1, B, N, 16, pad_frac=pad, norm_intensity=True)
    return enc, x.cuda()


def test_encoder_inference_chunked_is_bit_identical(ops):
    """The no_grad encoder (chunked over points, three reused buffers) must return the same bits as the training forward:
    same kernels per row, packed maxima are order-independent, column sums are added per 128-row block in block order."""
    enc, x = _enc_and_input(9, 3, 1000, 0.1)
    ops.set_precision("bf16")
    ref = [t.detach().clone() for t in enc.pooled(x)[:6]]          # autograd path (grad enabled)
    for chunk in (384, 1280, 1 << 19):
        old = ops.INFER_CHUNK_ROWS
        ops.INFER_CHUNK_ROWS = chunk
        try:
            with torch.no_grad():
                out = enc.pooled(x)[:6]
        finally:
            ops.INFER_CHUNK_ROWS = old
        for a, b, n in zip(out, ref, ("max_m", "avg_m", "max_u", "mean_u", "arg_m", "arg_u")):
            assert torch.equal(a, b), f"chunk {chunk}: {n} differs"
    # the L2-resident form: small chunks through two ping-pong buffers, LayerNorm in place (ops.encoder_pooled_infer)
    xin, params = enc.tc_inputs(x)
    with torch.no_grad():
        out = ops.encoder_pooled_infer(xin, [t.detach() for t in params], l2_resident=True, chunk_rows=512)
    for a, b, n in zip(out, ref, ("max_m", "avg_m", "max_u", "mean_u", "arg_m", "arg_u")):
        assert torch.equal(a, b), f"L2-resident form: {n} differs"


def test_encoder_point_sharded_matches_unsharded(ops):
    """SURVEY 8e config 4 (batch < world): clouds split by points over 2 'ranks' (emulated on one device: the two shards'
    partial pools are combined exactly as reduce_pool_shards' MAX / SUM all-reduces would).  Max pools and GLOBAL argmax
    indices must be bit-identical to the unsharded encoder, mean pools equal to fp32 summation order."""
    enc, x = _enc_and_input(4, 2, 1024, 0.1)
    ops.set_precision("bf16")
    with torch.no_grad():
        ref = enc.pooled(x)[:6]
    p = []
    for li in range(4):
        lin, ln = enc.mlp[4 * li], enc.mlp[4 * li + 1]
        p += [lin.weight, lin.bias, ln.weight, ln.bias]
    p += [enc.mlp[16].weight, enc.mlp[16].bias]
    world, n = 2, 512
    shards = [x[:, r * n:(r + 1) * n].contiguous() for r in range(world)]
    stash = []
    for r in range(world):                                   # pass 1: collect each rank's partial pools
        ops.encoder_pooled_infer(shards[r], p, index_offset=r * n, points_total=world * n,
                                 reduce_fn=lambda a, b, c: stash.append((a.clone(), b.clone(), c.clone())))
    flip = -(1 << 63)                                        # the words are unsigned: compare in unsigned order
    packed = torch.maximum(stash[0][0] ^ flip, stash[1][0] ^ flip) ^ flip; hsum = stash[0][1] + stash[1][1]; cnt = stash[0][2] + stash[1][2]

    def fake_allreduce(a, b, c):
        a.copy_(packed); b.copy_(hsum); c.copy_(cnt)
    for r in range(world):                                   # pass 2: every rank finalises the combined pools
        out = ops.encoder_pooled_infer(shards[r], p, index_offset=r * n, points_total=world * n, reduce_fn=fake_allreduce)
        assert torch.equal(out[0], ref[0]) and torch.equal(out[4], ref[4]), "masked max / global argmax"
        assert torch.equal(out[2], ref[2]) and torch.equal(out[5], ref[5]), "unmasked max / global argmax"
        assert_close(out[1], ref[1], 1e-5, "masked mean pool")
        assert_close(out[3], ref[3], 1e-5, "unmasked mean pool")


@pytest.mark.parametrize("M,N,K,tB", [(64, 2048, 4096, True), (64, 512, 4096, False), (64, 4096, 512, True), (100, 1024, 2048, True)])
def test_tf32_split_k_forward_is_deterministic(ops, M, N, K, tB, monkeypatch):
    """Few-tile TF32 products split the reduction over K; forward/dX products must stay bit-reproducible: partial tiles go to
    workspace slices that are added in a fixed order (wf_gemm_tf32_splitk).  Products with <= 128 rows take the rowmlp kernels
    by default (test_rowmlp_kernels); this test pins the kind::tf32 route, which keeps serving taller few-tile products."""
    monkeypatch.setattr(ops, "USE_ROWMLP", False)
    torch.manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda")
    B = torch.randn(N, K, device="cuda") / math.sqrt(K) if tB else torch.randn(K, N, device="cuda") / math.sqrt(K)
    bias = torch.randn(N, device="cuda")
    ops.set_precision("bf16")
    l0 = ops.LAUNCHES
    outs = [ops.gemm_f32(A, B, transB=tB, bias=bias) for _ in range(3)]
    assert ops.LAUNCHES - l0 == 6, "expected the split-K pair of launches (GEMM + ordered reduction)"
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    ref = A.double() @ (B.double().t() if tB else B.double()) + bias.double()
    assert_close(outs[0], ref, 2e-3, "tf32 split-K product")
    acc = outs[0].clone()
    ops.gemm_f32(A, B, transB=tB, out=acc, beta=1.0)
    assert_close(acc, 2 * ref - bias.double(), 2e-3, "tf32 split-K accumulate (beta = 1)")
