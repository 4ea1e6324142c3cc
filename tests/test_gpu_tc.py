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
    stats = torch.zeros(M, 2, device="cuda")
    ops.gemm_bf16(A, B, M=M, N=N, K=K, bias=bias, out=out, rowstats=stats)
    ref = A.float().double() @ B.float().double().t() + bias.double()
    assert_close(out, ref, 1e-4, f"kmajor f32 {M}x{N}x{K}")
    assert_close(stats[:, 0], ref.sum(1), 1e-4, "rowstats sum")
    assert_close(stats[:, 1], (ref * ref).sum(1), 1e-4, "rowstats sumsq")


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
    assert_close(dW2, d[1].grad, 5e-5, "l1 bwd dW (no dx)")


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
