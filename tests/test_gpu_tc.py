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


def test_encoder_tc_vs_fp32_path(ops):
    """bf16 tensor-core encoder (fwd pools + all parameter gradients) against the fp32 SIMT path of this
    library on the same weights.  Tolerance: bf16 operands, fp32 accumulation, 4 layers deep."""
    from oracle import wireframe_oracle as wo
    from models.PointNetEncoder import PointNetEncoder
    torch.manual_seed(0)
    enc = PointNetEncoder().cuda()
    sd = {k[len("encoder."):]: v for k, v in wo.make_state_dict(21, 16).items() if k.startswith("encoder.")}
    enc.load_state_dict(sd)
    x, _, _ = wo.make_inputs(3, 2, 700, 16, pad_frac=0.1, norm_intensity=True)
    x = x.cuda()
    gs = [torch.randn(2, 512, device="cuda") for _ in range(4)]
    res = {}
    for prec in ("fp32", "bf16"):
        ops.set_precision(prec)
        enc.zero_grad()
        r = enc.pooled(x)
        (r[0] * gs[0] + r[1] * gs[1] + r[2] * gs[2] + r[3] * gs[3]).sum().backward()
        res[prec] = ([t.detach().clone() for t in r[:4]], {k: p.grad.detach().clone() for k, p in enc.named_parameters() if p.grad is not None})
    ops.set_precision("bf16")
    for a, b, n in zip(res["bf16"][0], res["fp32"][0], ("max_m", "avg_m", "max_u", "mean_u")):
        assert_close(a, b, 3e-2, f"pooled {n}")
    worst = 0.0
    for k, g in res["fp32"][1].items():
        if k.startswith("mlp."):
            e = rel_err(res["bf16"][1][k], g)
            worst = max(worst, e)
            assert e < 8e-2, f"grad {k}: {e:.3e}"
    print("worst encoder grad rel err bf16 vs fp32:", worst)
