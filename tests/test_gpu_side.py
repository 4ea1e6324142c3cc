"""Side jobs (include/wf_b200.h, wf_side_seg): LayerNorm passes executed by spare warps inside the tcgen05 GEMM kernel.
The GEMM's own results must not change at all; the forward passes must be bit-identical to the stand-alone kernels (same
device code over 256 virtual threads); the backward pass agrees with the stand-alone kernel to rounding (its row sums are
added in a different order) and with an fp64 autograd reference."""
import pytest
import torch

from gpu_util import assert_close, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from wf_b200 import ops as o
    return o


def _ln_inputs(M, C, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    z = (torch.randn(M, C, device="cuda", generator=g) * 1.5 + 0.3).to(torch.bfloat16)
    zf = z.float()
    mean = zf.mean(1).contiguous()
    rstd = (zf.var(1, unbiased=False) + 1e-5).rsqrt().contiguous()
    gamma = (1.0 + 0.2 * torch.randn(C, device="cuda", generator=g)).contiguous()
    beta = (0.1 * torch.randn(C, device="cuda", generator=g)).contiguous()
    return z, mean, rstd, gamma, beta


def _gemm_inputs(M, N, K, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    W = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    return A, W, bias


@pytest.mark.parametrize("C", [1024, 2048])
@pytest.mark.parametrize("Mside", [128 * 37 + 5, 40000])
def test_side_ln_fwd_is_bit_identical_and_gemm_unchanged(ops, C, Mside):
    from wf_b200.ops import call, _p, _s
    M, N, K = 256 * 9 + 77, 1024, 512
    A, W, bias = _gemm_inputs(M, N, K, 1)
    parts = call("wf_gemm_rowstats_parts", N)
    out0 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16); st0 = torch.zeros(parts, M, 2, device="cuda")
    out1 = torch.empty_like(out0); st1 = torch.zeros_like(st0)
    ops.gemm_bf16(A, W, M=M, N=N, K=K, bias=bias, out=out0, rowstats=st0)
    z, mean, rstd, gamma, beta = _ln_inputs(Mside, C, 2)
    h_ref = torch.empty_like(z)
    call("wf_ln_relu_bf16_fwd", _p(z), _p(mean), _p(rstd), _p(gamma), _p(beta), _p(h_ref), Mside, C, _s())
    h = torch.full_like(z, float("nan"))
    # two segments: the pass is split at a row-block boundary like the scheduler does
    cut = (Mside // 2) // 128 * 128
    segs = [ops.side_ln_fwd(z, mean, rstd, gamma, beta, h, 0, cut), ops.side_ln_fwd(z, mean, rstd, gamma, beta, h, cut, Mside)]
    ops.gemm_bf16(A, W, M=M, N=N, K=K, bias=bias, out=out1, rowstats=st1, side=segs)
    torch.cuda.synchronize()
    assert torch.equal(out0, out1) and torch.equal(st0, st1), "the GEMM's own results changed"
    assert torch.equal(h.view(torch.int16), h_ref.view(torch.int16)), "side LayerNorm forward differs from the stand-alone kernel"


def test_side_ln_fwd_colsum_bit_identical_under_the_pool_gemm(ops):
    """The last LayerNorm (per-cloud column sums of h4, deterministic per row block) as a side job of the POOLING GEMM of
    another chunk; the packed maxima of the pooling epilogue must be unchanged."""
    from wf_b200.ops import call, _p, _s
    B, Np, C = 3, 1000, 1024
    Mside = B * Np
    z, mean, rstd, gamma, beta = _ln_inputs(Mside, C, 3)
    mask = (torch.rand(Mside, device="cuda") > 0.2).to(torch.uint8)
    row_base = 128 * 5                                         # the segment starts inside a larger row space
    pool_n = Np
    nfl = call("wf_seg_part_floats", row_base + Mside, C)
    part_ref = torch.zeros(nfl, device="cuda"); part = torch.zeros(nfl, device="cuda")
    h_ref = torch.empty_like(z); h = torch.empty_like(z)
    call("wf_ln_relu_bf16_fwd_colsum", _p(z), _p(mean), _p(rstd), _p(gamma), _p(beta), _p(h_ref), _p(mask), Mside, C, pool_n,
         row_base, _p(part_ref), _s())
    # host GEMM: pooling epilogue over 2 clouds x 640 points
    Mg, Ng, Kg = 2 * 640, 512, 1024
    A, W, bias = _gemm_inputs(Mg, Ng, Kg, 4)
    gmask = torch.ones(Mg, device="cuda", dtype=torch.uint8)
    pk0 = torch.zeros(2, 2, Ng, device="cuda", dtype=torch.int64); pk1 = torch.zeros_like(pk0)
    ops.gemm_bf16_pool(A, W, M=Mg, N=Ng, K=Kg, bias=bias, points_per_cloud=640, row_offset=0, mask=gmask, packed=pk0)
    seg = ops.side_ln_fwd(z, mean, rstd, gamma, beta, h, 0, Mside, colsum=(mask, part, pool_n, row_base))
    ops.gemm_bf16_pool(A, W, M=Mg, N=Ng, K=Kg, bias=bias, points_per_cloud=640, row_offset=0, mask=gmask, packed=pk1, side=[seg])
    torch.cuda.synchronize()
    assert torch.equal(pk0, pk1)
    assert torch.equal(h.view(torch.int16), h_ref.view(torch.int16))
    assert torch.equal(part, part_ref), "per-block column sums differ"


@pytest.mark.parametrize("C", [1024, 2048])
def test_side_ln_bwd_matches_standalone_and_fp64(ops, C):
    from wf_b200.ops import call, _p, _s
    Mside = 128 * 23 + 3
    z, mean, rstd, gamma, beta = _ln_inputs(Mside, C, 5)
    g = torch.Generator(device="cuda").manual_seed(6)
    dh = torch.randn(Mside, C, device="cuda", generator=g).to(torch.bfloat16)
    outs = []
    for mode in ("standalone", "side"):
        dz = torch.empty_like(z)
        dg, db, dc = (torch.zeros(C, device="cuda") for _ in range(3))
        if mode == "standalone":
            call("wf_ln_relu_bf16_bwd", _p(dh), _p(z), _p(mean), _p(rstd), _p(gamma), _p(beta), _p(dz), _p(dg), _p(db), _p(dc),
                 Mside, C, _s())
        else:
            # as a side job of a dW-shaped GEMM (MN-major operands, split-K accumulate)
            Mo, No, Kr = 512, 512, 4096
            gg = torch.Generator(device="cuda").manual_seed(7)
            At = torch.randn(Kr, Mo, device="cuda", generator=gg).to(torch.bfloat16)
            Bt = torch.randn(Kr, No, device="cuda", generator=gg).to(torch.bfloat16)
            o0 = torch.zeros(Mo, No, device="cuda"); o1 = torch.zeros(Mo, No, device="cuda")
            ops.gemm_bf16(At, Bt, M=Mo, N=No, K=Kr, kmajor=False, out=o0, accumulate=True, split_k=8)
            cut = 128 * 11
            segs = [ops.side_ln_bwd(dh, z, mean, rstd, gamma, beta, dz, dg, db, dc, 0, cut),
                    ops.side_ln_bwd(dh, z, mean, rstd, gamma, beta, dz, dg, db, dc, cut, Mside)]
            ops.gemm_bf16(At, Bt, M=Mo, N=No, K=Kr, kmajor=False, out=o1, accumulate=True, split_k=8, side=segs)
            assert_close(o1, o0, 1e-5, "dW GEMM under a side job (atomic split-K: order may differ)")
        outs.append((dz, dg, db, dc))
    torch.cuda.synchronize()
    (dz0, dg0, db0, dc0), (dz1, dg1, db1, dc1) = outs
    # bf16 outputs: identical except where the differently ordered row sums move a value across a rounding boundary
    diff = (dz0.float() - dz1.float()).abs()
    assert float((diff > 0).float().mean()) < 2e-3 and rel_err(dz1.float(), dz0.float()) < 8e-3
    assert_close(dg1, dg0, 1e-4, "dgamma"); assert_close(db1, db0, 1e-4, "dbeta"); assert_close(dc1, dc0, 2e-3, "dbias")
    # fp64 autograd of relu(LN(z)) with the SAME statistics semantics
    zd = z.double().requires_grad_(True); gd = gamma.double().requires_grad_(True); bd = beta.double().requires_grad_(True)
    y = torch.relu(torch.nn.functional.layer_norm(zd, (C,), gd, bd, 1e-5))
    y.backward(dh.double())
    assert_close(dz1.float(), zd.grad, 8e-3, "ln bwd dz (side)")
    assert_close(dg1, gd.grad, 1e-4, "ln bwd dgamma (side)")
    assert_close(db1, bd.grad, 1e-4, "ln bwd dbeta (side)")


@pytest.mark.parametrize("chunks", [1, 2, 3])
def test_encoder_with_side_jobs_equals_encoder_without(ops, chunks):
    """The whole tensor-core encoder (forward pools + all 18 parameter gradients) with its LayerNorm passes riding inside
    the GEMM launches of a 1/2/3-chunk pipeline, against the same encoder with stand-alone passes: the forward is
    bit-identical (pooled values AND argmax), the gradients agree to the rounding of the differently ordered row sums /
    atomics.  Inference (chunked, no grad) takes the same pipeline and must equal the training forward bit for bit."""
    from oracle import wireframe_oracle as wo
    from models.PointNetEncoder import PointNetEncoder
    torch.manual_seed(0)
    enc = PointNetEncoder().cuda()
    sd = {k[len("encoder."):]: v for k, v in wo.make_state_dict(21, 16).items() if k.startswith("encoder.")}
    enc.load_state_dict(sd)
    x, _, _ = wo.make_inputs(3, 3, 1100, 16, pad_frac=0.1, norm_intensity=True)
    x = x.cuda()
    gs = [torch.randn(3, 512, device="cuda") for _ in range(4)]
    saved = (ops.SIDE_JOBS, ops.SIDE_MIN_ROWS, ops.SIDE_CHUNKS)
    res = {}
    try:
        ops.set_precision("bf16")
        for mode in ("plain", "side"):
            ops.SIDE_JOBS, ops.SIDE_MIN_ROWS, ops.SIDE_CHUNKS = (False, saved[1], saved[2]) if mode == "plain" else (True, 0, chunks)
            enc.zero_grad()
            r = enc.pooled(x)
            sum((t * g).sum() for t, g in zip(r[:4], gs)).backward()
            with torch.no_grad():
                ri = enc.pooled(x)
            res[mode] = ([t.detach().clone() for t in r[:6]], {k: p.grad.detach().clone() for k, p in enc.named_parameters() if p.grad is not None},
                         [t.clone() for t in ri[:6]])
    finally:
        ops.SIDE_JOBS, ops.SIDE_MIN_ROWS, ops.SIDE_CHUNKS = saved
    for a, b in zip(res["plain"][0], res["side"][0]):
        assert torch.equal(a, b), "forward differs between stand-alone and side-job LayerNorm passes"
    for a, b in zip(res["side"][0], res["side"][2]):
        assert torch.equal(a, b), "inference pipeline differs from the training forward"
    for k, g in res["plain"][1].items():
        if k.startswith("mlp."):
            assert rel_err(res["side"][1][k], g) < 2e-3, k
            assert float((res["side"][1][k].double() - g.double()).norm() / g.double().norm()) < 5e-4, k


@pytest.mark.parametrize("M,N,K", [(256 * 5 + 77, 1024, 512), (256 * 40, 2048, 1024), (300000, 1024, 512), (40000, 2048, 1024)])
def test_own_output_layernorm_is_bit_identical(ops, M, N, K):
    """wf_gemm_bf16_ownln: Linear, row statistics and LayerNorm+ReLU in one launch -- the side warps normalise each 256-row
    unit once all its N tiles are stored, reading it back through L2.  Z, mean, rstd and H must equal the three-kernel
    sequence (wf_gemm_bf16 + wf_stats_finalize + wf_ln_relu_bf16_fwd) bit for bit, launch after launch (a missed
    dependency would show as stale rows)."""
    from wf_b200.ops import call, _p, _s
    A, W, bias = _gemm_inputs(M, N, K, 11)
    g = torch.Generator(device="cuda").manual_seed(12)
    gamma = (1.0 + 0.2 * torch.randn(N, device="cuda", generator=g)).contiguous()
    beta = (0.1 * torch.randn(N, device="cuda", generator=g)).contiguous()
    parts = call("wf_gemm_rowstats_parts", N)
    z0 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16); st0 = torch.empty(parts, M, 2, device="cuda")
    mean0 = torch.empty(M, device="cuda"); rstd0 = torch.empty(M, device="cuda"); h0 = torch.empty_like(z0)
    ops.gemm_bf16(A, W, M=M, N=N, K=K, bias=bias, out=z0, rowstats=st0)
    call("wf_stats_finalize", _p(st0), M, N, parts, 1e-5, _p(mean0), _p(rstd0), _s())
    call("wf_ln_relu_bf16_fwd", _p(z0), _p(mean0), _p(rstd0), _p(gamma), _p(beta), _p(h0), M, N, _s())
    for rep in range(3):
        z1 = torch.full_like(z0, float("nan")); st1 = torch.empty_like(st0)
        mean1 = torch.full_like(mean0, float("nan")); rstd1 = torch.full_like(rstd0, float("nan")); h1 = torch.full_like(z0, float("nan"))
        ops.gemm_bf16_ownln(A, W, M=M, N=N, K=K, bias=bias, z=z1, rowstats=st1, gamma=gamma, beta=beta, h=h1, mean=mean1, rstd=rstd1)
        torch.cuda.synchronize()
        assert torch.equal(z1.view(torch.int16), z0.view(torch.int16)), "Z differs"
        assert torch.equal(mean1, mean0) and torch.equal(rstd1, rstd0), "row statistics differ"
        bad = (h1.view(torch.int16) != h0.view(torch.int16)).any(dim=1).nonzero().flatten()
        assert bad.numel() == 0, f"H differs on {bad.numel()} rows (first {bad[:8].tolist()}) in repetition {rep}"
