"""GPU probe: does a LayerNorm side job inside the tensor-core GEMM cost time?  Times the encoder GEMM shapes at half the
bench's rows (320,000) alone and carrying LayerNorm segments of the other half."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "wireframe-3d-prediction_b200")); sys.path.insert(0, ROOT)
from wf_b200 import ops
from wf_b200.ops import call, _p, _s

M = int(os.environ.get("PROBE_M", 320000))
dev = "cuda"


def timeit(fn, n=6):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for i in range(n + 2):
        flush.fill_(1); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def ln_in(Ms, C):
    z = torch.randn(Ms, C, device=dev).to(torch.bfloat16)
    mean = z.float().mean(1); rstd = (z.float().var(1, unbiased=False) + 1e-5).rsqrt()
    gamma = torch.ones(C, device=dev); beta = torch.zeros(C, device=dev)
    return z, mean, rstd, gamma, beta


res = []
for (N, K, Cs, frac) in [(1024, 512, 1024, 1.0), (1024, 512, 1024, 0.5), (2048, 1024, 1024, 1.0), (2048, 1024, 2048, 1.0), (1024, 2048, 2048, 1.0),
                         (1024, 2048, 2048, 0.5), (512, 1024, 1024, 1.0)]:
    A = torch.randn(M, K, device=dev).to(torch.bfloat16); W = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    bias = torch.zeros(N, device=dev); out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    st = torch.empty(call("wf_gemm_rowstats_parts", N), M, 2, device=dev)
    Ms = int(M * frac) // 128 * 128
    z, mean, rstd, gamma, beta = ln_in(Ms, Cs); h = torch.empty_like(z)
    t_g = timeit(lambda: ops.gemm_bf16(A, W, M=M, N=N, K=K, bias=bias, out=out, rowstats=st))
    t_l = timeit(lambda: call("wf_ln_relu_bf16_fwd", _p(z), _p(mean), _p(rstd), _p(gamma), _p(beta), _p(h), Ms, Cs, _s()))
    seg = [ops.side_ln_fwd(z, mean, rstd, gamma, beta, h, 0, Ms)]
    t_s = timeit(lambda: ops.gemm_bf16(A, W, M=M, N=N, K=K, bias=bias, out=out, rowstats=st, side=seg))
    # empty-ish side job: measures the cost of the 5-stage ring + re-dealt registers alone
    seg0 = [ops.side_ln_fwd(z, mean, rstd, gamma, beta, h, 0, 128)]
    t_0 = timeit(lambda: ops.gemm_bf16(A, W, M=M, N=N, K=K, bias=bias, out=out, rowstats=st, side=seg0))
    r = {"what": "fwd", "gemm": f"{M}x{N}x{K}", "ln": f"{Ms}x{Cs}", "gemm_ms": round(t_g, 3), "ln_ms": round(t_l, 3), "side_ms": round(t_s, 3),
         "side_empty_ms": round(t_0, 3), "tflops_alone": round(2 * M * N * K / t_g / 1e9, 1), "tflops_with_side": round(2 * M * N * K / t_s / 1e9, 1),
         "hidden_frac": round((t_g + t_l - t_s) / t_l, 3), "ln_gbs_side": round(Ms * Cs * 4 / (t_s * 1e6), 1)}
    print(json.dumps(r), flush=True); res.append(r)
    del A, W, out, z, h

# backward: dX (K-major) and dW (MN-major, split-K) GEMMs carrying the LayerNorm backward
for (kind, N, K, Cs, frac) in [("dX", 1024, 2048, 1024, 1.0), ("dX", 1024, 2048, 2048, 0.5), ("dW", 2048, 1024, 2048, 1.0), ("dW", 2048, 1024, 1024, 1.0),
                               ("dX", 512, 1024, 1024, 0.5)]:
    Ms = int(M * frac) // 128 * 128
    z, mean, rstd, gamma, beta = ln_in(Ms, Cs)
    dh = torch.randn(Ms, Cs, device=dev).to(torch.bfloat16); dz = torch.empty_like(z)
    dg, db, dc = (torch.zeros(Cs, device=dev) for _ in range(3))
    if kind == "dX":
        A = torch.randn(M, K, device=dev).to(torch.bfloat16); W = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        run = lambda side=None: ops.gemm_bf16(A, W, M=M, N=N, K=K, out=out, side=side)
        flop = 2 * M * N * K
    else:
        A = torch.randn(M, N, device=dev).to(torch.bfloat16); Bm = torch.randn(M, K, device=dev).to(torch.bfloat16)
        out = torch.zeros(N, K, device=dev)
        run = lambda side=None: ops.gemm_bf16(A, Bm, M=N, N=K, K=M, kmajor=False, out=out, accumulate=True, split_k=37, side=side)
        flop = 2 * M * N * K
    t_g = timeit(lambda: run())
    t_l = timeit(lambda: call("wf_ln_relu_bf16_bwd", _p(dh), _p(z), _p(mean), _p(rstd), _p(gamma), _p(beta), _p(dz), _p(dg), _p(db), _p(dc), Ms, Cs, _s()))
    seg = [ops.side_ln_bwd(dh, z, mean, rstd, gamma, beta, dz, dg, db, dc, 0, Ms)]
    t_s = timeit(lambda: run(seg))
    r = {"what": "bwd " + kind, "gemm": f"{M}x{N}x{K}", "ln": f"{Ms}x{Cs}", "gemm_ms": round(t_g, 3), "ln_ms": round(t_l, 3), "side_ms": round(t_s, 3),
         "tflops_alone": round(flop / t_g / 1e9, 1), "tflops_with_side": round(flop / t_s / 1e9, 1),
         "hidden_frac": round((t_g + t_l - t_s) / t_l, 3), "ln_gbs_side": round(Ms * Cs * 6 / (t_s * 1e6), 1)}
    print(json.dumps(r), flush=True); res.append(r)
    del A, out, z, dh, dz
