"""GPU probe, sustained (power-capped) regime: GEMM alone, LayerNorm alone and GEMM + LayerNorm side job, each run back to
back for ~1.5 s without flushes; reports ms per iteration and the board power seen by NVML."""
import os, sys, json, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "wireframe-3d-prediction_b200")); sys.path.insert(0, ROOT)
from wf_b200 import ops
from wf_b200.ops import call, _p, _s
import pynvml
pynvml.nvmlInit(); H = pynvml.nvmlDeviceGetHandleByIndex(0)
dev = "cuda"
M = 320000


def sustained(fn, secs=1.5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    # calibrate
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); [fn() for _ in range(10)]; e1.record(); torch.cuda.synchronize()
    per = e0.elapsed_time(e1) / 10
    n = max(20, int(secs * 1e3 / per))
    pw, ck = [], []
    e0.record()
    for i in range(n):
        fn()
        if i % max(1, n // 20) == 0 and i > n // 3:
            pw.append(pynvml.nvmlDeviceGetPowerUsage(H) / 1e3); ck.append(pynvml.nvmlDeviceGetClockInfo(H, pynvml.NVML_CLOCK_SM))
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, sum(pw) / max(1, len(pw)), sum(ck) / max(1, len(ck))


def ln_in(Ms, C):
    z = torch.randn(Ms, C, device=dev).to(torch.bfloat16)
    mean = z.float().mean(1); rstd = (z.float().var(1, unbiased=False) + 1e-5).rsqrt()
    return z, mean, rstd, torch.ones(C, device=dev), torch.zeros(C, device=dev)


for (N, K, Ms, Cs) in [(2048, 1024, 160000, 1024), (2048, 1024, 320000, 1024), (1024, 2048, 160000, 2048)]:
    A = torch.randn(M, K, device=dev).to(torch.bfloat16); W = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    bias = torch.zeros(N, device=dev); out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    st = torch.empty(call("wf_gemm_rowstats_parts", N), M, 2, device=dev)
    z, mean, rstd, gamma, beta = ln_in(Ms, Cs); h = torch.empty_like(z)
    dh = torch.randn(Ms, Cs, device=dev).to(torch.bfloat16); dz = torch.empty_like(z)
    dg, db, dc = (torch.zeros(Cs, device=dev) for _ in range(3))
    g = lambda: ops.gemm_bf16(A, W, M=M, N=N, K=K, bias=bias, out=out, rowstats=st)
    l = lambda: call("wf_ln_relu_bf16_fwd", _p(z), _p(mean), _p(rstd), _p(gamma), _p(beta), _p(h), Ms, Cs, _s())
    lb = lambda: call("wf_ln_relu_bf16_bwd", _p(dh), _p(z), _p(mean), _p(rstd), _p(gamma), _p(beta), _p(dz), _p(dg), _p(db), _p(dc), Ms, Cs, _s())
    seg = [ops.side_ln_fwd(z, mean, rstd, gamma, beta, h, 0, Ms)]
    segb = [ops.side_ln_bwd(dh, z, mean, rstd, gamma, beta, dz, dg, db, dc, 0, Ms)]
    s = lambda: ops.gemm_bf16(A, W, M=M, N=N, K=K, bias=bias, out=out, rowstats=st, side=seg)
    sb = lambda: ops.gemm_bf16(A, W, M=M, N=N, K=K, bias=bias, out=out, rowstats=st, side=segb)
    both = lambda: (g(), l())
    bothb = lambda: (g(), lb())
    r = {"gemm": f"{M}x{N}x{K}", "ln": f"{Ms}x{Cs}"}
    for name, fn in (("gemm", g), ("ln_fwd", l), ("ln_bwd", lb), ("gemm_then_ln_fwd", both), ("gemm_with_side_fwd", s), ("gemm_then_ln_bwd", bothb), ("gemm_with_side_bwd", sb)):
        ms, p, c = sustained(fn)
        r[name] = {"ms": round(ms, 4), "watts": round(p), "sm_mhz": round(c)}
    print(json.dumps(r), flush=True)
    del A, W, out, z, h, dh, dz
