"""GPU probe: Linear + LayerNorm in one launch (own-output side job, z re-read through L2) against GEMM then LayerNorm,
burst and sustained; also DRAM bytes when run under ncu."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "wireframe-3d-prediction_b200")); sys.path.insert(0, ROOT)
from wf_b200 import ops
from wf_b200.ops import call, _p, _s
dev = "cuda"
M = int(os.environ.get("PROBE_M", 640000))
reps = int(os.environ.get("PROBE_REPS", 40))
for (N, K) in [(1024, 512), (2048, 1024)]:
    A = torch.randn(M, K, device=dev).to(torch.bfloat16); W = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    bias = torch.zeros(N, device=dev); z = torch.empty(M, N, device=dev, dtype=torch.bfloat16); h = torch.empty_like(z)
    parts = call("wf_gemm_rowstats_parts", N)
    st = torch.empty(parts, M, 2, device=dev); mean = torch.empty(M, device=dev); rstd = torch.empty(M, device=dev)
    gamma = torch.ones(N, device=dev); beta = torch.zeros(N, device=dev)

    def three():
        ops.gemm_bf16(A, W, M=M, N=N, K=K, bias=bias, out=z, rowstats=st)
        call("wf_stats_finalize", _p(st), M, N, parts, 1e-5, _p(mean), _p(rstd), _s())
        call("wf_ln_relu_bf16_fwd", _p(z), _p(mean), _p(rstd), _p(gamma), _p(beta), _p(h), M, N, _s())

    def own():
        ops.gemm_bf16_ownln(A, W, M=M, N=N, K=K, bias=bias, z=z, rowstats=st, gamma=gamma, beta=beta, h=h, mean=mean, rstd=rstd)

    res = {"shape": f"{M}x{N}x{K}"}
    for name, fn in (("three_kernels", three), ("own_ln", own)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        res[name + "_ms_sustained"] = round(e0.elapsed_time(e1) / reps, 4)
    print(json.dumps(res), flush=True)
    del A, W, z, h, st
