"""The K = 512 encoder layer alone (Linear 512 -> 1024 with bias, row statistics and the bf16 TMA-store epilogue) at the bench
shape: target for  ncu --set full --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 -o /tmp/k512 python tools/k512_probe.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "wireframe-3d-prediction_b200"))
from wf_b200 import ops  # noqa: E402
from wf_b200._lib import call  # noqa: E402

M, K, N = 640000, 512, 1024
A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
W = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
bias = torch.randn(N, device="cuda")
z = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
st = torch.empty(call("wf_gemm_rowstats_parts", N), M, 2, device="cuda")
for _ in range(4):
    ops.gemm_bf16(A, W, M=M, N=N, K=K, bias=bias, out=z, rowstats=st)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.gemm_bf16(A, W, M=M, N=N, K=K, bias=bias, out=z, rowstats=st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"K=512 layer: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TF/s")
