"""The pooling launch alone (final per-point Linear 1024 -> 512 with the max-pool epilogue) at the bench shape: target for
  ncu --set full --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 -o /tmp/pool python tools/pool_probe.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "wireframe-3d-prediction_b200"))
from wf_b200 import ops  # noqa: E402

M, K, N, clouds = 640000, 1024, 512, 64
A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
W = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
bias = torch.randn(N, device="cuda")
mask = torch.ones(M, device="cuda", dtype=torch.uint8)
packed = torch.zeros(2, clouds, N, device="cuda", dtype=torch.int64)
for _ in range(4):
    ops.gemm_bf16_pool(A, W, M=M, N=N, K=K, bias=bias, points_per_cloud=M // clouds, row_offset=0, mask=mask, packed=packed)
torch.cuda.synchronize()
print("done")
