"""One forward+backward of the bf16 encoder at the bench shape (64 clouds x 10,000 points) -- the 10 tcgen05 GEMM launches of a
training step (4 fwd incl. the pooling epilogue, 3 dX, 3 dW).  Target for
  ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 10 -c 10 -o gpurun_out/prof_gemm python tools/prof_gemm.py
(the first 10 matching launches are the warm-up pass)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "wireframe-3d-prediction_b200"), ROOT]
import torch  # noqa: E402
from wf_b200 import ops  # noqa: E402
from wf_b200.synthetic import make_inputs  # noqa: E402
from models.PointNetEncoder import PointNetEncoder  # noqa: E402

torch.manual_seed(0)
enc = PointNetEncoder().cuda()
x, _, _ = make_inputs(seed=0, B=64, N=10000, V=64, min_count=16, max_count=64)
x = x.cuda()
g = [torch.randn(64, 512, device="cuda") for _ in range(4)]
ops.set_precision("bf16")
for it in range(2):
    enc.zero_grad()
    r = enc.pooled(x)
    (r[0] * g[0] + r[1] * g[1] + r[2] * g[2] + r[3] * g[3]).sum().backward()
    torch.cuda.synchronize()
print("done; kernels launched:", ops.LAUNCHES)
