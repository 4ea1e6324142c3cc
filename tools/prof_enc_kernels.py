"""Run the non-GEMM encoder kernels once at bench shapes (M = 64 x 10,000 points) -- target for `ncu --set full`:
  ncu --set full --clock-control none --import-source on -k regex:'l1c|ln_relu' -o gpurun_out/prof_enc python tools/prof_enc_kernels.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "wireframe-3d-prediction_b200"), ROOT]
import torch  # noqa: E402
from wf_b200 import ops  # noqa: E402
from wf_b200._lib import call, BF16  # noqa: E402

M = 640000
dev = "cuda"
torch.manual_seed(0)
x = torch.randn(M, 8, device=dev)
W = torch.randn(512, 8, device=dev) / 8 ** 0.5
b = 0.1 * torch.randn(512, device=dev); g = 1 + 0.1 * torch.randn(512, device=dev); be = 0.1 * torch.randn(512, device=dev)
h = torch.empty(M, 512, device=dev, dtype=torch.bfloat16)
dh = torch.randn(M, 512, device=dev).to(torch.bfloat16)
dW = torch.zeros(512, 8, device=dev); db = torch.zeros(512, device=dev); dg = torch.zeros(512, device=dev); dbe = torch.zeros(512, device=dev)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ev = lambda: torch.cuda.Event(enable_timing=True)


def timeit(name, fn, bytes_):
    fn(); torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:28s} {ms:8.3f} ms  {bytes_ / ms / 1e9:8.2f} TB/s-equivalent", flush=True)


timeit("l1 fwd", lambda: call("wf_enc_l1_fwd", ops._p(x), ops._p(W), ops._p(b), ops._p(g), ops._p(be), ops._p(h), BF16, M, 8, 512, 1e-5, ops._s()),
       M * (32 + 1024))
timeit("l1 bwd", lambda: call("wf_enc_l1_bwd", ops._p(x), ops._p(W), ops._p(b), ops._p(g), ops._p(be), ops._p(dh), BF16, ops._p(dW), ops._p(db),
                              ops._p(dg), ops._p(dbe), None, M, 8, 512, 1e-5, ops._s()), M * (32 + 1024))
for C in (1024, 2048):
    z = torch.randn(M, C, device=dev).to(torch.bfloat16)
    d = torch.randn(M, C, device=dev).to(torch.bfloat16)
    out = torch.empty_like(z)
    mean = torch.zeros(M, device=dev); rstd = torch.ones(M, device=dev)
    gg = torch.ones(C, device=dev); bb = torch.zeros(C, device=dev)
    a1 = torch.zeros(C, device=dev); a2 = torch.zeros(C, device=dev); a3 = torch.zeros(C, device=dev)
    timeit(f"ln fwd C={C}", lambda: call("wf_ln_relu_bf16_fwd", ops._p(z), ops._p(mean), ops._p(rstd), ops._p(gg), ops._p(bb), ops._p(out), M, C, ops._s()),
           M * C * 4)
    timeit(f"ln bwd C={C}", lambda: call("wf_ln_relu_bf16_bwd", ops._p(d), ops._p(z), ops._p(mean), ops._p(rstd), ops._p(gg), ops._p(bb), ops._p(out),
                                         ops._p(a1), ops._p(a2), ops._p(a3), M, C, ops._s()), M * C * 6)
    if C == 1024:
        mask = torch.ones(M, device=dev, dtype=torch.uint8)
        part = torch.empty(call("wf_seg_part_floats", M, C), device=dev)
        timeit("ln fwd colsum C=1024", lambda: call("wf_ln_relu_bf16_fwd_colsum", ops._p(z), ops._p(mean), ops._p(rstd), ops._p(gg), ops._p(bb), ops._p(out),
                                                    ops._p(mask), M, C, 10000, 0, ops._p(part), ops._s()), M * C * 4)
    del z, d, out
