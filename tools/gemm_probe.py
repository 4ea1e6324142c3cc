"""Times wf_gemm_bf16 on the encoder's layer shapes next to torch.matmul (cuBLAS) on the same operands.
Usage: python tools/gemm_probe.py [points]   -- prints one line per shape; device-timed with CUDA events."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "wireframe-3d-prediction_b200"))
from wf_b200 import ops  # noqa: E402


def timeit(fn, iters=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 640000
    dev = "cuda"
    print(f"points M={M}")
    for (K, N) in ((512, 1024), (1024, 2048), (2048, 1024), (1024, 512)):
        A = torch.randn(M, K, device=dev).to(torch.bfloat16)
        W = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
        bias = torch.randn(N, device=dev)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        from wf_b200._lib import call
        stats = torch.empty(call("wf_gemm_rowstats_parts", N), M, 2, device=dev)
        t = timeit(lambda: ops.gemm_bf16(A, W, M=M, N=N, K=K, bias=bias, out=out, rowstats=stats))
        tt = timeit(lambda: torch.matmul(A, W.t()))
        fl = 2.0 * M * N * K
        err = (out.float() - (A.float() @ W.float().t() + bias)).abs().max().item() if M <= 200000 else float("nan")
        del stats
        print(f"fwd  K={K:5d} N={N:5d}: wf {t:8.3f} ms {fl / t / 1e9:8.1f} TF/s | cublas {tt:8.3f} ms {fl / tt / 1e9:8.1f} TF/s | maxerr {err:.3e}")
        # dX shape: dZ[M,N] * W[N,K] -> [M,K]  (B operand = W^T stored [K,N])
        dZ = torch.randn(M, N, device=dev).to(torch.bfloat16)
        Wt = W.t().contiguous()
        dX = torch.empty(M, K, device=dev, dtype=torch.bfloat16)
        t = timeit(lambda: ops.gemm_bf16(dZ, Wt, M=M, N=K, K=N, out=dX))
        tt = timeit(lambda: torch.matmul(dZ, W))
        print(f"dX   K={N:5d} N={K:5d}: wf {t:8.3f} ms {fl / t / 1e9:8.1f} TF/s | cublas {tt:8.3f} ms {fl / tt / 1e9:8.1f} TF/s")
        # dW shape: dZ^T[N,M] * A[M,K] -> [N,K], reduction over points
        dW = torch.zeros(N, K, device=dev)
        tiles = ((N + 127) // 128) * ((K + 255) // 256)
        split = max(1, (2 * 148 + tiles - 1) // tiles)
        t = timeit(lambda: ops.gemm_bf16(dZ, A, M=N, N=K, K=M, kmajor=False, out=dW, accumulate=True, split_k=split))
        tt = timeit(lambda: torch.matmul(dZ.t(), A))
        print(f"dW   M={N:5d} N={K:5d} split={split}: wf {t:8.3f} ms {fl / t / 1e9:8.1f} TF/s | cublas {tt:8.3f} ms {fl / tt / 1e9:8.1f} TF/s")
    # the pooling launch: final Linear 1024 -> 512 whose epilogue max-pools per cloud instead of storing
    K, N, clouds = 1024, 512, 64
    A = torch.randn(M, K, device=dev).to(torch.bfloat16)
    W = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device=dev)
    mask = torch.ones(M, device=dev, dtype=torch.uint8)
    packed = torch.zeros(2, clouds, N, device=dev, dtype=torch.int64)
    t = timeit(lambda: ops.gemm_bf16_pool(A, W, M=M, N=N, K=K, bias=bias, points_per_cloud=M // clouds, row_offset=0, mask=mask, packed=packed))
    mask[::7] = 0
    t2 = timeit(lambda: ops.gemm_bf16_pool(A, W, M=M, N=N, K=K, bias=bias, points_per_cloud=M // clouds, row_offset=0, mask=mask, packed=packed))
    fl = 2.0 * M * N * K
    print(f"pool K={K:5d} N={N:5d}: wf {t:8.3f} ms {fl / t / 1e9:8.1f} TF/s (all points valid) | {t2:8.3f} ms {fl / t2 / 1e9:8.1f} TF/s (1/7 masked out)")


if __name__ == "__main__":
    main()
