"""BASELINE.json configs[4]: loss-style cost + LSAP for B=4096 samples at V=16..64 on the GPU (device time, CUDA events)
next to scipy on the host for a 256-matrix sample.  python tools/lsap_probe.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "wireframe-3d-prediction_b200"))
from wf_b200 import ops  # noqa: E402
from scipy.optimize import linear_sum_assignment  # noqa: E402

B = 4096
rng = np.random.default_rng(0)
for V in (16, 32, 48, 64):
    pv = torch.from_numpy(rng.uniform(-1, 1, (B, V, 3)).astype(np.float32)).cuda()
    pe = torch.from_numpy(rng.uniform(0, 1, (B, V)).astype(np.float32)).cuda()
    tv = torch.from_numpy(rng.uniform(-1, 1, (B, V, 3)).astype(np.float32)).cuda()
    cnt = torch.from_numpy(rng.integers(V // 2, V + 1, (B,)).astype(np.int64)).cuda()
    ops.loss_match(pv, pe, tv, cnt); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        col, st, _ = ops.loss_match(pv, pe, tv, cnt)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    col, st, cost = ops.loss_match(pv, pe, tv, cnt, want_cost=True)
    c = cost[:256].cpu().numpy()
    t0 = time.perf_counter()
    ref = [linear_sum_assignment(c[b])[1] for b in range(256)]
    cpu = (time.perf_counter() - t0) / 256
    same = all(np.array_equal(ref[b], col[b].cpu().numpy()) for b in range(256))
    print(f"V={V:3d}: GPU {ms:7.3f} ms for {B} matrices = {B / ms * 1e3 / 1e6:6.2f} M matrices/s | scipy {cpu * 1e6:7.1f} us/matrix "
          f"= {1 / cpu / 1e6:6.3f} M/s (1 core) | identical on sample: {same}")
