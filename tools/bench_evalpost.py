"""Evaluation post-processing at the model's maximum shape: per sample 64 predicted corners, every one
of the 2,016 corner pairs predicted as an edge (the untrained-model case of evaluate.py), ~90 label edges.

    python tools/bench_evalpost.py [--batch 64] [--cpu-samples 4]

Prints one JSON line: wall time of `APCalculator.compute_metrics` over the whole batch (host packing, H2D, kernels,
D2H and the host set logic included), device time of the three kernels, and the same work through scipy on the host
cores (the reference's formulation: cdist over the 20-point samples + linear_sum_assignment), on a bounded sample."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "wireframe-3d-prediction_b200"), ROOT]
import numpy as np  # noqa: E402
import torch  # noqa: E402
from scipy.optimize import linear_sum_assignment  # noqa: E402
from scipy.spatial.distance import cdist  # noqa: E402

from wf_b200 import evalpost  # noqa: E402
from eval.ap_calculator import APCalculator  # noqa: E402


def make_sample(rng, corners=64, label_edges=90):
    gv = rng.uniform(-1, 1, (corners, 3)).astype(np.float32)
    ge = np.unique(np.sort(rng.integers(0, corners, (label_edges, 2)), axis=1), axis=0)
    ge = ge[ge[:, 0] != ge[:, 1]].astype(np.int64)
    pv = (gv + rng.normal(0, 0.02, gv.shape)).astype(np.float32)
    pe = np.array([[i, j] for i in range(corners) for j in range(i + 1, corners)], dtype=np.int64)
    return pv, pe, evalpost.segment_endpoints(pv, pe), gv, ge, evalpost.segment_endpoints(gv, ge)


def scipy_first_stage(pev, gev):
    """hausdorff_distance_line + linear_sum_assignment exactly as eval/ap_calculator.py:8-36,161 composes them."""
    n, m, s = len(pev), len(gev), 20
    lines = np.concatenate((pev, gev), axis=0)
    w = np.linspace(0, 1, s).reshape(1, s, 1)
    pts = lines[:, 0, :][:, None, :] + w * (lines[:, 1, :][:, None, :] - lines[:, 0, :][:, None, :])
    d = cdist(pts[:n].reshape(-1, 3), pts[n:].reshape(-1, 3)).reshape(n, s, m, s).transpose(0, 2, 1, 3)
    h = np.maximum(d.min(-1).max(-1), d.min(-2).max(-1))
    return h, linear_sum_assignment(h)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--cpu-samples", type=int, default=4)
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    rng = np.random.default_rng(0)
    samples = [make_sample(rng) for _ in range(a.batch)]
    keys = ("predicted_vertices", "predicted_edges", "pred_edges_vertices", "wf_vertices", "wf_edges", "wf_edges_vertices")

    def batch():
        return {k: [s[i].copy() for s in samples] for i, k in enumerate(keys)}

    calc = APCalculator(distance_thresh=1.0)
    calc.compute_metrics(batch())                                   # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(a.iters):
        calc.reset()
        calc.compute_metrics(batch())
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) / a.iters * 1e3

    # the two first-stage calls, device clock (events on the launching stream; includes the uploads of the packed
    # segments / offsets and, for the solver, the read-back of assignments -- not the 93 MB of matrices, which stay put)
    pev = [s[2] for s in samples]; gev = [s[5] for s in samples]
    dev = evalpost._device()
    h = evalpost.hausdorff_lines_batched(pev, gev)
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    torch.cuda.synchronize()
    e0.record()
    d_h, o_off, shapes = evalpost._hausdorff_device(pev, gev, 20, dev)
    e1.record()
    got = evalpost._solve_device(d_h, o_off, shapes, dev)
    e2.record(); torch.cuda.synchronize()
    haus_ms, lsap_ms = e0.elapsed_time(e1), e1.elapsed_time(e2)

    # host reference formulation on a bounded sample, and index-identity of the assignments
    t1 = time.perf_counter()
    ref = [scipy_first_stage(pev[b], gev[b]) for b in range(a.cpu_samples)]
    cpu_ms = (time.perf_counter() - t1) * 1e3 / a.cpu_samples
    same = all(np.array_equal(ref[b][0], h[b]) and np.array_equal(ref[b][1][0], got[b][0])
               and np.array_equal(ref[b][1][1], got[b][1]) for b in range(a.cpu_samples))
    print(json.dumps({
        "workload": f"APCalculator.compute_metrics, batch {a.batch}: 64 corners, 2016 predicted edges x ~90 label edges per sample",
        "gpu_wall_ms_per_batch": wall_ms, "gpu_samples_per_s": a.batch / wall_ms * 1e3,
        "hausdorff_call_ms": haus_ms, "lsap_call_ms": lsap_ms,
        "cpu_scipy_first_stage_ms_per_sample": cpu_ms, "cpu_samples_per_s": 1e3 / cpu_ms,
        "cpu_sample": f"{a.cpu_samples} samples, scipy cdist + linear_sum_assignment (first stage only), {os.cpu_count()} host cores visible, 1 used",
        "first_stage_bit_identical_on_cpu_sample": bool(same),
        "ap_dict": {k: float(calc.ap_dict[k]) for k in ("tp_corners", "tp_edges", "distance", "wed")},
    }), flush=True)


if __name__ == "__main__":
    main()
