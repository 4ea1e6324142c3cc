"""The BASELINE.json configurations that are NOT the bench.py line (configs[1], [3], [4]): parity-test-sized shapes live in
tests/, this tool measures them at full size on one B200 (device time, CUDA events) and prints one JSON line per case.

    python tools/bench_configs.py [--config 2|4|5|all] [--quick]

  config 2  inference: batch 64 x 10,000 points, 64 vertex slots, eval + no_grad (encoder + heads + edge head; counts from
            the existence head, and a forced-count variant c = 64 that bounds the edge head) + the wireframe matcher
  config 4  encoder-only sweep: 8 clouds x {100k, 250k, 500k, 1M} points (chunked inference encoder) -> Gpts/s and the share
            of the bf16 tensor peak (10,485,760 FLOP/point in the wide layers)
  config 5  matcher / edge-head sweep: B = 4096 samples, V in {16,32,48,64}: loss-style cost + LSAP, matcher-style cost +
            LSAP, edge head forward + backward
Multi-GPU (config 4 at 2/4/8 GPUs): clouds are independent -> batch-sharded, no collective; with fewer clouds than ranks
the points are sharded instead (wf_b200.parallel.encode_point_sharded); run under torchrun with --config 4.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "wireframe-3d-prediction_b200"), ROOT]
os.environ.setdefault("NCCL_DEBUG", "WARN")
import numpy as np  # noqa: E402
import torch  # noqa: E402

FLOP_PER_POINT_FWD = 10485760
CLOUDS = 8


def ev():
    return torch.cuda.Event(enable_timing=True)


def timeit(fn, iters, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1646.8), d.get("bf16_tflops_sustained", 1387.9)
    return 1590.0, 1400.0


def out(line):
    print(json.dumps(line), flush=True)


def config2(quick):
    from wf_b200 import ops
    from wf_b200.synthetic import make_inputs
    from models.PointCloudToWireframe import PointCloudToWireframe
    from models.WireframeHungarianMatcher import build_wireframe_matcher
    B, N, V = 64, 10000, 64
    torch.manual_seed(0)
    model = PointCloudToWireframe(input_dim=8, max_vertices=V).cuda().eval()
    x, tgt, counts = make_inputs(seed=0, B=B, N=N, V=V, min_count=16, max_count=64)
    x = x.cuda()
    forced = torch.full((B,), V, dtype=torch.int64, device="cuda")
    it = 3 if quick else 10
    with torch.no_grad():
        model(x[:2, :256])                                                   # materialise the lazy projection
        enc_ms = timeit(lambda: model.encoder.pooled(x), it)

        # the model takes counts from the existence head in eval mode; random-init weights put every probability near 0.5,
        # so the forced-count variant (train-mode count path, c = 64) is the one that bounds the edge head
        def fwd_pred():
            try:
                return model(x)
            except IndexError:
                return None                                                  # a sample with <= 1 predicted vertices (reference Q6)
        pred_ms = timeit(fwd_pred, it)
        model.train()
        for m in model.modules():
            if isinstance(m, torch.nn.Dropout):
                m.eval()
        model.edge_predictor.attention.dropout = 0.0
        full_ms = timeit(lambda: model(x, forced), it)
        pred = model(x, forced)
        matcher = build_wireframe_matcher(1.0, 1.0)
        tl = [{"vertices": tgt["vertices"][b, :counts[b]].cuda(), "existence": torch.ones(int(counts[b]), device="cuda")} for b in range(B)]
        t0 = time.perf_counter(); matcher(pred, tl); torch.cuda.synchronize(); match_ms = (time.perf_counter() - t0) * 1e3
    burst, sus = peaks()
    tf = B * N * FLOP_PER_POINT_FWD / (enc_ms * 1e-3) / 1e12
    out({"config": "2: inference 64 x 10k points, 64 slots, 1 GPU", "encoder_ms": enc_ms, "encoder_gpts_per_s": B * N / enc_ms / 1e6,
         "encoder_tflops_wide_layers": tf, "encoder_frac_bf16_burst_peak": tf / burst, "encoder_frac_bf16_sustained_peak": tf / sus,
         "forward_ms_predicted_counts": pred_ms, "forward_ms_forced_counts_64": full_ms,
         "samples_per_s_forced_counts_64": B / full_ms * 1e3, "wireframe_matcher_ms_host_timed_incl_sync": match_ms,
         "kernel_launches_counted": ops.LAUNCHES})


def config4(quick):
    import torch.distributed as dist
    from wf_b200.synthetic import make_inputs
    from wf_b200.parallel import encode_point_sharded
    from models.PointNetEncoder import PointNetEncoder
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    torch.manual_seed(0)
    enc = PointNetEncoder().cuda().eval()
    burst, sus = peaks()
    B = CLOUDS
    for N in ((100000,) if quick else (100000, 250000, 500000, 1000000)):
        if world <= B:                                       # batch-sharded: independent clouds, no collective
            bl = B // world
            x, _, _ = make_inputs(seed=rank, B=bl, N=min(N, 100000), V=64)
            if N > 100000:                                   # tile the 100k synthetic cloud (host RNG time, not the bench)
                x = x.repeat(1, -(-N // 100000), 1)[:, :N].contiguous()
            x = x.cuda()
            with torch.no_grad():
                ms = timeit(lambda: enc.pooled(x), max(2, min(20, int(4e6 // (bl * N)))), warm=2)
            mode, pts = f"batch-sharded {bl} clouds/GPU", bl * N * world
        else:                                                # fewer clouds than ranks: shard the points
            n = N // world
            x, _, _ = make_inputs(seed=0, B=B, N=min(N, 100000), V=64)
            x = x.repeat(1, -(-N // 100000), 1)[:, rank * n:(rank + 1) * n].contiguous().cuda()
            ms = timeit(lambda: encode_point_sharded(enc, x, rank, world), max(2, min(20, int(4e6 // (B * n)))), warm=2)
            mode, pts = f"point-sharded {n} points/GPU", B * N
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t)
        if rank == 0:
            tf = pts * FLOP_PER_POINT_FWD / (ms * 1e-3) / 1e12
            out({"config": f"4: encoder-only, {B} clouds x {N} points", "n_gpus": world, "mode": mode, "ms": ms,
                 "gpts_per_s": pts / ms / 1e6, "tflops_wide_layers": tf, "frac_bf16_burst_peak_per_gpu": tf / world / burst,
                 "frac_bf16_sustained_peak_per_gpu": tf / world / sus,
                 "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9})
        del x
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


def config5(quick):
    from scipy.optimize import linear_sum_assignment
    from wf_b200 import ops
    from wf_b200._lib import call
    from models.EdgePredictor import EdgePredictor
    B = 4096
    rng = np.random.default_rng(0)
    torch.manual_seed(0)
    edge = EdgePredictor().cuda().train()
    for m in edge.modules():
        if isinstance(m, torch.nn.Dropout):
            m.eval()
    edge.attention.dropout = 0.0
    for V in ((32,) if quick else (16, 32, 48, 64)):
        pv = torch.from_numpy(rng.uniform(-1, 1, (B, V, 3)).astype(np.float32)).cuda()
        pe = torch.from_numpy(rng.uniform(0, 1, (B, V)).astype(np.float32)).cuda()
        tv = torch.from_numpy(rng.uniform(-1, 1, (B, V, 3)).astype(np.float32)).cuda()
        cnt_h = rng.integers(V // 2, V + 1, (B,)).astype(np.int64)
        cnt = torch.from_numpy(cnt_h).cuda()
        loss_ms = timeit(lambda: ops.loss_match(pv, pe, tv, cnt), 5)
        col, st, cost = ops.loss_match(pv, pe, tv, cnt, want_cost=True)
        c = cost[:256].cpu().numpy()
        t0 = time.perf_counter()
        ref = [linear_sum_assignment(c[b])[1] for b in range(256)]
        cpu_us = (time.perf_counter() - t0) / 256 * 1e6
        same = all(np.array_equal(ref[b], col[b].cpu().numpy()) for b in range(256))
        # matcher-style: cost = cdist_L1 + |e_pred - e_tgt| over V x T_b, rectangular LSAP
        sizes = cnt_h.tolist()
        off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
        tvp = torch.cat([tv[b, :sizes[b]] for b in range(B)]).contiguous()
        tep = torch.ones(int(off[-1]), device="cuda")
        toff = torch.from_numpy(off).cuda()
        costm = torch.zeros(B, V, V, device="cuda")
        nr = torch.full((B,), V, dtype=torch.int32, device="cuda"); nc = torch.from_numpy(cnt_h.astype(np.int32)).cuda()

        def matcher():
            call("wf_wireframe_matcher_cost", ops._p(pv), ops._p(pe), ops._p(tvp), ops._p(tep), ops._p(toff), B, V, 1.0, 1.0,
                 ops._p(costm), V, ops._s())
            return ops.lsap_batched(costm, nr, nc)
        match_ms = timeit(matcher, 5)
        colm, _ = matcher()
        cm = costm[:128].cpu().numpy()
        same_m = True
        for b in range(128):
            r, cc = linear_sum_assignment(cm[b][:, :sizes[b]])
            got = colm[b].cpu().numpy()
            same_m &= bool(np.array_equal(cc, got[r]))
        # edge head forward + backward on the ragged batch (counts = cnt)
        rg = ops.Ragged(sizes, "cuda")
        verts = torch.cat([pv[b, :sizes[b]] for b in range(B)]).contiguous().requires_grad_(True)

        def edge_fb():
            edge.zero_grad(set_to_none=True)
            p = edge.forward_ragged(verts, rg)
            p.sum().backward()
        edge_ms = timeit(edge_fb, 2 if V >= 48 else 3, warm=1)
        with torch.no_grad():
            edge_f_ms = timeit(lambda: edge.forward_ragged(verts, rg), 3, warm=1)
        out({"config": f"5: B=4096, V={V}, counts~U{{{V // 2}..{V}}}", "loss_cost_lsap_ms": loss_ms,
             "loss_cost_lsap_M_matrices_per_s": B / loss_ms / 1e3, "scipy_us_per_matrix_1core": cpu_us,
             "loss_assignments_identical_to_scipy_on_256": bool(same), "matcher_cost_lsap_ms": match_ms,
             "matcher_assignments_identical_to_scipy_on_128": bool(same_m), "edges_total": int(rg.E),
             "edge_head_fwd_ms": edge_f_ms, "edge_head_fwd_bwd_ms": edge_ms,
             "edge_head_fwd_M_edges_per_s": rg.E / edge_f_ms / 1e3})
        del verts, rg, costm
        torch.cuda.empty_cache()


def config_prep(quick):
    """target preparation for a batch (train.py:48-88,112-115) -- device kernel vs the Python loops."""
    from oracle import targets_oracle as to
    from wf_b200.targets import prepare_targets
    for B in ((64,) if quick else (64, 512)):
        verts, edges = to.make_case(5, B, 64, min_c=16)
        prepare_targets(verts, edges, 64, "cuda"); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            got = prepare_targets(verts, edges, 64, "cuda")
        torch.cuda.synchronize()
        gpu_ms = (time.perf_counter() - t0) / 10 * 1e3
        t0 = time.perf_counter(); ref = to.prepare_targets(verts, edges, 64); cpu_ms = (time.perf_counter() - t0) * 1e3
        same = all(torch.equal(got[k].cpu(), ref[k]) for k in ref)
        out({"config": f"prep: targets for a batch of {B} (V=64, counts~U{{16..64}})", "device_path_ms_host_timed": gpu_ms,
             "reference_python_loops_ms": cpu_ms, "identical": bool(same), "batches_per_s_device_path": 1e3 / gpu_ms})


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="all")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--clouds", type=int, default=8, help="config 4: clouds in the batch (fewer than ranks -> point-sharded)")
    a = ap.parse_args()
    CLOUDS = a.clouds
    from wf_b200 import load
    load()
    if a.config in ("2", "all"):
        config2(a.quick)
    if a.config in ("4", "all"):
        config4(a.quick)
    if a.config in ("5", "all"):
        config5(a.quick)
    if a.config in ("prep", "all"):
        config_prep(a.quick)
