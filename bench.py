#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path: training samples/s on 10k-point clouds.

    python bench.py [--gpus N] [--steps K] [--warmup W]             (N>1: launched under torchrun, one rank per GPU)
    python bench.py --impl reference [--gpus N --steps K --warmup W]

A "step" is one pass of the hot path over one batch of synthetic input, exactly the body of the reference's
training loop (train.py:124-142): zero_grad -> PointCloudToWireframe forward -> WireframeLoss (Hungarian matching
inside) -> backward -> [gradient all-reduce when N>1] -> clip_grad_norm_(1.0) -> Adam step.

Workload (config.workload): BASELINE.json configs[2] -- data-parallel training, 64 clouds x 10,000 points x 8 features
per GPU, 64 vertex slots, GT vertex counts ~ U{16..64}, bf16 wide encoder layers (fp32 accumulate), fp32 everything
else; weak scaling (per-GPU batch fixed, global batch 64*N).  Random-init weights of the reference architecture,
synthetic data of the reference's shapes (no network for datasets/checkpoints).

One JSON line is printed by rank 0; see README / DESIGN.md for the keys.  The oracle (oracle/) is executed only by
the `cpu_baseline` leg and by `--impl reference`, as the thing being compared against -- never by the product path.
"""
import argparse
import gc
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "wireframe-3d-prediction_b200")
REF_DIR = os.path.join(ROOT, "oracle", "_ref")          # the unmodified reference (oracle/make_ref.py), git-ignored
REFERENCE_ARM = any(a == "--impl=reference" or (a == "--impl" and sys.argv[i + 1:i + 2] == ["reference"])
                    for i, a in enumerate(sys.argv))
if REFERENCE_ARM:
    # the reference arm never imports the product: `models` / `losses` resolve to the reference's own packages
    sys.path.insert(0, ROOT)
    if os.path.isdir(os.path.join(REF_DIR, "models")):
        sys.path.insert(0, REF_DIR)
else:
    for p in (PKG, ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)

# stdout must carry exactly one JSON line, but libraries write there too (NCCL prints its version banner to stdout at
# NCCL_DEBUG=VERSION and above): keep the real stdout aside and point fd 1 at stderr for the whole run.  NCCL_DEBUG itself
# is left as the caller set it (the driver reads the ranks' NCCL lines from stderr).
sys.stdout.flush()
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


import torch  # noqa: E402

POINTS, VERTS, PER_GPU_BATCH, FEATS = 10000, 64, 64, 8
FLOP_PER_POINT_TRAIN = 29360128          # wide layers: fwd (4 GEMMs) + dX + dW of layers 2-4; SURVEY 8(d)'s 31,457,280 also counts a
                                         # dense layer-5 backward, which the analytic pool backward replaces
FLOP_PER_POINT_FWD = 10485760
METRIC = "train samples/s (10k-pt clouds)"
UNIT = "samples/s"


def workload_name(n_gpus):
    return (f"train step: {PER_GPU_BATCH} clouds/GPU x {POINTS} pts x {FEATS} feats, {VERTS} vertex slots, "
            f"counts~U{{16..64}}, global batch {PER_GPU_BATCH * n_gpus} (BASELINE.json configs[2], weak scaling)")


def _synthetic():
    """wf_b200/synthetic.py loaded by path (pure numpy/torch): the reference arm must not import the product package."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_wf_synthetic", os.path.join(PKG, "wf_b200", "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_batch(rank, B):
    """Synthetic batch shaped like datasets/building3d.py:109-126 output + train.py:48-88 targets."""
    return _synthetic().make_inputs(seed=rank, B=B, N=POINTS, V=VERTS, min_count=16, max_count=64)


class ClockSampler:
    """SM clock, power and throttle reasons sampled DURING the timed region (profiling recipe's clocks line).
    NVML (nvidia_ml_py) is initialised before the warm-up and polled ONCE PER STEP from the main thread, right after the
    step's launches have been enqueued and before the host waits for them: the device is in the middle of the step, and a
    slow NVML call cannot delay a launch or another rank (a polling thread, and before it an external `nvidia-smi -lms`,
    showed up as 10-100 ms outlier steps at 2-8 GPUs: NVML takes driver locks the launch path needs).  Step times come
    from CUDA events, so host time spent here is not in them.  Falls back to `nvidia-smi -lms 100` without NVML."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.rows, self.on, self.nvml, self.handle = index, None, [], False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml
            self.flags = (("hw_slowdown", pynvml.nvmlClocksThrottleReasonHwSlowdown),
                          ("hw_thermal_slowdown", pynvml.nvmlClocksThrottleReasonHwThermalSlowdown),
                          ("sw_thermal_slowdown", pynvml.nvmlClocksThrottleReasonSwThermalSlowdown),
                          ("sw_power_cap", pynvml.nvmlClocksThrottleReasonSwPowerCap))
        except Exception:
            self.nvml = None

    def poll_once(self):
        """One sample; called by the timed loop while the device executes the step it has just been handed."""
        if not self.on or self.nvml is None:
            return
        n = self.nvml
        try:
            sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
            pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
            mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
            self.rows.append((float(sm), float(self.max_sm), pw, [k for k, bit in self.flags if mask & bit]))
        except Exception:
            pass

    def start(self):
        self.rows, self.on = [], True
        if self.nvml is not None:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.proc.stdout:
            f = [t.strip() for t in ln.split(",")]
            try:
                self.rows.append((float(f[0]), float(f[1]), float(f[2]),
                                  [k for k, v in zip(names, f[3:7]) if v.lower().startswith("active")]))
            except (ValueError, IndexError):
                continue

    def stop(self):
        self.on = False
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        rows = list(self.rows)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        reasons = sorted({r for row in rows for r in row[3]})
        return {"sm_mhz": statistics.median(r[0] for r in rows), "sm_max_mhz": max(r[1] for r in rows),
                "power_w_max": max(r[2] for r in rows), "reasons": reasons, "samples": len(rows),
                "source": "nvml, one sample per step while the step executes" if self.nvml is not None else "nvidia-smi -lms 100"}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d.get("bf16_tflops_sustained", d.get("bf16_tflops")), d.get("bf16_tflops"), "measured"
    return 1400.0, 1590.0, "fallback"


def hbm_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)).get("hbm_gbs", 6550.0)
    return 6550.0


def traffic_from_profile():
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)).get("gemm_dram_bytes_per_launch")
        except Exception:
            return None
    return None


# ----------------------------------------------------------------------------------------------
# Reference arm: the UNMODIFIED reference (oracle/_ref, copied by oracle/make_ref.py) through its own public API --
# PointCloudToWireframe.forward, WireframeLoss, the body of train.py:124-142 -- on the host cores (default) or, for the
# `gpu_eager_baseline` field, as PyTorch eager on the B200.  Falls back to the oracle port (CPU only) when oracle/_ref is absent.
# ----------------------------------------------------------------------------------------------
CPU_SAMPLE_B = 8            # clouds per CPU step: a bounded sample of the 64-cloud step (~2 s/step on 16 cores, ~3 GB)


def reference_step_factory(device, sample_b, tf32=False):
    have_ref = os.path.isdir(os.path.join(REF_DIR, "models"))
    dev = torch.device(device)
    if dev.type == "cpu":
        torch.set_num_threads(os.cpu_count() or 1)
    else:
        torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
        torch.backends.cudnn.allow_tf32 = bool(tf32)
    x, tgt, counts = make_batch(0, sample_b)
    if not have_ref:
        if dev.type != "cpu":
            raise SystemExit("oracle/_ref missing: the eager-GPU baseline needs the reference's own modules")
        from oracle import wireframe_oracle as wo
        sd = {k: v.clone().requires_grad_(True) for k, v in wo.make_state_dict(0, VERTS).items()}
        opt = torch.optim.Adam([v for k, v in sd.items() if "spatial_proj" not in k], lr=1e-3, weight_decay=1e-6)

        def step():
            opt.zero_grad(set_to_none=True)
            ld, _ = wo.train_step(sd, x, tgt, max_vertices=VERTS)
            torch.nn.utils.clip_grad_norm_([v for v in sd.values() if v.grad is not None], 1.0)
            opt.step()
            return float(ld["total_loss"].detach())
        return step, "port"
    from models.PointCloudToWireframe import PointCloudToWireframe      # the reference's (sys.path: oracle/_ref first)
    from losses.WireframeLoss import WireframeLoss
    assert os.path.realpath(sys.modules["models.PointCloudToWireframe"].__file__).startswith(os.path.realpath(REF_DIR))
    torch.manual_seed(0)
    model = PointCloudToWireframe(input_dim=FEATS, max_vertices=VERTS).to(dev)          # train.py:40
    model.train()
    crit = WireframeLoss(vertex_weight=3.0, edge_weight=1.0, existence_weight=1.5)      # train.py:90-94
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-6)              # train.py:96 (before the first forward)
    x = x.to(dev)
    tgt = {k: v.to(dev) for k, v in tgt.items()}
    counts = tgt["vertex_counts"]

    def step():                                                                         # train.py:124-145
        opt.zero_grad()
        pred = model(x, counts)
        ld = crit(pred, tgt)
        ld["total_loss"].backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)
        opt.step()
        return ld["total_loss"].item()
    return step, "reference"


def run_reference(device, steps, warmup, sample_b, tf32=False):
    step, kind = reference_step_factory(device, sample_b, tf32)
    cuda = torch.device(device).type == "cuda"
    for _ in range(warmup):
        step()
    if cuda:
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    if cuda:
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return sample_b * steps / dt, dt / steps * 1e3, kind


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return                                   # rank 0 alone runs the reference arm
    if args.ref_device == "cuda":
        # PyTorch eager on the B200 (SURVEY 8d "the real bar"): the full 64-cloud step, fp32 or TF32-allowed
        sample_b = args.ref_batch or PER_GPU_BATCH
        steps, warmup = max(1, min(args.steps, 10)), max(1, min(args.warmup, 3))
        val, ms, kind = run_reference("cuda", steps, warmup, sample_b, tf32=args.ref_tf32)
        emit({"impl": "reference", "device": "cuda", "metric": METRIC, "value": val, "unit": UNIT, "ms_per_step": ms,
              "steps": steps, "warmup": warmup, "batch": sample_b, "kind": kind, "tf32": bool(args.ref_tf32),
              "torch": torch.__version__, "gpu": torch.cuda.get_device_name(0),
              "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30})
        return
    sample_b = args.ref_batch or CPU_SAMPLE_B
    steps, warmup = max(1, min(args.steps, 10)), max(1, min(args.warmup, 2))
    val, ms, kind = run_reference("cpu", steps, warmup, sample_b)
    cores = os.cpu_count() or 1
    what = "the unmodified reference (oracle/_ref) through its own module API" if kind == "reference" else "oracle port"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.gpus),
                   "note": f"CPU arm: {what}; bounded sample of {sample_b} clouds per step"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{sample_b} clouds x {POINTS} pts per step (zero_grad, forward, loss incl. scipy matching, "
                                   f"backward, clip, Adam: train.py:124-142), {steps} steps after {warmup} warm-up, "
                                   f"torch {torch.__version__} CPU, {cores} threads"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def reference_subprocess(extra, timeout=900):
    """Run a reference leg in its own process (its `models`/`losses` packages shadow the product's by name) and return its
    JSON line, or {"unavailable": why}."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference"] + extra
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")}
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
        lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        if r.returncode != 0 or not lines:
            tail = r.stderr.strip().splitlines()[-1] if r.stderr.strip() else "no output"
            return {"unavailable": f"rc={r.returncode}: {tail}"[:300]}
        return json.loads(lines[-1])
    except Exception as e:                                    # noqa: BLE001
        return {"unavailable": f"{type(e).__name__}: {e}"[:300]}


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def main_gpu(args):
    import torch.distributed as dist
    from wf_b200 import ops, load
    from wf_b200.parallel import GradAllReduce, shard_loss_weights
    from models.PointCloudToWireframe import PointCloudToWireframe
    from losses.WireframeLoss import WireframeLoss
    load()                                         # fail loudly if the CUDA library is missing

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ops.set_precision("bf16")

    torch.manual_seed(0)                           # same weights on every rank
    model = PointCloudToWireframe(input_dim=FEATS, max_vertices=VERTS).to(dev)
    model.train()
    crit = WireframeLoss(vertex_weight=3.0, edge_weight=1.0, existence_weight=1.5)      # train.py:90-94
    B = PER_GPU_BATCH
    x_host, tgt_host, counts_host = make_batch(rank, B)
    x_pin = x_host.pin_memory()
    tgt_pin = {k: v.pin_memory() for k, v in tgt_host.items()}
    x = x_pin.to(dev, non_blocking=True)
    tgt = {k: v.to(dev, non_blocking=True) for k, v in tgt_pin.items()}
    counts = tgt["vertex_counts"]
    host_counts = [int(c) for c in counts_host.tolist()]
    all_counts = [make_counts_only(r, B) for r in range(world)]
    wv, wx, we = shard_loss_weights(counts_host.tolist(), all_counts, VERTS)

    # materialise the lazy point_pool_proj before the optimizer exists, so DP replicas stay identical
    with torch.no_grad():
        model(x[:2, :256], counts[:2])
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-6, eps=1e-8, betas=(0.9, 0.999))   # train.py:96
    reducer = GradAllReduce(model)

    def step(xin, targets, want_loss):
        reducer.zero()
        pred = model(xin, targets["vertex_counts"])
        ld = crit(pred, targets)
        loss = (3.0 * wv) * ld["vertex_loss"] + (1.5 * wx) * ld["existence_loss"] + (1.0 * we) * ld["edge_loss"] \
            if world > 1 else ld["total_loss"]
        loss.backward()
        reducer.finish()
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)
        opt.step()
        return loss if want_loss else None          # the caller reads it back (train.py:143 `.item()`)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step_ms = {False: [], True: []}
    copy_stream = torch.cuda.Stream(dev)

    # two preallocated device staging sets (double buffer): the per-step copy never touches the caching allocator
    stage = [(torch.empty_like(x), {k: torch.empty_like(v) for k, v in tgt.items()}) for _ in range(2)]
    stage_free = [None, None]                     # event: the step that last read this set has finished

    def h2d(slot):
        """This step's inputs from pinned host memory into staging set `slot`, on the copy stream (overlaps the previous
        step's compute; waits until the step that last used the set is done)."""
        xs, tg = stage[slot]
        with torch.cuda.stream(copy_stream):
            if stage_free[slot] is not None:
                copy_stream.wait_event(stage_free[slot])
            xs.copy_(x_pin, non_blocking=True)
            for k, v in tgt_pin.items():
                tg[k].copy_(v, non_blocking=True)
            ev = torch.cuda.Event(); ev.record(copy_stream)
            for t in tg.values():
                ops.mark_ready(t, ev)             # as wf_b200.targets.DevicePrefetcher tags what it stages
            # ... and the host's own copy of the counts travels with the device one (no read-back inside the step)
            tg["vertex_counts"]._wf_host_counts = (tg["vertex_counts"]._version, tuple(host_counts))
        return xs, tg, ev

    def timed(n, e2e):
        """n steps, each bracketed by CUDA events, L2 flushed between steps (outside the events).  e2e: every step issues the
        pinned host->device copy of the NEXT step's inputs inside its own timed region (side stream, as wf_b200.targets.
        DevicePrefetcher does; the first step also pays for its own copy) and reads the loss back to the host."""
        total_ms = 0.0
        nxt = None
        for i in range(n):
            flush.fill_(1)                                                  # L2 flush, outside the timed region
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if e2e:
                cur = nxt if nxt is not None else h2d(i % 2)
                xin, tg, ev = cur
                torch.cuda.current_stream().wait_event(ev)
                loss_t = step(xin, tg, True)
                stage_free[i % 2] = torch.cuda.Event(); stage_free[i % 2].record()
                # the next batch's copies are enqueued while the device works on this step (a prefetching loader's order), then
                # the loss scalar is read back (D2H, waits for the whole step)
                nxt = h2d((i + 1) % 2) if i + 1 < n else None
                loss_t.item()
            else:
                step(x, tgt, False)
            e1.record()
            if sampler is not None:
                sampler.poll_once()                                         # the device is still executing this step
            torch.cuda.synchronize()
            total_ms += e0.elapsed_time(e1)
            step_ms[e2e].append(round(e0.elapsed_time(e1), 3))
        return total_ms

    sampler = ClockSampler(local) if (rank == 0 and os.environ.get("WF_BENCH_NO_CLOCKS") != "1") else None   # NVML initialised before the warm-up
    # Warm-up.  With more than one rank the first ~15 steps after start-up carry isolated 10-100 ms stalls on one rank
    # (seen at 2, 4 and 8 GPUs with and without clock sampling or per-launch events, never later: diagnosis in DESIGN.md section 4),
    # so multi-GPU runs settle for 20 further untimed steps; the JSON line reports the warm-up actually done.
    # Every run also settles into the sustained power state first: inside a cold 20-step region the step time drifted by ~10 %
    # (r01: 22.4 -> 24.8 ms) as the package reached its power cap, so a short region flattered the steady state.
    # (WF_BENCH_SETTLE overrides the settling steps: the ncu launch-list pass, whose per-launch times are cold anyway, uses 0)
    n_warm = max(args.warmup, 3) + int(os.environ.get("WF_BENCH_SETTLE", "30" if world > 1 else "20"))
    for _ in range(n_warm):
        step(x, tgt, False)
    barrier()
    gc.collect()
    gc.freeze()                                   # the long-lived module/optimizer objects leave the collector's young
    #                                               generations: no multi-ms collection pause lands on one rank mid-step
    if sampler is not None:
        sampler.start()
    # ---- timed region 1: device-resident inputs, kernel-level GEMM timing on the launching stream
    if os.environ.get("WF_BENCH_DIAG") == "1":          # diagnosis: the same region without the per-GEMM events, twice
        for k in range(2):
            barrier(); timed(args.steps, e2e=False); barrier()
            if rank == 0:
                print(f"diag region {k} (no GEMM events):", step_ms[False], file=sys.stderr, flush=True)
            step_ms[False].clear()
    ops.GEMM_PROFILE = []
    l0 = ops.LAUNCHES
    barrier()
    ms = timed(args.steps, e2e=False)
    barrier()
    launches = ops.LAUNCHES - l0
    prof, ops.GEMM_PROFILE = ops.GEMM_PROFILE, None
    clocks = sampler.stop() if sampler is not None else None
    # ---- timed region 2: end to end through the public API with host buffers (two untimed steps first: the pinned
    #      staging path is exercised once per buffer set first)
    barrier()
    timed(2, e2e=True)
    step_ms[True].clear()
    barrier()
    ms_e2e = timed(args.steps, e2e=True)
    barrier()
    # ---- supplementary region (not part of `value`): the same steps with the LayerNorm side jobs switched off, so that the GEMM
    #      launches do nothing but the GEMM -- the kernel's own tensor roofline, measured live in this process
    pure = None
    if ops.SIDE_JOBS and os.environ.get("WF_BENCH_NO_PURE") != "1":
        ops.SIDE_JOBS = False
        try:
            timed(2, e2e=False)
            step_ms[False] = step_ms[False][:-2]
            ops.GEMM_PROFILE = []
            barrier()
            n_pure = max(2, min(args.steps, 5))
            ms_pure = timed(n_pure, e2e=False)
            step_ms[False] = step_ms[False][:-n_pure]
            barrier()
            prof_pure, ops.GEMM_PROFILE = ops.GEMM_PROFILE, None
            g_ms = sum(r[0].elapsed_time(r[1]) for r in prof_pure)
            g_fl = sum(r[2] for r in prof_pure)
            pure = {"steps": n_pure, "ms_per_step": ms_pure / n_pure, "gemm_ms_per_step": g_ms / n_pure,
                    "launches_per_step": len(prof_pure) // n_pure,
                    "achieved": g_fl / (g_ms * 1e-3) / 1e12 if g_ms > 0 else None}
        finally:
            ops.SIDE_JOBS = True
            ops.GEMM_PROFILE = None
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
    gemm_ms = sum(r[0].elapsed_time(r[1]) for r in prof)
    gemm_flop = sum(r[2] for r in prof)
    side_bytes = sum(r[3] for r in prof)          # LayerNorm passes executed by spare warps inside these launches
    if rank == 0:
        sustained, burst, src = peaks()
        samples = B * world * args.steps
        h2d = x_pin.numel() * 4 + sum(v.numel() * v.element_size() for v in tgt_pin.values())
        achieved = gemm_flop / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
        hbm_peak = hbm_peak_gbs()
        # the same launches also stream the LayerNorm passes of other row chunks (side jobs): time the two would take one after
        # the other, each at its own measured peak, over the time the launches took
        serial_ms = gemm_flop / (sustained * 1e12) * 1e3 + side_bytes / (hbm_peak * 1e9) * 1e3
        line = {
            "metric": METRIC, "value": samples / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": n_warm, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(world), "parallelism": f"dp{world}",
                       "l2": "256 MB buffer written between timed steps (outside the events); per-step working set ~11 GB",
                       "precision": "bf16 operands / fp32 accumulate on encoder layers 2-5, fp32 elsewhere",
                       "dropout": "enabled (edge head, p=0.1)"},
            "clocks": clocks,
            "e2e": {"value": samples / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "roofline": {"kernel": f"wf::tc::gemm_tc_kernel<bf16, 2-SM MMA> ({len(prof) // max(1, args.steps)} launches/step: 4 fwd incl. the pooling "
                                   "epilogue + 3 dX + 3 dW, each over the row chunks of the encoder pipeline; the layer-5 backward is "
                                   "analytic).  The launches also carry the LayerNorm passes of other chunks as side jobs "
                                   "(side_jobs below), so `achieved` = tensor FLOP / launch time understates the kernel", "bound": "tensor",
                         "achieved": achieved, "peak": sustained, "peak_burst": burst, "peak_source": src,
                         "unit": "TFLOP/s", "frac": achieved / sustained if sustained else None,
                         "traffic": traffic_from_profile(),
                         "gemm_ms_per_step": gemm_ms / args.steps, "gemm_share_of_step": gemm_ms / ms if ms else None,
                         "flop_per_launch_avg": gemm_flop / max(1, len(prof)),
                         "without_side_jobs": None if pure is None else dict(
                             pure, frac=(pure["achieved"] / sustained if pure["achieved"] and sustained else None),
                             what="supplementary steps after the timed regions with WF_B200_SIDE=0 semantics: the launches run "
                                  "the GEMM only, the LayerNorm passes are stand-alone kernels"),
                         "side_jobs": {"hbm_bytes_per_step": side_bytes / max(1, args.steps), "hbm_peak_gbs": hbm_peak,
                                       "gemm_at_peak_then_passes_at_peak_ms_per_step": serial_ms / max(1, args.steps),
                                       "frac_of_serial_peaks": serial_ms / gemm_ms if gemm_ms > 0 else None},
                         "encoder_train_mpts_per_s": (B * POINTS * args.steps) / (ms * 1e-3) / 1e6},
        }
        sm = step_ms[False]
        if len(sm) >= 4:
            h = min(10, len(sm) // 2)
            line["step_ms_first_last"] = {"first": round(sum(sm[:h]) / h, 3), "last": round(sum(sm[-h:]) / h, 3), "n": h}
        if world == 1 and not args.no_cpu:
            # baselines, each in its own process after the timed regions (their `models` package shadows ours by name)
            del flush
            torch.cuda.empty_cache()
            r = reference_subprocess(["--steps", "3", "--warmup", "1"])
            line["cpu_baseline"] = r.get("cpu_baseline", r)
            if not args.no_eager:
                eager = {}
                for name, extra in (("fp32", []), ("tf32", ["--ref-tf32"])):
                    e = reference_subprocess(["--ref-device", "cuda", "--steps", "5", "--warmup", "2"] + extra)
                    eager[name] = {k: e[k] for k in ("value", "ms_per_step", "batch", "kind", "peak_mem_gb", "unavailable") if k in e}
                line["gpu_eager_baseline"] = {
                    "what": "the unmodified reference modules as PyTorch eager on this B200 (cuBLAS/ATen kernels, per-sample "
                            "Python loops, scipy matching on the host): the same 64 x 10,000-point train step (train.py:124-142)",
                    "unit": UNIT, **eager}
        emit(line)
        print("per-step ms (device-resident):", step_ms[False], "\nper-step ms (e2e):", step_ms[True], file=sys.stderr)
    if world > 1:
        dist.destroy_process_group()


def make_counts_only(rank, B):
    from wf_b200.synthetic import make_counts
    return make_counts(seed=rank, B=B, V=VERTS, min_count=16, max_count=64)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="wf_b200", choices=["wf_b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline and gpu_eager_baseline legs")
    ap.add_argument("--no-eager", action="store_true", help="skip the gpu_eager_baseline leg")
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"], help="reference arm: host cores (default) or eager on the GPU")
    ap.add_argument("--ref-batch", type=int, default=0, help="reference arm: clouds per step (default 8 on CPU, 64 on the GPU)")
    ap.add_argument("--ref-tf32", action="store_true", help="reference arm on the GPU: allow TF32 matmuls")
    a = ap.parse_args()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_gpu(a)
