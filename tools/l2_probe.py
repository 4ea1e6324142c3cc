"""GPU probe: the inference encoder with L2-resident activations (ops.encoder_pooled_infer(l2_resident=True): 9472-row chunks
through two ping-pong buffers, LayerNorm in place) against the default form (chunks of 2 x 320k rows, activations through HBM).
Checks that both give the same bits, times them eagerly and as a replayed CUDA graph, sustained (back to back).

    python tools/l2_probe.py                      timing lines (JSON)
    PROBE_MODE=default|l2|l2graph python ...      one warm call, then ONE call inside cudaProfilerStart/Stop (for ncu
                                                  --profile-from-start off: DRAM bytes of every launch of that call)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "wireframe-3d-prediction_b200"), ROOT]
import torch  # noqa: E402

from wf_b200 import ops  # noqa: E402
from wf_b200.synthetic import make_inputs  # noqa: E402
from models.PointNetEncoder import PointNetEncoder  # noqa: E402

B = int(os.environ.get("PROBE_B", 64)); N = int(os.environ.get("PROBE_N", 10000))
reps = int(os.environ.get("PROBE_REPS", 30))
rows = int(os.environ.get("PROBE_ROWS", ops.INFER_L2_ROWS))
mode = os.environ.get("PROBE_MODE", "")
torch.manual_seed(0)
enc = PointNetEncoder().cuda().eval()
x, _, _ = make_inputs(seed=0, B=min(B, 64), N=min(N, 100000), V=64)
if N > 100000:
    x = x.repeat(1, N // 100000, 1)
x = x.cuda()
x, p = enc.tc_inputs(x)
p = [t.detach() for t in p]


def default():
    return ops.encoder_pooled_infer(x, p)


def l2():
    return ops.encoder_pooled_infer(x, p, l2_resident=True, chunk_rows=rows)


def same(a, b):
    return all(torch.equal(u, v) for u, v in zip(a, b))


with torch.no_grad():
    ref = default()
    got = l2()
    torch.cuda.synchronize()
    assert same(ref, got), "L2-resident encoder differs from the default one"
    # CUDA graph of the whole call (every launch goes to torch's current stream; tensor maps are kernel parameters)
    fns = {"default": default, "l2": l2}
    graph_err = None
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            l2()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            gout = l2()
        graph.replay()
        torch.cuda.synchronize()
        assert same(ref, gout), "graph replay differs"
        fns["l2graph"] = graph.replay
    except Exception as e:  # noqa: BLE001
        graph_err = repr(e)[:300]
    if mode:
        fns[mode]()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        fns[mode]()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        sys.exit(0)
    res = {"shape": f"{B}x{N}", "l2_rows": rows, "bit_identical": True, "graph_error": graph_err}
    flops = B * N * 10485760
    for name, fn in fns.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        n0 = ops.LAUNCHES
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res[name] = {"ms": round(ms, 4), "mpts_per_s": round(B * N / ms / 1e3, 1), "tflops_wide_layers": round(flops / ms / 1e9, 1),
                     "launches": (ops.LAUNCHES - n0) // reps}
    print(json.dumps(res), flush=True)
