"""In-step GPU timeline of the bench.py training step through torch.profiler (CUPTI): per-kernel totals, GPU busy time
and the idle gaps between kernels (where the host could not keep the queue full).  Not a benchmark: profiler overhead
inflates host time.  Usage: python tools/step_timeline.py [steps]"""
import collections
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "wireframe-3d-prediction_b200"), ROOT]
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402
from wf_b200 import ops  # noqa: E402
from wf_b200.parallel import GradAllReduce  # noqa: E402
from models.PointCloudToWireframe import PointCloudToWireframe  # noqa: E402
from losses.WireframeLoss import WireframeLoss  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = PointCloudToWireframe(input_dim=8, max_vertices=bench.VERTS).to(dev).train()
crit = WireframeLoss(vertex_weight=3.0, edge_weight=1.0, existence_weight=1.5)
x_host, tgt_host, _ = bench.make_batch(0, bench.PER_GPU_BATCH)
x = x_host.to(dev); tgt = {k: v.to(dev) for k, v in tgt_host.items()}
with torch.no_grad():
    model(x[:2, :256], tgt["vertex_counts"][:2])
opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-6)
reducer = GradAllReduce(model)


def step():
    reducer.zero()
    pred = model(x, tgt["vertex_counts"])
    ld = crit(pred, tgt)
    ld["total_loss"].backward()
    reducer.finish()
    torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
path = os.path.join(ROOT, "gpurun_out", "step_trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
busy, cur_end, gaps = 0.0, t0, []
for e in ev:
    s, d = e["ts"], e["dur"]
    if s > cur_end:
        gaps.append((s - cur_end, e["name"][:60]))
        busy += d; cur_end = s + d
    else:
        busy += max(0.0, s + d - cur_end); cur_end = max(cur_end, s + d)
agg = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    n = re.sub(r"\(.*", "", e["name"]); n = re.sub(r"^void ", "", n)[:70]
    agg[n][0] += 1; agg[n][1] += e["dur"]
print(f"{steps} steps: span {(t1 - t0) / 1e3 / steps:.2f} ms/step, GPU busy {busy / 1e3 / steps:.2f} ms/step, "
      f"idle {(t1 - t0 - busy) / 1e3 / steps:.2f} ms/step in {len(gaps) // steps} gaps/step")
if len(sys.argv) > 2 and sys.argv[2] == "--seq":
    # kernel sequence of the last step (from its first point_mask kernel)
    starts = [i for i, e in enumerate(ev) if "point_mask" in e["name"]]
    prev_end = None
    for e in ev[starts[-1]:]:
        n = re.sub(r"\(.*", "", e["name"]); n = re.sub(r"^void ", "", n).replace("at::native::", "at::")[:64]
        gap = 0.0 if prev_end is None else e["ts"] - prev_end
        print(f"{e['dur']:9.1f} us  gap {gap:6.1f}  {n}")
        prev_end = e["ts"] + e["dur"]
gaps.sort(reverse=True)
print("largest gaps (us, kernel that followed):", [(round(g, 1), n) for g, n in gaps[:12]])
for n, (c, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{s / 1e3 / steps:8.3f} ms  x{c // steps:<3d} {n}")
os.remove(path)
