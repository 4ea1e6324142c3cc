"""Diagnosis (GPU): where do the element-wise bf16 gradient differences against the reference come from?
Runs a golden training case in bf16 mode with (a) reference matching injected, (b) + reference argmax injected,
(c) + TF32 heads off, and prints per-parameter norm / element errors."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "wireframe-3d-prediction_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import wireframe_oracle as wo
from wf_b200 import ops
from losses.WireframeLoss import WireframeLoss
from test_gpu_model import _model, _ref_col

name = sys.argv[1] if len(sys.argv) > 1 else "train_b2_n384_v12"
g = dict(np.load(os.path.join(ROOT, "tests", "golden", name + ".npz")))
seed, B, N, V, pad, norm_i = [int(v) for v in g["meta"]]
cmin, cmax = [int(v) for v in g["count_range"]]
x, tgt, counts = wo.make_inputs(seed, B, N, V, pad_frac=pad / 1000.0, norm_intensity=bool(norm_i), min_count=cmin, max_count=None if cmax < 0 else cmax)


def run(prec, inject_match, inject_argmax, tf32_heads=True):
    ops.set_precision(prec)
    ops.USE_TF32_HEADS = tf32_heads
    m = _model(seed, V, True)
    xg = x.cuda().requires_grad_(True)
    tg = {k: v.cuda() for k, v in tgt.items()}
    if inject_argmax:
        a = torch.from_numpy(g["pf_argmax"]).to(torch.int32).cuda()
        ops.ARGMAX_OVERRIDE = (a, a)
    pred = m(xg, counts.cuda())
    ops.ARGMAX_OVERRIDE = None
    crit = WireframeLoss(3.0, 1.0, 1.5)
    if inject_match:
        col = _ref_col(g, B, V).cuda()
        crit._match_device = lambda p, t, sync=None: col
    ld = crit(pred, tg)
    ld["total_loss"].backward()
    out = {}
    for k, p in m.named_parameters():
        if "gnone/" + k in g:
            continue
        gr = p.grad.detach().double().reshape(-1).cpu()
        ref_norm = float(g["gnorm/" + k][0])
        samp = torch.from_numpy(g["gsamp/" + k]).double()
        stride = max(1, -(-gr.numel() // 1024))
        scale = max(float(samp.abs().max()), ref_norm / np.sqrt(gr.numel()))
        d = gr[::stride] - samp
        out[k] = (abs(float(gr.norm()) - ref_norm) / ref_norm, float(d.abs().max()) / scale, float(d.norm() / samp.norm().clamp_min(1e-30)))
    return out


cfgs = [("fp32", False, False, True), ("bf16", True, False, True), ("bf16", True, True, True), ("bf16", True, True, False)]
res = [run(*c) for c in cfgs]
print("case", name, "| columns per config: norm-err, max elem err / scale, Frobenius of sample diff")
print("configs:", cfgs)
for k in res[0]:
    print(f"{k:52s} " + " | ".join(f"{r[k][0]:.1e} {r[k][1]:.1e} {r[k][2]:.1e}" for r in res))
