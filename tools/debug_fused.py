import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "wireframe-3d-prediction_b200"), ROOT]
import torch
from wf_b200 import ops
from oracle import wireframe_oracle as wo
from models.PointNetEncoder import PointNetEncoder
torch.manual_seed(0)
enc = PointNetEncoder().cuda()
sd = {k[len("encoder."):]: v for k, v in wo.make_state_dict(5, 16).items() if k.startswith("encoder.")}
enc.load_state_dict(sd)
x, _, _ = wo.make_inputs(7, 3, 900, 16, pad_frac=0.15, norm_intensity=True)
x = x.cuda()
res = {}
for fused in (False, True):
    ops.FUSED_POOL = fused
    with torch.no_grad():
        r = enc.pooled(x)
    res[fused] = [t.clone() for t in r[:6]]
names = ["max_m", "avg_m", "max_u", "mean_u", "arg_m", "arg_u"]
for i, n in enumerate(names):
    a, b = res[True][i], res[False][i]
    ne = (a != b)
    print(n, "mismatch", int(ne.sum()), "of", a.numel())
    if ne.any() and i in (0, 2, 4, 5):
        idx = ne.nonzero()[:5]
        for b_, c_ in idx.tolist():
            print("   ", b_, c_, "fused", a[b_, c_].item(), "unfused", b[b_, c_].item(),
                  "max:", res[True][i - 4 if i >= 4 else i][b_, c_].item(), res[False][i - 4 if i >= 4 else i][b_, c_].item())
