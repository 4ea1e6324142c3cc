"""Runs the reference's UNMODIFIED callers -- train.py:train_model and evaluate.py:evaluate_with_ap_calculator (copied to
oracle/_ref by oracle/make_ref.py) -- in one process, either against the drop-in package (`--impl dropin`: its models/,
losses/ and eval/ packages shadow the reference's by PYTHONPATH order, exactly as INTEGRATION.md tells a user to deploy
it) or against the reference's own modules (`--impl reference`, CPU).  Used by tests/test_gpu_real_callers.py.

Nothing in train.py / evaluate.py / datasets/ is edited.  What this driver does around them, identically for both arms:
  * supplies `easydict` (not installed in this image; a 3-line attribute dict) and a no-op `wandb` if it is missing;
  * seeds torch / numpy / random, so that both arms draw the same weights and the same augmented, resampled batch;
  * switches the edge head's dropout off by constructing nn.Dropout / nn.MultiheadAttention with p = 0 (SURVEY Q4:
    the two arms cannot share dropout masks), a construction-time default, not a code change;
  * records every total_loss the criterion returns (train.py keeps its loss_history to itself) and the final ap_dict.
Prints one JSON line."""
import argparse
import json
import os
import random
import sys
import types

ap = argparse.ArgumentParser()
ap.add_argument("--impl", required=True, choices=["dropin", "reference"])
ap.add_argument("--workdir", required=True)
ap.add_argument("--epochs", type=int, default=20)
ap.add_argument("--precision", default="fp32")
ap.add_argument("--skip-eval", action="store_true")
args = ap.parse_args()

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
PKG = os.path.join(ROOT, "wireframe-3d-prediction_b200")
if args.impl == "reference":
    os.environ["CUDA_VISIBLE_DEVICES"] = ""                 # train.py:21 / evaluate.py:41 pick cuda when available
    sys.path[:0] = [REF]
else:
    os.environ["WF_B200_PRECISION"] = args.precision
    sys.path[:0] = [PKG, REF]                               # models/ losses/ eval/ from the drop-in; train, evaluate, datasets from the reference

# ---- stand-ins for packages this image lacks
if "easydict" not in sys.modules:
    try:
        import easydict  # noqa: F401
    except ImportError:
        m = types.ModuleType("easydict")

        class EasyDict(dict):
            def __init__(self, d=None, **kw):
                super().__init__()
                for k, v in {**(d or {}), **kw}.items():
                    self[k] = EasyDict(v) if isinstance(v, dict) else v
            __getattr__ = dict.__getitem__
            __setattr__ = dict.__setitem__
        m.EasyDict = EasyDict
        sys.modules["easydict"] = m
try:
    import wandb  # noqa: F401
except Exception:                                           # noqa: BLE001
    sys.modules["wandb"] = types.ModuleType("wandb")

import numpy as np  # noqa: E402
import torch  # noqa: E402

# ---- dropout off (construction-time defaults)
_drop_init = torch.nn.Dropout.__init__
torch.nn.Dropout.__init__ = lambda self, p=0.5, inplace=False: _drop_init(self, 0.0, inplace)
_mha_init = torch.nn.MultiheadAttention.__init__


def _mha(self, embed_dim, num_heads, dropout=0.0, *a, **k):
    _mha_init(self, embed_dim, num_heads, 0.0, *a, **k)


torch.nn.MultiheadAttention.__init__ = _mha

os.makedirs(args.workdir, exist_ok=True)
os.chdir(args.workdir)
if not os.path.exists("datasets"):
    os.symlink(os.path.join(REF, "datasets"), "datasets")

torch.manual_seed(0); np.random.seed(0); random.seed(0)

import train as ref_train  # noqa: E402      (oracle/_ref/train.py, unmodified)
from torch.utils.data import DataLoader  # noqa: E402
from datasets import build_dataset  # noqa: E402
from easydict import EasyDict  # noqa: E402
import yaml  # noqa: E402
import losses.WireframeLoss as wl  # noqa: E402
import models.PointCloudToWireframe as pcw  # noqa: E402

which = os.path.realpath(pcw.__file__)
assert which.startswith(os.path.realpath(PKG if args.impl == "dropin" else REF)), which
assert os.path.realpath(ref_train.__file__).startswith(os.path.realpath(REF))

loss_log = []
_fwd = wl.WireframeLoss.forward


def _recording_forward(self, predictions, targets):
    out = _fwd(self, predictions, targets)
    loss_log.append(float(out["total_loss"].detach()))
    return out


wl.WireframeLoss.forward = _recording_forward

cfg = EasyDict(yaml.safe_load(open("datasets/dataset_config.yaml")))
ds = build_dataset(cfg.Building3D)                          # main.py:35
loader = DataLoader(ds["train"], batch_size=3, shuffle=False, drop_last=True, collate_fn=ds["train"].collate_batch)   # main.py:42-48, unshuffled
model = ref_train.train_model(loader, num_epochs=args.epochs, learning_rate=0.001, wandb_run=None)     # main.py:50
torch.save(model.state_dict(), "trained_model.pth")         # main.py:53
res = {"impl": args.impl, "module_file": which, "device": str(next(model.parameters()).device), "losses": loss_log,
       "state_dict_keys": len(model.state_dict()), "max_vertices": model.max_vertices}

if not args.skip_eval:
    import evaluate as ref_eval  # noqa: E402 (oracle/_ref/evaluate.py, unmodified)
    import eval.ap_calculator as apc  # noqa: E402
    assert os.path.realpath(apc.__file__).startswith(os.path.realpath(PKG if args.impl == "dropin" else REF))
    captured = {}
    _out = apc.APCalculator.output_accuracy

    def _capturing(self):
        _out(self)
        captured.update({k: float(v) for k, v in self.ap_dict.items()})
        captured["samples"] = int(self.batch_size)

    apc.APCalculator.output_accuracy = _capturing
    torch.manual_seed(1); np.random.seed(1); random.seed(1)  # evaluate.py draws a fresh point_pool_proj and resamples points
    ref_eval.evaluate_with_ap_calculator()
    res["ap_dict"] = captured
print("RESULT " + json.dumps(res))
