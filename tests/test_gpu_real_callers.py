"""The reference's UNMODIFIED callers against the drop-in: `train.py:train_model` (20 epochs on the shipped
buildings, as main.py drives it) followed by `evaluate.py:evaluate_with_ap_calculator` on the checkpoint it saved, run once
with the drop-in's models/ losses/ eval/ packages shadowing the reference's (PYTHONPATH order) and
once with the reference's own modules on the CPU.  tests/real_callers_driver.py is the harness; oracle/_ref (git-ignored
copy made by oracle/make_ref.py) holds train.py, evaluate.py and datasets/ as they are upstream."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
EPOCHS = 20


def _run(impl, workdir, precision="fp32"):
    cmd = [sys.executable, os.path.join(ROOT, "tests", "real_callers_driver.py"), "--impl", impl, "--workdir", str(workdir),
           "--epochs", str(EPOCHS), "--precision", precision]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")]
    assert r.returncode == 0 and lines, f"{impl} arm failed:\n{r.stdout[-2000:]}\n{r.stderr[-3000:]}"
    return json.loads(lines[-1][len("RESULT "):])


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "datasets", "train")), reason="oracle/_ref not present (python oracle/make_ref.py)")
def test_unmodified_train_and_evaluate_scripts_run_on_the_dropin(tmp_path):
    from gpu_util import record
    ref = _run("reference", tmp_path / "ref")
    ours = _run("dropin", tmp_path / "fp32", "fp32")
    bf16 = _run("dropin", tmp_path / "bf16", "bf16")
    assert "wireframe-3d-prediction_b200" in ours["module_file"] and "oracle/_ref" in ref["module_file"]
    assert ours["device"].startswith("cuda") and ref["device"] == "cpu"
    assert ours["state_dict_keys"] == ref["state_dict_keys"] == 80 and ours["max_vertices"] == ref["max_vertices"]
    lr, lo, lb = np.array(ref["losses"]), np.array(ours["losses"]), np.array(bf16["losses"])
    assert len(lr) == len(lo) == len(lb) == EPOCHS
    rel = np.abs(lo - lr) / np.abs(lr)
    relb = np.abs(lb - lr) / np.abs(lr)
    obs = {"loss_rel_epoch0": float(rel[0]), "loss_rel_epoch1": float(rel[1]), "loss_rel_max": float(rel.max()),
           "bf16_loss_rel_epoch0": float(relb[0]), "bf16_loss_rel_max": float(relb.max()),
           "final_loss_ref": float(lr[-1]), "final_loss_fp32": float(lo[-1]), "final_loss_bf16": float(lb[-1]),
           "ap_ref": ref["ap_dict"], "ap_fp32": ours["ap_dict"], "ap_bf16": bf16["ap_dict"]}
    record("real_callers", **{k: v for k, v in obs.items() if not k.startswith("ap_")},
           ap_ref=json.dumps(ref["ap_dict"]), ap_fp32=json.dumps(ours["ap_dict"]), ap_bf16=json.dumps(bf16["ap_dict"]))
    print(json.dumps(obs, indent=1))
    # same weights (same torch seed, same registration order), same batch: the first loss is the reference's to rounding (the
    # shipped buildings carry un-normalised intensity ~5e4, SURVEY D6, which costs fp32 two digits in the first LayerNorm --
    # in the reference as well; observed 7e-5 in fp32 mode, 1.5e-4 in bf16 mode)
    assert rel[0] < 5e-4 and relb[0] < 2e-3
    # afterwards Adam (train.py:96) turns every gradient entry into a +-lr step, including the entries whose true gradient is
    # zero (all Linear biases in front of a LayerNorm): two correct implementations part ways at the 1e-3 level after one
    # step and chaotically later (tools/dp_proof.py shows the same between 2-GPU and 1-GPU runs of THIS code); the curves
    # must stay in the same band and both must have learnt
    # (observed: 6e-3 after one step, at most 6 % / 10 % anywhere on the 20-epoch curve, final losses 1.0732 / 1.0759 / 1.0726)
    assert rel[1] < 3e-2
    assert lo[-1] < lo[0] and lb[-1] < lb[0] and lr[-1] < lr[0]
    assert rel.max() < 0.3 and relb.max() < 0.3
    assert abs(lo[-1] - lr[-1]) < 0.05 * lr[-1] and abs(lb[-1] - lr[-1]) < 0.05 * lr[-1]
    # evaluate.py ran on each arm's own checkpoint; the metric dictionary has the reference's keys
    assert set(ours["ap_dict"]) == set(ref["ap_dict"]) == set(bf16["ap_dict"])
    assert ours["ap_dict"]["samples"] == ref["ap_dict"]["samples"]
    assert ours["ap_dict"]["tp_fn_corners"] == ref["ap_dict"]["tp_fn_corners"]          # ground-truth corner count: data only
    assert ours["ap_dict"]["tp_fn_edges"] == ref["ap_dict"]["tp_fn_edges"]


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "datasets", "train")), reason="oracle/_ref not present (python oracle/make_ref.py)")
def test_evaluate_script_on_one_checkpoint_gives_the_reference_metrics(tmp_path):
    """evaluate.py on the SAME checkpoint in both arms (the reference arm's trained_model.pth): with identical weights the
    drop-in's eval-mode forward and its device APCalculator must reproduce the reference's counters."""
    import shutil
    ref = _run("reference", tmp_path / "ref")
    wd = tmp_path / "same"
    os.makedirs(wd, exist_ok=True)
    shutil.copy(tmp_path / "ref" / "trained_model.pth", wd / "trained_model.pth")
    code = (
        "import sys, os, json, types, random; sys.argv=['x']\n"
        f"ROOT={ROOT!r}; sys.path[:0]=[os.path.join(ROOT,'wireframe-3d-prediction_b200'), os.path.join(ROOT,'oracle','_ref')]\n"
        "os.environ['WF_B200_PRECISION']='fp32'\n"
        "m=types.ModuleType('easydict')\n"
        "class EasyDict(dict):\n"
        "    def __init__(s,d=None,**k):\n"
        "        super().__init__()\n"
        "        for a,b in {**(d or {}),**k}.items(): s[a]=EasyDict(b) if isinstance(b,dict) else b\n"
        "    __getattr__=dict.__getitem__\n"
        "m.EasyDict=EasyDict; sys.modules['easydict']=m\n"
        "import numpy as np, torch\n"
        f"os.chdir({str(wd)!r})\n"
        "os.path.exists('datasets') or os.symlink(os.path.join(ROOT,'oracle','_ref','datasets'),'datasets')\n"
        "torch.manual_seed(1); np.random.seed(1); random.seed(1)\n"
        "import evaluate as ev, eval.ap_calculator as apc\n"
        "cap={}\n"
        "o=apc.APCalculator.output_accuracy\n"
        "def f(self): o(self); cap.update({k:float(v) for k,v in self.ap_dict.items()})\n"
        "apc.APCalculator.output_accuracy=f\n"
        "ev.evaluate_with_ap_calculator(); print('RESULT '+json.dumps(cap))\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")]
    assert r.returncode == 0 and lines, r.stdout[-2000:] + r.stderr[-3000:]
    ours = json.loads(lines[-1][7:])
    refd = ref["ap_dict"]
    print("reference:", refd, "\ndrop-in:", ours)
    # evaluate.py draws a fresh random point_pool_proj in each arm (strict=False drops the checkpoint's) from the
    # same seed on the CPU generator, and resamples the test clouds with the same numpy seed: identical inputs and weights
    for k in ("tp_corners", "tp_fp_corners", "tp_fn_corners", "tp_edges", "tp_fp_edges", "tp_fn_edges"):
        assert ours[k] == refd[k], (k, ours[k], refd[k])
    assert abs(ours["distance"] - refd["distance"]) <= 1e-4 * max(1.0, abs(refd["distance"]))
    assert abs(ours["wed"] - refd["wed"]) <= 1e-4 * max(1.0, abs(refd["wed"]))
