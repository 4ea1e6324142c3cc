"""Generates tests/golden/targets.npz in the AUTHORING container (needs /root/reference): the reference's own
models.utils.create_edge_labels_from_edge_set is imported unmodified and driven by the loop structure of train.py:48-88,112-115
(that code is inline in train_overfit_model and cannot be imported on its own; oracle/targets_oracle.py restates it line by
line).  Usage: PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_targets.py"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import targets_oracle as to  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_utils", "/root/reference/models/utils.py")
ref_utils = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_utils)

out = {}
cases = [(11, 3, 8), (12, 5, 20), (13, 4, 64)]
out["cases"] = np.array(cases)
for seed, B, V in cases:
    verts, edges = to.make_case(seed, B, V)
    t = to.prepare_targets(verts, edges, V, label_fn=ref_utils.create_edge_labels_from_edge_set)
    for k, v in t.items():
        out[f"{seed}/{k}"] = v.numpy()
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "targets.npz"), **out)
print("wrote targets.npz", {k: v.shape for k, v in out.items()})
