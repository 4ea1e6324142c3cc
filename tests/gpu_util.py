"""Helpers shared by the GPU parity tests."""
import numpy as np
import torch


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max(|b|) -- scale-relative max error."""
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    den = float(b.abs().max())
    return float((a - b).abs().max()) / (den if den > 0 else 1.0)


def assert_close(a, b, tol, what=""):
    e = rel_err(a, b)
    assert e <= tol, f"{what}: scale-relative max error {e:.3e} > {tol:.1e}"
    return e


def scipy_pairs(cost):
    from scipy.optimize import linear_sum_assignment
    return linear_sum_assignment(cost)


def col_to_pairs(col, nr, nc):
    """col_of_row (host int array) -> scipy-style (rows, cols)."""
    c = np.asarray(col[:nr])
    rows = np.nonzero(c >= 0)[0] if nc < nr else np.arange(nr)
    return rows.astype(np.int64), c[rows].astype(np.int64)


def record(test, **vals):
    """Append the errors a parity test OBSERVED to gpurun_out/parity_observed.jsonl (when that directory exists): the
    asserted tolerances are set from these records (<= 10x observed) and the file is copied to profiles/ per round."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = os.path.join(root, "gpurun_out")
    if not os.path.isdir(d):
        return
    with open(os.path.join(d, "parity_observed.jsonl"), "a") as f:
        f.write(json.dumps({"test": test, **{k: (float(v) if isinstance(v, (int, float)) else v) for k, v in vals.items()}}) + "\n")
