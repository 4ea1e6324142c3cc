"""
Generates tests/golden/ap_calculator.npz by running the UNMODIFIED reference `eval/ap_calculator.py`
(imported from /root/reference, authoring container only) on wireframes from the reference's own test
split plus perturbed "predictions".

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_ap.py

Each case stores the six arrays evaluate.py:99-107 puts in the batch dict and the eight accumulators the
reference adds for that sample (eval/ap_calculator.py:262-272); the file also holds the final
output_accuracy() dictionaries of a run over all cases, per distance threshold.
"""
import contextlib
import glob
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, REF)

from eval.ap_calculator import APCalculator  # noqa: E402  (reference)
from datasets.building3d import load_wireframe  # noqa: E402  (reference)

KEYS = ("tp_corners", "tp_fp_corners", "tp_fn_corners", "distance", "tp_edges", "wed", "tp_fp_edges", "tp_fn_edges")
FINAL = ("average_corner_offset", "average_wed", "corners_precision", "corners_recall", "corners_f1",
         "edges_precision", "edges_recall", "edges_f1")


def edge_points(v, e):
    """evaluate.py:87-98 verbatim semantics (the reference has no function for it)."""
    if len(e) == 0:
        return np.empty((0, 2, 3))
    pts = np.stack((v[e[:, 0]], v[e[:, 1]]), axis=1)
    return pts[np.arange(pts.shape[0])[:, np.newaxis], np.flip(np.argsort(pts[:, :, -1]), axis=1)]


def all_pairs(c):
    return np.array([[i, j] for i in range(c) for j in range(i + 1, c)], dtype=np.int64).reshape(-1, 2)


def make_cases():
    rng = np.random.default_rng(20261018)
    files = sorted(glob.glob(os.path.join(REF, "datasets/test/wireframe/*.obj")))
    cases = []
    for k, f in enumerate(files):
        v, e = load_wireframe(f)
        v = v - v.mean(axis=0)
        v = (v / np.abs(v).max()).astype(np.float32)          # same scale as the loader's unit-ball normalisation
        e = e.astype(np.int64)
        c = len(v)
        kind = k % 6
        if kind == 0:      # perfect prediction
            pv, pe = v.copy(), e.copy()
        elif kind == 1:    # noisy corners, a few wrong and missing edges, V=64 slots with an unused tail
            pv = np.concatenate((v + rng.normal(0, 0.02, v.shape), rng.uniform(-1, 1, (64 - c, 3)))).astype(np.float32)
            keep = rng.random(len(e)) > 0.2
            extra = all_pairs(c)[rng.choice(c * (c - 1) // 2, size=5, replace=False)]
            pe = np.concatenate((e[keep], extra))
        elif kind == 2:    # every pair predicted (the untrained-model case): many more predictions than labels
            pv = (v + rng.normal(0, 0.05, v.shape)).astype(np.float32)
            pe = all_pairs(c)
        elif kind == 3:    # no edge above threshold: corners-only branch
            pv = (v + rng.normal(0, 0.03, v.shape)).astype(np.float32)
            pe = np.zeros((0, 2), dtype=np.int64)
        elif kind == 4:    # duplicated predicted corners, fewer predictions than labels
            pv = (v + rng.normal(0, 0.01, v.shape)).astype(np.float32)
            pv[1] = pv[0]
            pe = e[: max(2, len(e) // 3)].copy()
        else:              # coarse noise: with thresh 0.1 most edges stay unmatched
            pv = (v + rng.normal(0, 0.08, v.shape)).astype(np.float32)
            pe = np.concatenate((e, all_pairs(c)[:7]))
        cases.append(dict(pv=pv, pe=pe, pev=edge_points(pv, pe), gv=v, ge=e, gev=edge_points(v, e)))
    # one synthetic large case at the model's maximum: 64 corners, all 2016 pairs against 90 label edges
    gv = rng.uniform(-1, 1, (64, 3)).astype(np.float32)
    ge = np.unique(np.sort(rng.integers(0, 64, (90, 2)), axis=1), axis=0)
    ge = ge[ge[:, 0] != ge[:, 1]].astype(np.int64)
    pv = (gv + rng.normal(0, 0.02, gv.shape)).astype(np.float32)
    cases.append(dict(pv=pv, pe=all_pairs(64), pev=edge_points(pv, all_pairs(64)), gv=gv, ge=ge, gev=edge_points(gv, ge)))
    # nothing within 0.1: the reference dies in np.min of an empty matrix (eval/ap_calculator.py:227)
    gv = rng.uniform(-1, 1, (6, 3)).astype(np.float32)
    ge = np.array([[0, 1], [1, 2], [2, 3], [3, 4], [4, 5], [0, 5]], dtype=np.int64)
    pv = (gv + 0.6).astype(np.float32)
    cases.append(dict(pv=pv, pe=ge.copy(), pev=edge_points(pv, ge), gv=gv, ge=ge, gev=edge_points(gv, ge)))
    return cases


def run_reference(case, calc):
    batch = {
        "predicted_vertices": case["pv"][np.newaxis, :].copy(),
        "predicted_edges": case["pe"][np.newaxis, :].copy(),
        "pred_edges_vertices": case["pev"].reshape((1, -1, 2, 3)).copy(),
        "wf_vertices": case["gv"][np.newaxis, :].copy(),
        "wf_edges": case["ge"][np.newaxis, :].copy(),
        "wf_edges_vertices": case["gev"].reshape((1, -1, 2, 3)).copy(),
    }
    before = {k: calc.ap_dict[k] for k in KEYS}
    with contextlib.redirect_stdout(io.StringIO()):
        calc.compute_metrics(batch)
    return np.array([float(calc.ap_dict[k] - before[k]) for k in KEYS])


def main():
    cases = make_cases()
    out = {"n_cases": np.array(len(cases))}
    for i, c in enumerate(cases):
        for k, a in c.items():
            out[f"c{i}_{k}"] = a
    for thresh in (1.0, 0.1):
        calc = APCalculator(distance_thresh=thresh)
        tag = "t1" if thresh == 1.0 else "t01"
        for i, c in enumerate(cases):
            try:
                out[f"c{i}_{tag}"] = run_reference(c, calc)
            except ValueError as err:          # zero matched edges -> np.min of an empty matrix (:227)
                out[f"c{i}_{tag}"] = np.full(len(KEYS), np.nan)
                out[f"c{i}_{tag}_error"] = np.array(str(err))
        with contextlib.redirect_stdout(io.StringIO()):
            calc.output_accuracy()
        out[f"final_{tag}"] = np.array([float(calc.ap_dict[k]) for k in FINAL])
        out[f"totals_{tag}"] = np.array([float(calc.ap_dict[k]) for k in KEYS])
    np.savez_compressed(os.path.join(HERE, "ap_calculator.npz"), **out)
    for i in range(len(cases)):
        print(i, cases[i]["pe"].shape, cases[i]["ge"].shape, out[f"c{i}_t1"], out[f"c{i}_t01"])
    print(out["final_t1"], out["final_t01"])


if __name__ == "__main__":
    main()
