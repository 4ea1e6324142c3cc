"""CPU: the HOST side of the drop-in `eval.ap_calculator.APCalculator` (batch phases, set logic, wireframe edit distance)
against the golden file of the unmodified reference class.  The three device primitives it calls (wf_b200.evalpost, fp64
CUDA kernels) are replaced here by the scipy functions they reproduce bit for bit on the GPU (tests/test_gpu_evalpost.py);
nothing in the product imports these stand-ins."""
import os

import numpy as np
import pytest
from scipy.optimize import linear_sum_assignment
from scipy.spatial.distance import cdist

from oracle import ap_oracle as ao

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ap_calculator.npz"))
N_CASES = int(GOLD["n_cases"])
INT_KEYS = ("tp_corners", "tp_fp_corners", "tp_fn_corners", "tp_edges", "tp_fp_edges", "tp_fn_edges")


@pytest.fixture()
def calc_cls(monkeypatch):
    from wf_b200 import evalpost
    e = np.zeros(0, dtype=np.int64)

    def assign(ps, ts, samples=20):
        out = []
        for p, t in zip(ps, ts):
            h = ao.hausdorff_lines(np.asarray(p), np.asarray(t), samples)
            r, c = linear_sum_assignment(h) if h.ndim == 2 and h.size else (e, e)
            out.append((r, c, h[r, c] if len(r) else np.zeros(0)))
        return out

    def cd(a_list, b_list):
        return [cdist(np.asarray(a, dtype=np.float64).reshape(-1, 3), np.asarray(b, dtype=np.float64).reshape(-1, 3))
                for a, b in zip(a_list, b_list)]

    monkeypatch.setattr(evalpost, "hausdorff_assign_batched", assign)
    monkeypatch.setattr(evalpost, "cdist_batched", cd)
    monkeypatch.setattr(evalpost, "lsap_batched_f64", lambda mats: [linear_sum_assignment(m) if m.size else (e, e) for m in mats])
    from eval.ap_calculator import APCalculator
    return APCalculator


def batch_of(indices):
    keys = (("predicted_vertices", "pv"), ("predicted_edges", "pe"), ("pred_edges_vertices", "pev"), ("wf_vertices", "gv"),
            ("wf_edges", "ge"), ("wf_edges_vertices", "gev"))
    return {k: [GOLD[f"c{i}_{s}"].copy() for i in indices] for k, s in keys}


@pytest.mark.parametrize("tag,thresh", [("t1", 1.0), ("t01", 0.1)])
def test_host_logic_matches_reference_golden(calc_cls, tag, thresh):
    calc = calc_cls(distance_thresh=thresh)
    for i in range(N_CASES):
        want = GOLD[f"c{i}_{tag}"]
        before = dict(calc.ap_dict)
        if np.isnan(want).all():
            with pytest.raises(ValueError, match="zero-size array"):
                calc.compute_metrics(batch_of([i]))
            continue
        calc.compute_metrics(batch_of([i]))
        got = {k: calc.ap_dict[k] - before[k] for k in ao.KEYS}
        for k in INT_KEYS:
            assert int(got[k]) == int(dict(zip(ao.KEYS, want))[k]), (i, k)
        assert got["distance"] == pytest.approx(dict(zip(ao.KEYS, want))["distance"], rel=1e-12, abs=1e-14)
        assert got["wed"] == pytest.approx(dict(zip(ao.KEYS, want))["wed"], rel=1e-6, abs=1e-9), i


def test_whole_batch_equals_sample_by_sample(calc_cls):
    ok = [i for i in range(N_CASES) if not np.isnan(GOLD[f"c{i}_t1"]).all()]
    whole = calc_cls(distance_thresh=1.0)
    whole.compute_metrics(batch_of(ok))
    for k, v in zip(ao.KEYS, GOLD["totals_t1"]):
        assert whole.ap_dict[k] == pytest.approx(v, rel=1e-6 if k == "wed" else 1e-12), k
