"""CPU: the evaluation post-processing oracle (oracle/ap_oracle.py) against the golden file made by the
unmodified reference `eval/ap_calculator.py` (tests/golden/make_golden_ap.py) and against the installed scipy."""
import os

import numpy as np
import pytest
from scipy.spatial.distance import cdist

from oracle import ap_oracle as ao

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ap_calculator.npz"))
N_CASES = int(GOLD["n_cases"])
INT_KEYS = ("tp_corners", "tp_fp_corners", "tp_fn_corners", "tp_edges", "tp_fp_edges", "tp_fn_edges")


def case_arrays(i):
    return [GOLD[f"c{i}_{k}"] for k in ("pv", "pe", "pev", "gv", "ge", "gev")]


def check_sample(got, want):
    want = dict(zip(ao.KEYS, want))
    for k in INT_KEYS:
        assert got[k] == int(want[k]), k
    for k in ("distance", "wed"):
        assert got[k] == pytest.approx(want[k], rel=1e-12, abs=1e-14), k


@pytest.mark.parametrize("tag,thresh", [("t1", 1.0), ("t01", 0.1)])
def test_oracle_matches_reference_golden(tag, thresh):
    per_sample = []
    for i in range(N_CASES):
        want = GOLD[f"c{i}_{tag}"]
        if np.isnan(want).all():
            with pytest.raises(ValueError, match="zero-size array"):
                ao.sample_metrics(*case_arrays(i), thresh)
            assert "zero-size array" in str(GOLD[f"c{i}_{tag}_error"])
            continue
        got = ao.sample_metrics(*case_arrays(i), thresh)
        check_sample(got, want)
        per_sample.append(got)
    tot = ao.accumulate(per_sample, last_batch_size=1)
    final = dict(zip(("average_corner_offset", "average_wed", "corners_precision", "corners_recall", "corners_f1",
                      "edges_precision", "edges_recall", "edges_f1"), GOLD[f"final_{tag}"]))
    for k, v in final.items():
        assert tot[k] == pytest.approx(v, rel=1e-12), k


def test_cdist_restatement_is_bit_equal_to_scipy():
    rng = np.random.default_rng(0)
    for n, m in ((1, 1), (7, 5), (64, 90), (300, 17)):
        a = rng.normal(size=(n, 3)).astype(np.float32)
        b = rng.normal(size=(m, 3))
        assert np.array_equal(ao.cdist_euclid(a, b), cdist(a, b))
    assert ao.cdist_euclid(np.zeros((0, 3)), np.zeros((4, 3))).shape == (0, 4)


def test_hausdorff_restatement_matches_the_scipy_formulation():
    rng = np.random.default_rng(1)
    for dtype in (np.float32, np.float64):
        p = rng.uniform(-1, 1, (23, 2, 3)).astype(dtype)
        t = rng.uniform(-1, 1, (9, 2, 3)).astype(np.float32)
        pts = ao.line_samples(np.concatenate((p, t)), 20)
        d = cdist(pts[:23].reshape(-1, 3), pts[23:].reshape(-1, 3)).reshape(23, 20, 9, 20).transpose(0, 2, 1, 3)
        want = np.maximum(d.min(-1).max(-1), d.min(-2).max(-1))
        assert np.array_equal(ao.hausdorff_lines(p, t), want)
    assert ao.hausdorff_lines(np.zeros((0, 2, 3)), t).size == 0


def test_edge_points_puts_the_higher_endpoint_first():
    v = np.array([[0, 0, 1], [1, 0, 2], [2, 0, 2]], dtype=np.float32)
    e = np.array([[0, 1], [1, 2]])
    pts = ao.edge_points(v, e)
    assert np.array_equal(pts[0], v[[1, 0]])
    assert np.array_equal(pts[1], v[[2, 1]])     # equal z: argsort is stable, flip puts the second endpoint first
    assert ao.edge_points(v, np.zeros((0, 2), dtype=np.int64)).shape == (0, 2, 3)
