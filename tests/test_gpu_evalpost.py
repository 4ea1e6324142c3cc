"""GPU parity of the evaluation post-processing: the fp64 kernels behind
`eval/ap_calculator.py` against the oracle (oracle/ap_oracle.py), the installed scipy and the golden file made by
the unmodified reference class.  Distances and assignments are required BIT/INDEX-identical."""
import contextlib
import io
import os

import numpy as np
import pytest
from scipy.optimize import linear_sum_assignment
from scipy.spatial.distance import cdist

from oracle import ap_oracle as ao

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ap_calculator.npz"))
N_CASES = int(GOLD["n_cases"])
INT_KEYS = ("tp_corners", "tp_fp_corners", "tp_fn_corners", "tp_edges", "tp_fp_edges", "tp_fn_edges")


@pytest.fixture(scope="module")
def ep():
    from wf_b200 import evalpost
    return evalpost


def test_hausdorff_bit_equal_ragged_batch(ep):
    rng = np.random.default_rng(3)
    ps, ts = [], []
    for n, m, dt in ((23, 9, np.float32), (1, 1, np.float64), (0, 5, np.float32), (2016, 90, np.float32), (40, 130, np.float64)):
        ps.append(rng.uniform(-1, 1, (n, 2, 3)).astype(dt))
        ts.append(rng.uniform(-1, 1, (m, 2, 3)).astype(np.float32))
    ps[1][0, 1] = ps[1][0, 0]                                  # a degenerate (zero-length) segment
    got = ep.hausdorff_lines_batched(ps, ts)
    for g, p, t in zip(got, ps, ts):
        want = ao.hausdorff_lines(p, t)
        assert g.shape == want.shape and np.array_equal(g, want)
    # other sample counts go through the generic kernel
    g7 = ep.hausdorff_lines_batched(ps[:1], ts[:1], samples=7)[0]
    assert np.array_equal(g7, ao.hausdorff_lines(ps[0], ts[0], samples=7))
    g64 = ep.hausdorff_lines_batched(ps[:1], ts[:1], samples=64)[0]
    assert np.array_equal(g64, ao.hausdorff_lines(ps[0], ts[0], samples=64))


def test_cdist_bit_equal_to_scipy(ep):
    rng = np.random.default_rng(4)
    a_list = [rng.normal(size=(n, 3)).astype(np.float32) for n in (5, 64, 0, 1, 300)]
    b_list = [rng.normal(size=(m, 3)) for m in (7, 90, 4, 1, 17)]
    for g, a, b in zip(ep.cdist_batched(a_list, b_list), a_list, b_list):
        assert g.shape == (len(a), len(b))
        if g.size:
            assert np.array_equal(g, cdist(a, b))


def test_lsap_f64_index_identical_to_scipy(ep):
    rng = np.random.default_rng(5)
    mats = []
    for k in range(400):
        nr, nc = (int(x) for x in rng.integers(1, 200, size=2))
        kind = k % 5
        if kind == 0:
            c = rng.uniform(0, 1, (nr, nc))
        elif kind == 1:
            c = rng.integers(0, 3, (nr, nc)).astype(np.float64)      # heavy ties
        elif kind == 2:
            c = np.full((nr, nc), float(rng.integers(0, 2)))          # constant
        elif kind == 3:
            c = rng.normal(size=(nr, nc))
            c[rng.random((nr, nc)) < 0.3] = np.inf                    # forbidden pairs, maybe still feasible
            c[np.arange(min(nr, nc)), np.arange(min(nr, nc))] = 0.0   # keep a finite assignment
        else:
            c = np.round(rng.uniform(0, 4, (nr, nc)), 1)
        mats.append(c)
    mats.append(rng.uniform(0, 2, (2016, 90)))                        # every pair of 64 corners vs 90 label edges
    mats.append(rng.uniform(0, 2, (90, 2016)))
    mats.append(rng.integers(0, 4, (2016, 64)).astype(np.float64))
    mats.append(np.zeros((0, 5)))
    mats.append(np.zeros((4, 0)))
    got = ep.lsap_batched_f64(mats)
    for (r, c), m in zip(got, mats):
        wr, wc = linear_sum_assignment(m)
        assert np.array_equal(r, wr) and np.array_equal(c, wc), m.shape


def test_hausdorff_assign_keeps_matrices_on_device(ep):
    rng = np.random.default_rng(8)
    ps = [rng.uniform(-1, 1, (n, 2, 3)).astype(np.float32) for n in (300, 7, 0, 2016)]
    ts = [rng.uniform(-1, 1, (m, 2, 3)).astype(np.float32) for m in (20, 40, 3, 64)]
    for (r, c, d), p, t in zip(ep.hausdorff_assign_batched(ps, ts), ps, ts):
        if len(p) == 0:
            assert r.size == 0 and c.size == 0 and d.size == 0
            continue
        h = ao.hausdorff_lines(p, t)
        wr, wc = linear_sum_assignment(h)
        assert np.array_equal(r, wr) and np.array_equal(c, wc) and np.array_equal(d, h[wr, wc])


def test_lsap_f64_errors_like_scipy(ep):
    bad = np.ones((5, 6)); bad[2, 3] = np.nan
    with pytest.raises(ValueError, match="invalid numeric"):
        ep.lsap_batched_f64([np.ones((3, 3)), bad])
    inf = np.ones((4, 4)); inf[1, :] = np.inf
    with pytest.raises(ValueError, match="infeasible"):
        ep.lsap_batched_f64([inf])
    neg = np.ones((3, 4)); neg[0, 0] = -np.inf
    with pytest.raises(ValueError, match="invalid numeric"):
        ep.lsap_batched_f64([neg])


def golden_batch(indices):
    return {
        "predicted_vertices": [GOLD[f"c{i}_pv"].copy() for i in indices],
        "predicted_edges": [GOLD[f"c{i}_pe"].copy() for i in indices],
        "pred_edges_vertices": [GOLD[f"c{i}_pev"].copy() for i in indices],
        "wf_vertices": [GOLD[f"c{i}_gv"].copy() for i in indices],
        "wf_edges": [GOLD[f"c{i}_ge"].copy() for i in indices],
        "wf_edges_vertices": [GOLD[f"c{i}_gev"].copy() for i in indices],
    }


def check_increment(after, before, want):
    got = {k: after[k] - before[k] for k in ao.KEYS}
    want = dict(zip(ao.KEYS, want))
    for k in INT_KEYS:
        assert int(got[k]) == int(want[k]), k
    assert got["distance"] == pytest.approx(want["distance"], rel=1e-12, abs=1e-14)
    assert got["wed"] == pytest.approx(want["wed"], rel=1e-6, abs=1e-9)      # float32 edge-length sums in the reference


@pytest.mark.parametrize("tag,thresh", [("t1", 1.0), ("t01", 0.1)])
def test_ap_calculator_matches_reference_golden(tag, thresh):
    from eval.ap_calculator import APCalculator
    calc = APCalculator(distance_thresh=thresh)
    for i in range(N_CASES):                                   # one sample per call, as evaluate.py:110 does
        want = GOLD[f"c{i}_{tag}"]
        before = dict(calc.ap_dict)
        if np.isnan(want).all():
            with pytest.raises(ValueError, match="zero-size array"):
                calc.compute_metrics(golden_batch([i]))
            assert {k: calc.ap_dict[k] for k in ao.KEYS} == {k: before[k] for k in ao.KEYS}
            continue
        calc.compute_metrics(golden_batch([i]))
        check_increment(calc.ap_dict, before, want)
    with contextlib.redirect_stdout(io.StringIO()) as text:
        calc.output_accuracy()
    assert "Edges F1" in text.getvalue()
    names = ("average_corner_offset", "average_wed", "corners_precision", "corners_recall", "corners_f1",
             "edges_precision", "edges_recall", "edges_f1")
    for k, v in zip(names, GOLD[f"final_{tag}"]):
        assert calc.ap_dict[k] == pytest.approx(v, rel=1e-6 if "wed" in k else 1e-12), k
    calc.reset()
    assert calc.ap_dict["tp_corners"] == 0 and calc.ap_dict["wed"] == 0


def test_ap_calculator_whole_batch_equals_per_sample_calls():
    """The batched phases (all samples in two device round trips) give the totals of sample-by-sample calls."""
    from eval.ap_calculator import APCalculator
    ok = [i for i in range(N_CASES) if not np.isnan(GOLD[f"c{i}_t1"]).all()]
    whole = APCalculator(distance_thresh=1.0)
    whole.compute_metrics(golden_batch(ok))
    want = GOLD["totals_t1"]
    check_increment(whole.ap_dict, {k: 0 for k in ao.KEYS}, want)
    assert whole.batch_size == len(ok)
    # the segments matched within the threshold were snapped onto their labels in the caller's arrays (:233-234)
    b = golden_batch([0]); APCalculator(distance_thresh=1.0).compute_metrics(b)
    assert np.array_equal(np.sort(b["pred_edges_vertices"][0].reshape(-1, 6), axis=0),
                          np.sort(b["wf_edges_vertices"][0].reshape(-1, 6), axis=0))


def test_module_level_functions(ep):
    from eval import ap_calculator as apc
    rng = np.random.default_rng(6)
    p = rng.uniform(-1, 1, (12, 2, 3)).astype(np.float32)
    t = rng.uniform(-1, 1, (5, 2, 3)).astype(np.float32)
    assert np.array_equal(apc.hausdorff_distance_line(p, t), ao.hausdorff_lines(p, t))
    assert apc.hausdorff_distance_line(p[:0], t).size == 0
    v = np.unique(t.reshape(-1, 3), axis=0)
    assert np.array_equal(apc.computer_edges(t, v), ao.index_edges(t, v))
    assert np.array_equal(apc.remove_corners(v, v[::2].copy()), ao.rows_not_in(v, v[::2].copy()))
    e = ao.index_edges(t, v)
    got = apc.graph_edit_distance(v.copy(), e.copy(), v.copy(), e.copy(), 0.25)
    assert got == pytest.approx(ao.edit_distance(v.copy(), e.copy(), v.copy(), e.copy(), 0.25), rel=1e-6)


def test_make_ap_batch_follows_evaluate_py(ep):
    import torch
    rng = np.random.default_rng(7)
    B, V = 3, 12
    counts = [5, 12, 8]
    idx = [[[i, j] for i in range(c) for j in range(i + 1, c)] for c in counts]
    max_e = max(len(x) for x in idx)
    probs = np.zeros((B, max_e), dtype=np.float32)
    for b in range(B):
        probs[b, :len(idx[b])] = rng.random(len(idx[b]))
    pred = {"vertices": torch.tensor(rng.normal(size=(B, V, 3)).astype(np.float32)).cuda(),
            "edge_probs": torch.tensor(probs).cuda(), "edge_indices": idx}
    gv = [torch.tensor(rng.normal(size=(c, 3)).astype(np.float32)) for c in counts]
    ge = [torch.tensor(np.array(idx[b])[rng.random(len(idx[b])) < 0.3].astype(np.float32)) for b in range(B)]
    batch = ep.make_ap_batch(pred, gv, ge)
    for b in range(B):
        want = ao.eval_sample_inputs(pred["vertices"][b].cpu().numpy(), idx[b], probs[b, :len(idx[b])],
                                     gv[b].numpy(), ge[b].numpy().astype(np.int64))
        keys = ("predicted_vertices", "predicted_edges", "pred_edges_vertices", "wf_vertices", "wf_edges", "wf_edges_vertices")
        for k, w in zip(keys, want):
            assert np.array_equal(batch[k][b], w), k
