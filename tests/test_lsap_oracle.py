"""Pins oracle/lsap_oracle.c against the installed scipy (the third-party solver the reference
calls at losses/WireframeLoss.py:236): identical index arrays, not merely equal cost."""
import numpy as np
import pytest
from scipy.optimize import linear_sum_assignment

from oracle import wireframe_oracle as wo


def _same(cost):
    r0, c0 = linear_sum_assignment(cost)
    r1, c1 = wo.lsap(cost)
    assert np.array_equal(r0, r1) and np.array_equal(c0, c1), (cost.shape, r0, c0, r1, c1)


def test_random_square_and_rect():
    rng = np.random.default_rng(0)
    for _ in range(4000):
        nr, nc = rng.integers(1, 20, size=2)
        _same(rng.uniform(0, 1, (nr, nc)).astype(np.float32))
    for _ in range(300):
        nr, nc = rng.integers(20, 65, size=2)
        _same(rng.normal(size=(nr, nc)))


def test_ties_small_integers():
    rng = np.random.default_rng(1)
    for _ in range(6000):
        nr, nc = rng.integers(1, 14, size=2)
        _same(rng.integers(0, 3, (nr, nc)).astype(np.float64))
    for n in range(1, 12):
        _same(np.zeros((n, n)))
        _same(np.ones((n, n + 3)))
        _same(np.ones((n + 3, n)))


def test_loss_style_dummy_columns():
    """Constant-per-row dummy columns as in losses/WireframeLoss.py:216-219."""
    rng = np.random.default_rng(2)
    for _ in range(3000):
        V = int(rng.integers(2, 40)); c = int(rng.integers(1, V + 1))
        e = rng.uniform(0, 1, (V, 1)).astype(np.float32)
        real = (rng.uniform(0, 2, (V, c)).astype(np.float32) + np.abs(e - 1)).astype(np.float32)
        _same(np.concatenate([real, np.repeat(e, V - c, axis=1)], axis=1))
    for _ in range(500):                                   # quantised -> heavy ties
        V = int(rng.integers(2, 24)); c = int(rng.integers(1, V + 1))
        e = (rng.integers(0, 5, (V, 1)) / 4).astype(np.float32)
        real = (rng.integers(0, 4, (V, c)) / 2 + np.abs(e - 1)).astype(np.float32)
        _same(np.concatenate([real, np.repeat(e, V - c, axis=1)], axis=1))


def test_inf_entries_and_errors():
    rng = np.random.default_rng(3)
    for _ in range(500):
        n = int(rng.integers(2, 10))
        c = rng.uniform(0, 1, (n, n))
        c[rng.uniform(size=(n, n)) < 0.3] = np.inf
        try:
            ref = linear_sum_assignment(c)
        except ValueError as e:
            with pytest.raises(ValueError, match=str(e)):
                wo.lsap(c)
            continue
        got = wo.lsap(c)
        assert np.array_equal(ref[0], got[0]) and np.array_equal(ref[1], got[1])
    for bad in (np.nan, -np.inf):
        c = np.ones((3, 3)); c[1, 1] = bad
        with pytest.raises(ValueError, match="invalid numeric"):
            wo.lsap(c)
    with pytest.raises(ValueError, match="infeasible"):
        wo.lsap(np.array([[1.0, np.inf], [2.0, np.inf]]))
    r, c = wo.lsap(np.zeros((0, 4)))
    assert r.size == 0 and c.size == 0


def test_loss_cost_matrix_c_matches_torch():
    import ctypes
    import torch
    rng = np.random.default_rng(4)
    lib = wo._lib()
    fp = ctypes.POINTER(ctypes.c_float)
    for _ in range(200):
        V = int(rng.integers(1, 33)); c = int(rng.integers(0, V + 1))
        pv = rng.uniform(-1, 1, (V, 3)).astype(np.float32)
        pe = rng.uniform(0, 1, (V,)).astype(np.float32)
        tv = rng.uniform(-1, 1, (V, 3)).astype(np.float32)
        ref = wo.loss_cost_matrix(torch.from_numpy(pv), torch.from_numpy(pe), torch.from_numpy(tv), c).numpy()
        out = np.empty((V, V), np.float32)
        lib.wfo_loss_cost_f32(pv.ctypes.data_as(fp), pe.ctypes.data_as(fp), tv.ctypes.data_as(fp), V, c,
                              out.ctypes.data_as(fp))
        assert np.array_equal(ref, out)
