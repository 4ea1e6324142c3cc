"""Synthetic inputs of the reference's shapes: point clouds as datasets/building3d.py:109-126 emits them
(xyz centred and scaled to unit max-norm, RGBA/256, raw intensity) and targets as train.py:48-88,112-115 builds them.
Used by bench.py; same recipe (distributions) as the oracle generator in oracle/wireframe_oracle.py."""
from typing import Optional

import numpy as np
import torch


def make_counts(seed: int, B: int, V: int, min_count: int = 2, max_count: Optional[int] = None):
    """Only the GT vertex counts of make_inputs(seed, B, ...): they come from their own generator stream, so every
    rank can know every shard's counts (needed for the batch-global loss normalisers) without building the points."""
    max_count = V if max_count is None else max_count
    return np.random.Generator(np.random.PCG64(99991 + seed)).integers(min_count, max_count + 1, size=(B,)).tolist()


def make_inputs(seed: int, B: int, N: int, V: int, *, pad_frac: float = 0.0, norm_intensity: bool = False,
                min_count: int = 2, max_count: Optional[int] = None, dtype=torch.float32):
    rng = np.random.Generator(np.random.PCG64(1234 + seed))
    max_count = V if max_count is None else max_count
    xyz = rng.uniform(-1, 1, size=(B, N, 3))
    xyz = xyz - xyz.mean(axis=1, keepdims=True)
    xyz = xyz / np.linalg.norm(xyz, axis=2).max(axis=1)[:, None, None]
    rgba = rng.integers(0, 256, size=(B, N, 4)) / 256.0
    inten = rng.uniform(2e4, 6e4, size=(B, N, 1))
    if norm_intensity:
        inten = inten / 65536.0
    pts = np.concatenate([xyz, rgba, inten], axis=2)
    if pad_frac > 0:
        npad = int(N * pad_frac)
        if npad:
            pts[:, N - npad:, :] = 0.0
    counts = np.random.Generator(np.random.PCG64(99991 + seed)).integers(min_count, max_count + 1, size=(B,))
    tv = np.zeros((B, V, 3)); te = np.zeros((B, V))
    for b in range(B):
        c = int(counts[b])
        tv[b, :c] = rng.uniform(-0.5, 0.5, size=(c, 3))
        te[b, :c] = 1.0
    max_e = int(max(c * (c - 1) // 2 for c in counts)) if B else 0
    el = np.zeros((B, max_e))
    for b in range(B):
        c = int(counts[b]); e = c * (c - 1) // 2
        el[b, :e] = (rng.uniform(size=(e,)) < min(1.0, 3.0 / max(c, 1))).astype(np.float64)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dtype)
    counts_t = torch.from_numpy(counts.astype(np.int64))
    targets = {"vertices": t(tv), "vertex_existence": t(te), "edge_labels": t(el), "vertex_counts": counts_t}
    return t(pts), targets, counts_t
