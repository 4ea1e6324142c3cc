"""Host-side planner for the side jobs of the tensor-core GEMM launches (include/wf_b200.h: wf_side_seg).

The encoder's wide GEMMs (models/PointNetEncoder.py:37-45) are bound by the tensor pipe, the LayerNorm+ReLU passes between
them (:38-39, and their backward) by HBM.  The per-point MLP is therefore run over row chunks, and every GEMM launch carries
LayerNorm rows of OTHER chunks whose inputs already exist: a LayerNorm item L is *available* once the launch that produces its
input has been issued (same stream: it has completed when a later launch starts) and has a *deadline*: the first launch that
consumes its output.  Between the two it may be cut into row segments and dealt to any launches.

`plan` spreads every item over its window in proportion to the launches' durations (so it is always finished in time and
never piles up on the last launch of the window) and then pulls later work forward into launches that still have spare
bandwidth.  Pure Python, deterministic, no device access -- tests/test_sidesched.py checks the invariants on the CPU."""
from __future__ import annotations

from dataclasses import dataclass, field
from functools import lru_cache
from typing import Dict, List, Sequence, Tuple

SIDE_MAX = 4                 # include/wf_b200.h WF_SIDE_MAX
ROW_ALIGN = 128              # segments start and end on LayerNorm row blocks (lnb::CS_R)
CHUNK_ALIGN = 256            # chunks start on the GEMM's 256-row cluster tiles


@dataclass
class Item:
    key: object              # caller's identifier
    r0: int                  # rows [r0, r1) of the caller's tensors
    r1: int
    bytes_per_row: float     # HBM bytes the pass moves per row
    avail: int               # index of the launch that produces its input (-1: available from the start)
    deadline: int            # index of the first launch that needs its output (len(launches): none)
    done: int = field(default=0)          # rows already dealt

    @property
    def left(self) -> int:
        return (self.r1 - self.r0) - self.done


def split_rows(m: int, n_chunks: int, align: int = CHUNK_ALIGN) -> List[Tuple[int, int]]:
    """[0, m) in at most n_chunks nearly equal ranges whose starts are multiples of `align`."""
    n_chunks = max(1, min(n_chunks, (m + align - 1) // align))
    per = -(-m // n_chunks)
    per = -(-per // align) * align
    out, r = [], 0
    while r < m:
        out.append((r, min(m, r + per)))
        r += per
    return out


def _round_rows(rows: float, left: int) -> int:
    """Rows to take now: a multiple of ROW_ALIGN, at most `left`; a remainder below one block goes along."""
    take = int(-(-rows // ROW_ALIGN) * ROW_ALIGN) if rows > 0 else 0
    take = min(take, left)
    if 0 < left - take < ROW_ALIGN:
        take = left
    return take


def plan(durations: Sequence[float], items: List[Item], side_bw: float) -> Tuple[List[list], Dict[int, list]]:
    """durations[j]: estimated seconds of launch j; side_bw: HBM bytes/s a side job sustains beside the GEMM.
    Returns (side, pre): side[j] = [(key, r0, r1), ...] (<= SIDE_MAX) segments carried by launch j; pre[j] = segments that must
    run as stand-alone kernels BEFORE launch j (items without any launch inside their window)."""
    n = len(durations)
    side: List[list] = [[] for _ in range(n)]
    pre: Dict[int, list] = {}
    for it in items:
        it.done = 0
    for j in range(n + 1):
        # whatever is due now and unfinished cannot ride any more
        for it in items:
            if it.deadline == j and it.left > 0:
                pre.setdefault(j, []).append((it.key, it.r0 + it.done, it.r1))
                it.done = it.r1 - it.r0
        if j == n:
            break
        elig = sorted((it for it in items if it.avail < j < it.deadline and it.left > 0), key=lambda t: (t.deadline, t.avail))
        budget = durations[j] * side_bw
        taken = []
        for it in elig[:SIDE_MAX]:
            window = sum(durations[k] for k in range(j, min(it.deadline, n)))
            share = it.left * durations[j] / window if window > 0 else it.left
            take = it.left if it.deadline == j + 1 else _round_rows(share, it.left)
            taken.append([it, take])
            budget -= take * it.bytes_per_row
        # spare bandwidth: pull the earliest-deadline work forward
        for rec in taken:
            it, take = rec
            if budget <= 0:
                break
            extra = _round_rows(min(it.left - take, budget / it.bytes_per_row), it.left - take)
            # leave nothing smaller than one block behind
            if extra > 0:
                rec[1] = take + extra
                budget -= extra * it.bytes_per_row
        for it, take in taken:
            if take > 0:
                side[j].append((it.key, it.r0 + it.done, it.r0 + it.done + take))
                it.done += take
    return side, pre


@lru_cache(maxsize=64)
def _plan_cached(durations: tuple, specs: tuple, side_bw: float):
    items = [Item(*sp) for sp in specs]
    side, pre = plan(durations, items, side_bw)
    return tuple(tuple(s) for s in side), {j: tuple(v) for j, v in pre.items()}


def plan_cached(durations: Sequence[float], items: List[Item], side_bw: float):
    """`plan` memoised on its inputs: a training step asks for the same two plans (forward, backward) every time, and the ~0.2 ms
    of Python they cost sat between the first kernels of the step and the first GEMM launch, with the device idle.  The results
    are shared: callers must not modify them."""
    specs = tuple((it.key, it.r0, it.r1, it.bytes_per_row, it.avail, it.deadline) for it in items)
    return _plan_cached(tuple(durations), specs, float(side_bw))
