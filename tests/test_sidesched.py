"""CPU: the side-job planner (wf_b200/sidesched.py) -- every LayerNorm item is dealt exactly once, inside its window
(after the launch that produces its input, before the launch that consumes its output), in row-block aligned segments,
at most WF_SIDE_MAX per launch; items without a window fall back to stand-alone passes."""
import random

from wf_b200 import sidesched as ss


def _check(durs, items, side, pre):
    n = len(durs)
    cover = {id(it): [] for it in items}
    by_key = {it.key: it for it in items}
    for j, segs in enumerate(side):
        assert len(segs) <= ss.SIDE_MAX
        for key, r0, r1 in segs:
            it = by_key[key]
            assert it.avail < j < it.deadline, (key, j, it.avail, it.deadline)
            cover[id(it)].append((r0, r1))
    for j, segs in pre.items():
        for key, r0, r1 in segs:
            it = by_key[key]
            assert j == it.deadline
            cover[id(it)].append((r0, r1))
    for it in items:
        segs = sorted(cover[id(it)])
        assert segs and segs[0][0] == it.r0 and segs[-1][1] == it.r1, (it.key, segs)
        for (a0, a1), (b0, b1) in zip(segs, segs[1:]):
            assert a1 == b0                                   # contiguous, no overlap
        for r0, r1 in segs:
            assert r1 > r0 and (r0 - it.r0) % ss.ROW_ALIGN == 0


def test_split_rows():
    assert ss.split_rows(640000, 2) == [(0, 320000), (320000, 640000)]
    ch = ss.split_rows(640000, 3)
    assert ch[0][0] == 0 and ch[-1][1] == 640000 and all(a % 256 == 0 for a, _ in ch) and all(b == c for (_, b), (c, _) in zip(ch, ch[1:]))
    assert ss.split_rows(100, 4) == [(0, 100)]
    assert ss.split_rows(513, 2) == [(0, 512), (512, 513)]


def test_forward_plan_of_the_encoder():
    # layer-major forward over 2 chunks: G[li][c] for three layers, then the pooling GEMM per chunk
    chunks = ss.split_rows(640000, 2)
    flops = [2 * 512 * 1024, 2 * 1024 * 2048, 2 * 2048 * 1024, 2 * 1024 * 512]
    widths = [1024, 2048, 1024]
    idx = {}
    durs = []
    for li in range(4):
        for c, (r0, r1) in enumerate(chunks):
            idx[(li, c)] = len(durs)
            durs.append((r1 - r0) * flops[li] / 1.3e15)
    items = [ss.Item((li, c), r0, r1, widths[li] * 4, idx[(li, c)], idx[(li + 1, c)]) for li in range(3) for c, (r0, r1) in enumerate(chunks)]
    side, pre = ss.plan(durs, items, 2.0e12)
    _check(durs, items, side, pre)
    assert not pre                                             # every item has a window
    assert not side[0]                                         # nothing exists yet under the first launch


def test_items_without_a_window_run_stand_alone_and_random_plans_are_valid():
    durs = [1e-3] * 4
    items = [ss.Item("a", 0, 1000, 4096, -1, 0), ss.Item("b", 0, 777, 4096, 0, 1), ss.Item("c", 256, 5000, 8192, 0, 4)]
    side, pre = ss.plan(durs, items, 2.0e12)
    _check(durs, items, side, pre)
    assert pre[0] == [("a", 0, 1000)] and pre[1] == [("b", 0, 777)]
    rnd = random.Random(0)
    for trial in range(200):
        n = rnd.randint(1, 12)
        durs = [rnd.uniform(1e-4, 2e-3) for _ in range(n)]
        items = []
        for k in range(rnd.randint(1, 8)):
            a = rnd.randint(-1, n - 1)
            d = rnd.randint(a + 1, n)
            r0 = 256 * rnd.randint(0, 50)
            items.append(ss.Item(k, r0, r0 + rnd.randint(1, 400000), rnd.choice([4096, 8192, 6144, 12288]), a, d))
        side, pre = ss.plan(durs, items, rnd.choice([5e11, 2e12, 1e13]))
        _check(durs, items, side, pre)


def test_plan_cached_equals_plan_and_is_not_disturbed_by_reuse():
    """plan_cached memoises the plan on (durations, item windows, bandwidth): same result as plan(), also on the second request
    (Item.done of the caller's objects is not part of the key and is left alone)."""
    from wf_b200 import sidesched as ss
    durs = [1.0e-3, 0.5e-3, 2.0e-3, 1.0e-3, 1.5e-3, 0.7e-3, 1.1e-3, 0.9e-3]
    mk = lambda: [ss.Item((li, c), c * 320000, (c + 1) * 320000, 4096.0 * (1 + li % 2), li * 2 + c, (li + 1) * 2 + c)
                  for li in range(3) for c in range(2)]
    ref_side, ref_pre = ss.plan(durs, mk(), 2.0e12)
    for _ in range(2):
        items = mk()
        side, pre = ss.plan_cached(durs, items, 2.0e12)
        assert [list(x) for x in side] == ref_side and {k: list(v) for k, v in pre.items()} == ref_pre
        assert all(it.done == 0 for it in items)
    # a different window is a different plan (the key covers the items' windows, not only their rows)
    wide = mk()
    for it in wide:
        it.deadline += 2
    side2, _ = ss.plan_cached(durs, wide, 2.0e12)
    ref2, _ = ss.plan(durs, [ss.Item(it.key, it.r0, it.r1, it.bytes_per_row, it.avail, it.deadline) for it in wide], 2.0e12)
    assert [list(x) for x in side2] == ref2 and ref2 != ref_side
