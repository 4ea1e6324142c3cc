"""CPU: host-side helpers of the step that need no device -- the per-step zero-filled gradient arena (ops.zeros_f32) and
the rule by which WireframeLoss decides whether its matching may run beside the edge head (_overlap_events)."""
import torch

from wf_b200 import ops


def test_zero_arena_serves_views_of_one_buffer_per_step():
    ops._ZERO_ARENAS.clear()
    ops.new_step()
    a = ops.zeros_f32(3, 5, device="cpu")
    b = ops.zeros_f32(7, device="cpu")
    assert a.shape == (3, 5) and b.shape == (7,) and float(a.sum()) == 0.0
    assert a.untyped_storage().data_ptr() != b.untyped_storage().data_ptr()          # first step: nothing learned yet
    a.add_(1.0); b.add_(2.0)
    ops.new_step()                                                                    # second step: one buffer, sized by the first
    c = ops.zeros_f32(3, 5, device="cpu")
    d = ops.zeros_f32(7, device="cpu")
    assert c.untyped_storage().data_ptr() == d.untyped_storage().data_ptr()
    assert float(c.sum()) == 0.0 and float(d.sum()) == 0.0
    assert c.data_ptr() % 16 == 0 and d.data_ptr() % 16 == 0                          # 256-byte slots inside the buffer
    c.add_(3.0)
    assert float(d.sum()) == 0.0                                                      # views do not overlap
    e = ops.zeros_f32(1000, device="cpu")                                             # more than the previous step asked for
    assert e.untyped_storage().data_ptr() != c.untyped_storage().data_ptr() and float(e.sum()) == 0.0
    ops.new_step()                                                                    # a buffer is never zeroed twice:
    f = ops.zeros_f32(3, 5, device="cpu")
    assert f.untyped_storage().data_ptr() != c.untyped_storage().data_ptr()           # last step's views keep their values
    assert float(c.sum()) == 45.0 and float(f.sum()) == 0.0
    # without step marks every request is an individual allocation again only once the arena is exhausted
    g = ops.zeros_f32(0, device="cpu")
    assert g.numel() == 0


def test_loss_overlap_rule():
    from losses.WireframeLoss import WireframeLoss
    crit = WireframeLoss()
    pv = torch.zeros(2, 4, 3)
    tgt = {"vertices": torch.zeros(2, 4, 3), "vertex_counts": torch.tensor([2, 3])}
    assert crit._overlap_events({"vertices": pv}, tgt) is None                        # predictions carry no ready-event
    pv._wf_ready = "EV_PRED"
    assert crit._overlap_events({"vertices": pv}, tgt) is None                        # fresh, untagged targets: main stream
    assert crit._overlap_events({"vertices": pv}, tgt) == ["EV_PRED"]                 # the same tensors again: known complete
    tgt["vertices"].add_(1.0)                                                         # in-place change -> unknown again
    assert crit._overlap_events({"vertices": pv}, tgt) is None
    tagged = {k: v.clone() for k, v in tgt.items()}
    for v in tagged.values():
        v._wf_ready = "EV_TGT"
    assert crit._overlap_events({"vertices": pv}, tagged) == ["EV_PRED", "EV_TGT", "EV_TGT"]
    half = {"vertices": tagged["vertices"], "vertex_counts": tgt["vertex_counts"].clone()}
    assert crit._overlap_events({"vertices": pv}, half) is None                       # one untagged, unseen target is enough


def test_loss_overlap_rule_is_by_identity_not_by_address():
    """fresh target tensors of the next batch are routinely handed the previous batch's recycled address with
    version 0.  They must NOT be taken for "the tensors of the previous call"."""
    from losses.WireframeLoss import WireframeLoss
    crit = WireframeLoss()
    pv = torch.zeros(2, 4, 3)
    pv._wf_ready = "EV_PRED"
    seen_addresses = set()
    hits = 0
    for step in range(6):
        # a per-step function that builds its targets afresh: the allocator recycles the storage of the previous step's
        tgt = {"vertices": torch.full((2, 4, 3), float(step)), "vertex_counts": torch.tensor([2 + step % 2, 3])}
        seen_addresses.add((tgt["vertices"].data_ptr(), tgt["vertex_counts"].data_ptr()))
        if crit._overlap_events({"vertices": pv}, tgt) is not None:
            hits += 1
        del tgt
    assert hits == 0, "fresh untagged targets were mistaken for the previous call's tensors"
    # (on this allocator the 6 batches typically produce 1-2 distinct address pairs: the aliasing scenario is the common case)


def test_host_counts_never_keyed_on_an_address():
    """PointCloudToWireframe._host_counts must return the counts of THE tensor it is given."""
    from models.PointCloudToWireframe import PointCloudToWireframe
    m = PointCloudToWireframe.__new__(PointCloudToWireframe)      # no parameters needed for the host helper
    m._count_cache = None
    for step in range(6):
        vals = [2 + step, 3 + step, 4]
        t = torch.tensor(vals)                                     # fresh tensor per batch, recycled address
        assert m._host_counts(t) == vals
        del t
    assert m._host_counts([5, 6]) == [5, 6] and m._host_counts((7,)) == [7]
    # the host tag set by wf_b200.targets is honoured only at the version it was made for
    class FakeCuda(torch.Tensor):
        is_cuda = True
    t = torch.tensor([9, 9]).as_subclass(FakeCuda)
    t._wf_host_counts = (t._version, (4, 5))
    assert m._host_counts(t) == [4, 5]
    t.add_(1)                                                      # modified after tagging: tag is stale, values are read
    assert m._host_counts(t) == [10, 10]
    assert m._host_counts(t) == [10, 10]                           # same object, same version: remembered
    t.add_(1)
    assert m._host_counts(t) == [11, 11]
