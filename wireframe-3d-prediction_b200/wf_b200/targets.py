"""Target preparation and input staging for the training step (the code immediately before the hot path).

`prepare_targets` replaces the per-batch Python loops of train.py:48-88,112-115 (O(c^2) set look-ups per sample, one
`.item()` per edge endpoint) by one ragged concatenation on the host, one pinned host->device copy per array and ONE kernel
(wf_pack_targets).  `DevicePrefetcher` overlaps the host->device copy of batch i+1 with the step on batch i."""
from __future__ import annotations

from typing import Dict, Iterable, Iterator, List, Sequence

import torch

from . import ops
from ._lib import call


def prepare_targets(wf_vertices: Sequence[torch.Tensor], wf_edges: Sequence[torch.Tensor], max_vertices: int,
                    device) -> Dict[str, torch.Tensor]:
    """wf_vertices[b]: (n_b, 3) float; wf_edges[b]: (e_b, 2) vertex indices (float32 as the reference's collate_batch
    delivers them, datasets/building3d.py:180-183, or any integer dtype).  Returns the `targets` dict of
    losses/WireframeLoss.py:38 -- vertices (B,V,3), vertex_existence (B,V), edge_labels (B,max_E), vertex_counts (B,) int64 --
    with the values train.py:48-88,112-115 computes."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise ops._lib.WfError("prepare_targets runs on a CUDA device (there is no CPU path)")
    B, V = len(wf_vertices), int(max_vertices)
    counts = [int(v.shape[0]) for v in wf_vertices]
    if any(c > V for c in counts):
        raise ValueError(f"a sample has more than max_vertices={V} vertices")       # train.py:115 fails on the slice assignment
    v_off = [0] * (B + 1); e_off = [0] * (B + 1)
    for b in range(B):
        v_off[b + 1] = v_off[b] + counts[b]
        e_off[b + 1] = e_off[b] + int(wf_edges[b].shape[0])
    max_e = max((c * (c - 1) // 2 for c in counts), default=0)                       # train.py:80
    T, E = v_off[B], e_off[B]
    host_v = torch.empty(max(T, 1), 3, dtype=torch.float32, pin_memory=True)
    host_e = torch.empty(max(E, 1), 2, dtype=torch.float32, pin_memory=True)
    if T:
        torch.cat([v.reshape(-1, 3).to(torch.float32) for v in wf_vertices], out=host_v[:T])
    if E:
        torch.cat([e.reshape(-1, 2).to(torch.float32) for e in wf_edges if e.numel()], out=host_e[:E])
    host_off = torch.tensor([v_off, e_off], dtype=torch.int32).pin_memory()
    dv = host_v.to(dev, non_blocking=True); de = host_e.to(dev, non_blocking=True); doff = host_off.to(dev, non_blocking=True)
    out = {
        "vertices": torch.empty(B, V, 3, device=dev, dtype=torch.float32),
        "vertex_existence": torch.empty(B, V, device=dev, dtype=torch.float32),
        "edge_labels": torch.empty(B, max_e, device=dev, dtype=torch.float32),
        "vertex_counts": torch.empty(B, device=dev, dtype=torch.int64),
    }
    with torch.cuda.device(dev):
        call("wf_pack_targets", ops._p(dv), ops._p(doff[0]), ops._p(de), ops._p(doff[1]), B, V, max_e, E, ops._p(out["vertices"]),
             ops._p(out["vertex_existence"]), ops._p(out["vertex_counts"]), ops._p(out["edge_labels"]), ops._s())
    ops._count()
    done = torch.cuda.Event()
    done.record()
    for t in out.values():
        ops.mark_ready(t, done)          # lets WireframeLoss start its matching on a side stream (see _match_device)
    # the host already knows the counts it packed: PointCloudToWireframe.forward needs no device->host read for them
    # (valid while the tensor is not modified: the tag carries the version it was made at)
    out["vertex_counts"]._wf_host_counts = (out["vertex_counts"]._version, tuple(counts))
    return out


class DevicePrefetcher:
    """Iterates `(point_clouds, targets)` batches one step ahead of the consumer: the pinned host->device copies of the next
    batch run on a side stream while the current step computes.  `batches` yields dicts of CPU tensors (pinned or not;
    unpinned ones are staged through a pinned buffer) -- e.g. {'point_clouds': (B,N,8), 'vertices': ..., ...}."""

    def __init__(self, batches: Iterable[Dict[str, torch.Tensor]], device):
        self.it: Iterator = iter(batches)
        self.dev = torch.device(device)
        self.stream = torch.cuda.Stream(self.dev)
        self._next = None
        self._preload()

    def _preload(self) -> None:
        try:
            host = next(self.it)
        except StopIteration:
            self._next = None
            return
        with torch.cuda.stream(self.stream):
            dev_batch = {}
            for k, v in host.items():
                if torch.is_tensor(v):
                    src = v if v.is_pinned() else v.pin_memory()
                    dev_batch[k] = src.to(self.dev, non_blocking=True)
                    if k == "vertex_counts":          # host values travel with the device copy (no read-back in the step)
                        dev_batch[k]._wf_host_counts = (dev_batch[k]._version, tuple(int(c) for c in v.tolist()))
                else:
                    dev_batch[k] = v
            ev = torch.cuda.Event()
            ev.record(self.stream)
            for v in dev_batch.values():
                if torch.is_tensor(v):
                    ops.mark_ready(v, ev)
        self._next = (dev_batch, ev)

    def __iter__(self):
        return self

    def __next__(self) -> Dict[str, torch.Tensor]:
        if self._next is None:
            raise StopIteration
        batch, ev = self._next
        torch.cuda.current_stream(self.dev).wait_event(ev)
        for v in batch.values():
            if torch.is_tensor(v):
                v.record_stream(torch.cuda.current_stream(self.dev))      # allocated on the side stream, used on this one
        self._preload()
        return batch
