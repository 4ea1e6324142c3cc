"""Data-parallel training for the drop-in modules (net-new: the reference is single-device).

One process per GPU (torch.distributed, NCCL over NVLink).  The batch of point clouds is sharded across
ranks; nothing in the model or the loss couples samples (LayerNorm is per row, pooling per cloud, matching
per sample -- SURVEY 8e), so the only exchange is the gradient all-reduce:

  * parameters are cut into ~32 MB buckets in the order their gradients become ready (edge head -> vertex head ->
    fusion -> encoder MLP); the gradient tensors themselves are reduced in place, grouped per bucket into one NCCL
    launch (no flat staging buffer, no extra pass over the gradients);
  * as soon as a bucket's last gradient exists, its all-reduce is issued on a side stream, so the transfers (124 MB
    total) hide under the encoder backward, which is ~97 % of the backward time;
  * parameters that never receive a gradient (EdgePredictor.spatial_proj) keep `grad = None`, so the
    optimizer skips them exactly as in the reference.

Loss normalisation: the vertex term is normalised by the batch-total match count and the edge
term by B * max_edges, both batch-global.  `shard_loss_weights` rescales each rank's local terms so that the SUM
all-reduce of the gradients equals the full-batch gradient of the single-GPU run."""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_loss_weights(local_counts: Sequence[int], all_counts: Sequence[Sequence[int]], max_vertices: int):
    """Per-rank multipliers (vertex, existence, edge) for the three loss terms.

    local term * multiplier summed over ranks == the full-batch term of losses/WireframeLoss.py:82-96,276-281.
    Counts are ground-truth vertex counts (host-known before the step), so no collective is needed."""
    cl = [min(int(c), max_vertices) for c in local_counts]
    flat = [min(int(c), max_vertices) for r in all_counts for c in r]
    n_local, n_all = sum(cl), sum(flat)
    b_local, b_all = len(cl), len(flat)
    me_local = max((c * (c - 1) // 2 for c in cl), default=0)
    me_all = max((c * (c - 1) // 2 for c in flat), default=0)
    wv = n_local / n_all if n_all else 0.0
    wx = b_local / b_all if b_all else 0.0
    we = (b_local * me_local) / (b_all * me_all) if me_all and b_all else 0.0
    return wv, wx, we


_INT64_MIN = -(1 << 63)


def reduce_pool_shards(packed: torch.Tensor, hsum: torch.Tensor, count: torch.Tensor, group=None) -> None:
    """Combine per-rank partial pools of clouds that are sharded by POINTS (batch < world size).

    packed : int64 [2, B, C]  order-preserving (value, ~point index) words of the max pools -> integer MAX all-reduce gives
             the global max and, on ties, the smallest GLOBAL point index (ranks pass their first point as index_offset);
    hsum   : fp32 [2, B, K]   column sums of the last hidden layer (all points / valid points) -> SUM;
    count  : fp32 [B]         valid points of this shard -> SUM.
    In place; a no-op without an initialised process group."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    # the words are UNSIGNED 64-bit; torch's int64 MAX is signed -> flip the top bit around the reduction
    packed.bitwise_xor_(_INT64_MIN)
    dist.all_reduce(packed, op=dist.ReduceOp.MAX, group=group)
    packed.bitwise_xor_(_INT64_MIN)
    dist.all_reduce(hsum, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(count, op=dist.ReduceOp.SUM, group=group)


def encode_point_sharded(encoder, x_local: torch.Tensor, rank: int, world: int, group=None, chunk_rows=None):
    """Encoder pools for clouds whose points are split across ranks: rank r holds points [r*n, (r+1)*n) of every cloud
    (x_local: [B, n, 8]).  Every rank returns the pools of the FULL clouds (max_m, avg_m, max_u, mean_u, arg_m, arg_u);
    argmax indices are global point indices.  One exchange: reduce_pool_shards (3 small all-reduces)."""
    from . import ops
    n = x_local.shape[1]
    x_local, p = encoder.tc_inputs(x_local)
    with torch.no_grad():
        return ops.encoder_pooled_infer(x_local, p, chunk_rows=chunk_rows, index_offset=rank * n, points_total=world * n,
                                        reduce_fn=lambda a, b, c: reduce_pool_shards(a, b, c, group))


class GradAllReduce:
    """Bucketed, backward-overlapped gradient all-reduce (SUM) for a module replicated on every rank.

    Zero-copy: every step starts from `grad = None` (train.py:138 zero_grad), autograd adopts the gradient tensors the
    kernels return, and each bucket is reduced IN PLACE as one grouped NCCL launch over its tensors (ncclGroupStart/End
    through torch's coalescing manager) -- no flat staging buffer, no copy or accumulate pass over the 124 MB."""

    def __init__(self, module: torch.nn.Module, bucket_bytes: int = 32 << 20, process_group=None):
        self.module = module
        self.group = process_group
        self.bucket_bytes = bucket_bytes
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.buckets: List[dict] = []
        self._order: List[torch.nn.Parameter] = []
        self._hooks = []
        self._handles = []
        self._stream = None
        self._built = False
        self._bucket_of: Dict[torch.nn.Parameter, dict] = {}
        self._early_ptrs = set()
        self._armed = False
        if self.world > 1:
            from . import ops
            ops.GRAD_READY_HOOK = self._early
        self._hooked = set()
        for p in module.parameters():
            if p.requires_grad:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
                self._hooked.add(id(p))

    # ---- step protocol: zero() -> forward/backward -> finish() -> optimizer.step()
    def zero(self) -> None:
        for p in self.module.parameters():
            p.grad = None
        if self.world == 1:
            return
        if self._built:
            for b in self.buckets:
                b["pending"] = b["n"]
        else:
            self._order = []
        self._handles = []
        self._early_ptrs = set()
        self._armed = True

    def _early(self, tensors: List[torch.Tensor]) -> None:
        """ops.GRAD_READY_HOOK: gradient tensors that are final INSIDE the encoder's backward (one layer at a time) are
        reduced at once, under the GEMMs of the remaining layers, instead of in the bucket that only fills when the whole
        Function has returned (its 21 MB were the exposed tail of the step).  They are skipped when their bucket fires."""
        if self.world == 1 or not self._built or not tensors or not self._armed:
            return                       # only between zero() and finish(): another model's backward must not start a collective
        # aliases, not the tensors themselves: the NCCL work objects keep their operands alive, and autograd adopts a returned
        # gradient as .grad without a copy only while nobody else references that tensor object
        self._launch([t.detach() for t in tensors])
        self._early_ptrs.update(t.data_ptr() for t in tensors)

    def _not_early(self, grads: List[torch.Tensor]) -> List[torch.Tensor]:
        return [g for g in grads if g.data_ptr() not in self._early_ptrs]

    def _on_grad(self, p: torch.nn.Parameter) -> None:
        if self.world == 1:
            return
        if not self._built:
            self._order.append(p)
            return
        b = self._bucket_of.get(p)
        if b is None:
            return
        b["pending"] -= 1
        if b["pending"] < 0:
            raise RuntimeError("GradAllReduce: a parameter received a second gradient before finish(); the bucket it belongs to "
                               "has already been handed to NCCL (one backward per zero()/finish() pair)")
        if b["pending"] == 0:
            self._launch(self._not_early([q.grad for q in b["params"] if q.grad is not None]))

    def _reduce_list(self, tensors: List[torch.Tensor]) -> None:
        """One grouped in-place SUM all-reduce over `tensors` (NCCL); per-tensor calls on backends without grouping."""
        if not tensors:
            return
        if tensors[0].is_cuda and hasattr(dist, "_coalescing_manager"):
            with dist._coalescing_manager(group=self.group, device=tensors[0].device, async_ops=True) as cm:
                for t in tensors:
                    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            self._handles.append(cm)
        else:
            for t in tensors:
                self._handles.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def _launch(self, tensors: List[torch.Tensor]) -> None:
        if self.world == 1 or not tensors:
            return
        if tensors[0].is_cuda:
            if self._stream is None:
                self._stream = torch.cuda.Stream()
            self._stream.wait_stream(torch.cuda.current_stream())       # the gradients were produced on the compute stream
            with torch.cuda.stream(self._stream):
                self._reduce_list(tensors)
        else:
            self._reduce_list(tensors)

    def _check_registered(self) -> None:
        """Every parameter that received a gradient must have been seen at construction: a parameter created later (the
        lazy vertex_predictor.point_pool_proj when the reducer is built before the first forward) would never
        be reduced and the replicas would drift apart silently."""
        hooked = self._hooked
        late = [n for n, p in self.module.named_parameters() if p.requires_grad and p.grad is not None and id(p) not in hooked]
        if late:
            raise RuntimeError("GradAllReduce: parameters created after the reducer was built received gradients and would "
                               f"not be all-reduced: {late}.  Run one forward (materialising lazy layers, with identical "
                               "weights on every rank) before constructing GradAllReduce.")

    def finish(self) -> None:
        """Call after backward() -- ONE backward per zero()/finish() pair.  First step: fixes the buckets from the observed
        gradient order and reduces everything at once; later steps: waits for the in-flight bucket reductions."""
        self._armed = False
        if self.world == 1:
            return
        self._check_registered()
        if self._early_ptrs:
            # the early hand-off reduces the tensors a Function is about to return; that is only right if autograd then ADOPTS
            # them as .grad (no copy).  A copied gradient would have been cloned while its source was being reduced.
            held = {q.grad.data_ptr() for q in self.module.parameters() if q.grad is not None}
            if not self._early_ptrs <= held:
                raise RuntimeError("GradAllReduce: a gradient handed off inside backward was copied by autograd instead of being "
                                   "adopted as .grad; the early all-reduce cannot be used with this graph (set "
                                   "ops.GRAD_READY_HOOK = None)")
        if not self._built:
            self._build()
            self._launch([p.grad for b in self.buckets for p in b["params"]])
        else:
            for b in self.buckets:                  # a bucket whose hooks did not all fire (a parameter without gradient)
                if 0 < b["pending"] < b["n"] or (b["pending"] == b["n"] and any(p.grad is not None for p in b["params"])):
                    self._launch(self._not_early([q.grad for q in b["params"] if q.grad is not None])); b["pending"] = 0
        for h in self._handles:
            h.wait()
        if self._stream is not None:
            torch.cuda.current_stream().wait_stream(self._stream)
        self._handles = []

    def _build(self) -> None:
        seen = set()
        params = [p for p in self._order if p.grad is not None and not (id(p) in seen or seen.add(id(p)))]
        cur = {"n": 0, "params": [], "bytes": 0}
        for p in params:
            cur["params"].append(p); cur["n"] += 1; cur["bytes"] += p.numel() * 4
            if cur["bytes"] >= self.bucket_bytes:
                self.buckets.append(cur)
                cur = {"n": 0, "params": [], "bytes": 0}
        if cur["n"]:
            self.buckets.append(cur)
        for b in self.buckets:
            b["pending"] = b["n"]
            for p in b["params"]:
                self._bucket_of[p] = b
        self._built = True

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []
        from . import ops
        if ops.GRAD_READY_HOOK == self._early:
            ops.GRAD_READY_HOOK = None
