"""WireframeLoss -- drop-in for the reference's losses/WireframeLoss.py (same ctor, attributes,
forward(predictions, targets) -> {total_loss, vertex_loss, existence_loss, edge_loss}).

The cost matrices (L1 cdist + |e-1|, constant dummy columns) and the Jonker-Volgenant assignment run
on the device, one warp per sample (wf_loss_match), instead of B device->host syncs + scipy calls
(losses/WireframeLoss.py:129-244); the three loss terms and their backward are one fused kernel each."""
import numpy as np
import torch
import torch.nn as nn

from wf_b200 import ops
from wf_b200._lib import LSAP_INFEASIBLE, LSAP_INVALID


class WireframeLoss(nn.Module):
    def __init__(self, vertex_weight=1.0, edge_weight=1.0, existence_weight=1.0):
        super(WireframeLoss, self).__init__()
        self.vertex_weight = vertex_weight
        self.edge_weight = edge_weight
        self.existence_weight = existence_weight
        self.smooth_l1_loss = nn.SmoothL1Loss()      # kept for attribute parity; the kernels implement beta=1
        self.bce_loss = nn.BCELoss()
        # The reference gets scipy's ValueError (infeasible / invalid cost matrix, i.e. NaN or inf predictions) inside
        # forward, because it synchronises with the device B times per call anyway.  Here the solver's status words are
        #   True       read back immediately (one device->host sync per call; what _hungarian_matching() always does),
        #   "deferred" copied to pinned memory asynchronously and checked at the start of the NEXT call or by
        #              check_pending() -- same exception, one step later, no sync inside the step (default),
        #   False      never read.
        self.check_status = "deferred"
        self._pending = None
        # The matching needs the vertex head's outputs only, not the edge head's.  When the model tagged them with a
        # ready-event (PointCloudToWireframe.forward) and the targets are known to be complete -- tagged the same way
        # by wf_b200.targets, or the very tensors already used by the previous call, as in train.py's loop -- the
        # assignment runs on a side stream beside the edge head instead of after it.  Otherwise: main stream.
        self.overlap_matching = True
        self._seen_targets = None

    @staticmethod
    def _raise_for(st):
        if any(s == LSAP_INFEASIBLE for s in st):
            raise ValueError("cost matrix is infeasible")              # scipy's message (reference Q9)
        if any(s == LSAP_INVALID for s in st):
            raise ValueError("matrix contains invalid numeric entries")

    def check_pending(self):
        """Raise the ValueError of the last forward() if its assignment problems were infeasible/invalid."""
        if self._pending is not None:
            buf, ev = self._pending
            self._pending = None
            ev.synchronize()
            self._raise_for(buf.tolist())

    def _match_device(self, predictions, targets, sync=None):
        mode = self.check_status if sync is None else sync
        if mode == "deferred":
            self.check_pending()
        waits = self._overlap_events(predictions, targets) if self.overlap_matching else None
        if waits is None:
            return self._launch_match(predictions, targets, mode)
        main = torch.cuda.current_stream()
        side = ops.side_stream(predictions['vertices'].device)
        for ev in waits:
            side.wait_event(ev)
        with torch.cuda.stream(side):
            col = self._launch_match(predictions, targets, mode)
            done = torch.cuda.Event()
            done.record(side)
        main.wait_event(done)
        col.record_stream(main)
        return col

    def _overlap_events(self, predictions, targets):
        """Events the side stream must wait for, or None when the targets' completion cannot be established."""
        ready = getattr(predictions['vertices'], '_wf_ready', None)
        if ready is None:
            return None
        waits = [ready]
        sig = []
        for k in ('vertices', 'vertex_counts'):
            t = targets[k]
            ev = getattr(t, '_wf_ready', None)
            if ev is not None:
                waits.append(ev)
            sig.append((t, t._version, ev is not None))
        # "already used by the previous call" means THE SAME TENSOR OBJECTS at the same version.  The previous call's
        # tensors are held here, so a new batch cannot be mistaken for them through a recycled address (fresh tensors
        # copied on the main stream after the model forward would otherwise be read by the side stream mid-copy).
        seen, self._seen_targets = self._seen_targets, sig
        if all(s[2] for s in sig):
            return waits
        if seen is not None and all(a[0] is b[0] and a[1] == b[1] for a, b in zip(seen, sig)):
            return waits
        return None

    def _launch_match(self, predictions, targets, mode):
        col, status, _ = ops.loss_match(predictions['vertices'], predictions['existence_probabilities'],
                                        targets['vertices'], targets['vertex_counts'])
        if mode is True:
            self._raise_for(status.tolist())
        elif mode == "deferred":
            buf = torch.empty(status.shape, dtype=status.dtype, pin_memory=True)
            buf.copy_(status, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            self._pending = (buf, ev)
        return col

    def _hungarian_matching(self, predictions, targets):
        """Reference API (losses/WireframeLoss.py:106): list of (pred_indices, target_indices) numpy int64
        arrays per sample, dummy-column assignments filtered out."""
        col = self._match_device(predictions, targets, sync=True).cpu().numpy()
        counts = targets['vertex_counts'].cpu().numpy()
        out = []
        for b in range(col.shape[0]):
            rows = np.arange(col.shape[1], dtype=np.int64)
            cols = col[b].astype(np.int64)
            keep = (cols >= 0) & (cols < counts[b])
            out.append((rows[keep], cols[keep]))
        return out

    def forward(self, predictions, targets):
        col = self._match_device(predictions, targets)
        pe = predictions['edge_probs']
        te = targets['edge_labels']
        total, vloss, xloss, eloss = ops.WireframeLossFn.apply(
            predictions['vertices'], predictions['existence_probabilities'], pe, targets['vertices'],
            targets['vertex_existence'], te, col, targets['vertex_counts'],
            float(self.vertex_weight), float(self.edge_weight), float(self.existence_weight))
        return {'total_loss': total, 'vertex_loss': vloss, 'existence_loss': xloss, 'edge_loss': eloss}
