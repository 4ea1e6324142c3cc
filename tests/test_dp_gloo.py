"""Data-parallel host logic on CPU (gloo, world_size 2): the bucketed gradient all-reduce and the batch-global
loss normalisation.  The loss here is the oracle's (CPU); the product's reducer and
shard_loss_weights are the code under test."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


class TinyHead(torch.nn.Module):
    """x (B, F) -> predictions dict with the same fields/semantics the real model emits."""

    def __init__(self, F, V, max_e):
        super().__init__()
        self.v = torch.nn.Linear(F, V * 3)
        self.e = torch.nn.Linear(F, V)
        self.g = torch.nn.Linear(F, max_e)
        self.unused = torch.nn.Linear(3, 3)          # never used -> grad None (like EdgePredictor.spatial_proj)
        self.V = V

    def forward(self, x, counts):
        B = x.shape[0]
        verts = self.v(x).reshape(B, self.V, 3)
        ex = torch.sigmoid(self.e(x))
        ne = [int(c) * (int(c) - 1) // 2 for c in counts]
        me = max(ne)
        ep = torch.sigmoid(self.g(x))[:, :me]
        mask = torch.zeros(B, me)
        for b, n in enumerate(ne):
            mask[b, :n] = 1.0
        return {"vertices": verts, "existence_probabilities": ex, "edge_probs": ep * mask}


def _data(V):
    rng = np.random.default_rng(0)
    B, F = 6, 10
    counts = [3, 7, 2, 5, 8, 4]
    x = torch.from_numpy(rng.normal(size=(B, F)).astype(np.float32))
    tv = torch.zeros(B, V, 3); te = torch.zeros(B, V)
    for b, c in enumerate(counts):
        tv[b, :c] = torch.from_numpy(rng.uniform(-1, 1, (c, 3)).astype(np.float32)); te[b, :c] = 1
    max_e = max(c * (c - 1) // 2 for c in counts)
    el = torch.from_numpy((rng.uniform(size=(B, max_e)) < 0.3).astype(np.float32))
    for b, c in enumerate(counts):
        el[b, c * (c - 1) // 2:] = 0
    return x, counts, tv, te, el, max_e


def _worker(rank, world, port, out):
    sys.path.insert(0, os.path.join(ROOT, "wireframe-3d-prediction_b200")); sys.path.insert(0, ROOT)
    from oracle import wireframe_oracle as wo
    from wf_b200.parallel import GradAllReduce, shard_loss_weights
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    V = 8
    x, counts, tv, te, el, max_e = _data(V)
    torch.manual_seed(0)
    model = TinyHead(x.shape[1], V, max_e)
    shards = [[0, 1, 2, 3], [4, 5]]                       # unequal shards on purpose
    idx = shards[rank]
    red = GradAllReduce(model, bucket_bytes=256)         # tiny buckets -> several all-reduces
    w = (3.0, 1.0, 1.5)
    res = []
    for step in range(3):                                 # step 0 builds the buckets, later steps use the hooks
        red.zero()
        lc = [counts[i] for i in idx]
        pred = model(x[idx], lc)
        me_local = max(c * (c - 1) // 2 for c in lc)
        tgt = {"vertices": tv[idx], "vertex_existence": te[idx], "edge_labels": el[idx][:, :me_local],
               "vertex_counts": torch.tensor(lc)}
        ld = wo.loss_forward(pred, tgt, *w)
        sv, sx, se = shard_loss_weights(lc, [[counts[i] for i in s] for s in shards], V)
        loss = w[0] * sv * ld["vertex_loss"] + w[2] * sx * ld["existence_loss"] + w[1] * se * ld["edge_loss"]
        loss.backward()
        red.finish()
        res.append({k: (None if p.grad is None else p.grad.detach().clone()) for k, p in model.named_parameters()})
    if rank == 0:
        # single-process full-batch reference
        torch.manual_seed(0)
        ref = TinyHead(x.shape[1], V, max_e)
        pred = ref(x, counts)
        tgt = {"vertices": tv, "vertex_existence": te, "edge_labels": el, "vertex_counts": torch.tensor(counts)}
        wo.loss_forward(pred, tgt, *w)["total_loss"].backward()
        ok = True
        for r in res:
            for k, p in ref.named_parameters():
                if p.grad is None:
                    ok &= r[k] is None
                else:
                    ok &= bool(torch.allclose(r[k], p.grad, rtol=1e-4, atol=1e-6))
        out.put(ok)
    g0 = torch.cat([g.reshape(-1) for g in res[-1].values() if g is not None])
    gathered = [torch.zeros_like(g0) for _ in range(world)]
    dist.all_gather(gathered, g0)
    if rank == 0:
        out.put(bool(torch.equal(gathered[0], gathered[1])))
    dist.destroy_process_group()


def test_sharded_gradients_equal_full_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True, "sum of rescaled shard gradients != full-batch gradient"
    assert q.get(timeout=10) is True, "ranks disagree after all-reduce"


def _early_worker(rank, world, port, out):
    """The per-layer hand-off (ops.GRAD_READY_HOOK, what EncoderPointMLP_TC.backward calls for every finished layer): gradients
    reduced INSIDE a Function's backward must arrive once (not again when their bucket fires), only between zero() and
    finish(), and autograd must adopt the handed-off tensors as .grad."""
    sys.path.insert(0, os.path.join(ROOT, "wireframe-3d-prediction_b200")); sys.path.insert(0, ROOT)
    from wf_b200 import ops
    from wf_b200.parallel import GradAllReduce
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    calls = []

    class EarlyLinear(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, W, b):
            ctx.save_for_backward(x, W)
            return x @ W.t() + b

        @staticmethod
        def backward(ctx, dy):
            x, W = ctx.saved_tensors
            dW, db, dx = dy.t() @ x, dy.sum(0), dy @ W
            hook = ops.GRAD_READY_HOOK
            if hook is not None:
                calls.append(1)
                hook([dW, db])                              # final here: reduced while the rest of the backward still runs
            return dx, dW, db

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a = torch.nn.Linear(6, 16); self.b = torch.nn.Linear(16, 16); self.c = torch.nn.Linear(16, 3)

        def forward(self, x):
            h = torch.relu(self.a(x))
            h = torch.relu(EarlyLinear.apply(h, self.b.weight, self.b.bias))
            return self.c(h)

    torch.manual_seed(0)
    net = Net()
    g = torch.Generator().manual_seed(5)
    X = torch.randn(8, 6, generator=g); Y = torch.randn(8, 3, generator=g)
    sl = slice(rank * 4, rank * 4 + 4)
    red = GradAllReduce(net, bucket_bytes=128)
    ok = True
    for step in range(3):                                   # step 0 builds the buckets (the hand-off is inactive until then)
        red.zero()
        ((net(X[sl]) - Y[sl]) ** 2).sum().backward()
        red.finish()
        grads = {k: p.grad.clone() for k, p in net.named_parameters()}
    ok &= len(calls) == 3
    torch.manual_seed(0)
    ref = Net()
    ((ref(X) - Y) ** 2).sum().backward()
    for k, p in ref.named_parameters():
        ok &= bool(torch.allclose(grads[k], p.grad, rtol=1e-5, atol=1e-6))
    # outside zero()/finish() the hook must not start a collective (another model's backward may run between steps)
    n_before = len(red._handles)
    probe = torch.ones(4) * (rank + 1)
    ops.GRAD_READY_HOOK([probe])
    ok &= len(red._handles) == n_before and bool(torch.equal(probe, torch.ones(4) * (rank + 1)))
    red.remove()
    ok &= ops.GRAD_READY_HOOK is None
    if rank == 0:
        out.put(bool(ok))
    dist.destroy_process_group()


def test_early_gradient_handoff_is_reduced_exactly_once():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_early_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True


def test_shard_loss_weights_sum_rules():
    sys.path.insert(0, os.path.join(ROOT, "wireframe-3d-prediction_b200"))
    from wf_b200.parallel import shard_loss_weights
    allc = [[3, 7, 2, 5], [8, 4]]
    ws = [shard_loss_weights(c, allc, 8) for c in allc]
    assert abs(sum(w[0] for w in ws) - 1.0) < 1e-12 and abs(sum(w[1] for w in ws) - 1.0) < 1e-12
    # edge: (B_r * maxE_r) / (B * maxE)
    assert abs(ws[0][2] - (4 * 21) / (6 * 28)) < 1e-12 and abs(ws[1][2] - (2 * 28) / (6 * 28)) < 1e-12


def _pool_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from wf_b200.parallel import reduce_pool_shards
    g = torch.Generator().manual_seed(100 + rank)
    B, C, K, n = 3, 16, 8, 50
    vals = torch.randn(B, n, C, generator=g).clamp(max=2.5)
    vals[:, 7] = 3.0                                        # a tie across ranks on the maximum: smallest GLOBAL index must win
    mask = torch.rand(B, n, generator=g) > 0.3

    def pack(v, idx):                                       # the kernels' packing (gemm_tc.cu pool_push), in numpy
        import numpy as np
        u = v.numpy().astype(np.float32).view(np.uint32).astype(np.uint64)
        o = np.where(u & 0x80000000, (~u) & 0xFFFFFFFF, u | 0x80000000)
        return torch.from_numpy(((o << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - idx.numpy().astype(np.uint64))).astype(np.int64))
    gidx = (torch.arange(n) + rank * n).view(1, n, 1).expand(B, n, C)
    pk = pack(vals, gidx)
    flip = -(1 << 63)                                       # unsigned order for torch's signed int64 max
    pu = (pk ^ flip).max(dim=1).values ^ flip
    pm = torch.where(mask.unsqueeze(-1), pk ^ flip, torch.full_like(pk, flip)).max(dim=1).values ^ flip
    packed = torch.stack([pu, pm])
    hsum = torch.stack([vals[..., :K].sum(1), (vals[..., :K] * mask.unsqueeze(-1)).sum(1)])
    cnt = mask.sum(1).float()
    reduce_pool_shards(packed, hsum, cnt)
    # by value (numpy): torch tensors travel through a Queue as shared-memory handles that die with this process
    q.put((rank, packed.numpy(), hsum.numpy(), cnt.numpy(), vals.numpy(), mask.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_point_sharded_pool_reduce_gloo():
    """reduce_pool_shards over 2 gloo ranks: integer MAX of the packed (value, ~global index) words = global max with the
    smallest global index on ties; sums and counts add."""
    import numpy as np
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_pool_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted([q.get(timeout=120) for _ in ps], key=lambda t: t[0])
    res = [(r[0], *[torch.from_numpy(a) for a in r[1:]]) for r in res]
    [p.join(timeout=60) for p in ps]
    allv = torch.cat([res[0][4], res[1][4]], dim=1); allm = torch.cat([res[0][5], res[1][5]], dim=1)
    for r in range(2):
        packed, hsum, cnt = res[r][1], res[r][2], res[r][3]
        idx = (0xFFFFFFFF - (packed[0].numpy().astype(np.uint64) & np.uint64(0xFFFFFFFF))).astype(np.int64)
        ref_max = allv.max(dim=1).values
        first = torch.where(allv == ref_max.unsqueeze(1), torch.arange(allv.shape[1]).view(1, -1, 1), 10 ** 6).min(dim=1).values
        assert np.array_equal(idx, first.numpy())
        assert (first == 7).all()                           # the planted tie: rank 0's point 7 wins over rank 1's (global 57)
        assert torch.allclose(hsum[0], allv[..., :8].sum(1), atol=1e-5)
        assert torch.allclose(hsum[1], (allv[..., :8] * allm.unsqueeze(-1)).sum(1), atol=1e-5)
        assert torch.equal(cnt, allm.sum(1).float())
