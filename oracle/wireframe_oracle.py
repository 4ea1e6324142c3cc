"""
oracle/wireframe_oracle.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement (torch CPU tensors, fp32 or fp64, autograd for the backward) of the hot path
BASELINE.json names: PointCloudToWireframe forward/backward + WireframeLoss + the matchers.
It is the checker for the CUDA path; nothing under wireframe-3d-prediction_b200/ may import it.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.

Every function cites the reference lines it restates (paths relative to /root/reference).
The restatement is *functional*: weights come in as a flat dict with the reference's
state_dict key names, so the same dict drives the reference model (in
tests/golden/make_golden.py, run in the authoring container), this oracle, and the CUDA path.

Pinning: tests/test_oracle_golden.py checks this file against tests/golden/*.npz, which were
produced by importing the unmodified reference from /root/reference (script committed beside
them).  The LSAP step is pinned separately against the installed scipy (tests/test_lsap_oracle.py).
Dropout: the reference's edge head has live Dropout(0.1) (models/EdgePredictor.py:37,44,60,64).
The oracle, the goldens and the parity tests all run with those four sites disabled.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))

# --------------------------------------------------------------------------------------
# parameter inventory (registration order of the reference modules)
# --------------------------------------------------------------------------------------
ENC_DIMS = [8, 512, 1024, 2048, 1024, 512]          # models/PointNetEncoder.py:20,35-45
FUSION_DIMS = [1024, 2048, 1024, 512]               # models/PointNetEncoder.py:57-65
HIDDEN = 512                                        # models/EdgePredictor.py:19
HEADS = 8


def param_shapes(max_vertices: int = 64, input_dim: int = 8) -> "list[tuple[str, tuple, str]]":
    """(key, shape, kind) for all 80 state_dict entries. kind in {w,b,lnw,lnb}."""
    out = []

    def lin(name, o, i):
        out.append((name + ".weight", (o, i), "w"))
        out.append((name + ".bias", (o,), "b"))

    def ln(name, d):
        out.append((name + ".weight", (d,), "lnw"))
        out.append((name + ".bias", (d,), "lnb"))

    dims = [input_dim] + ENC_DIMS[1:]
    for li in range(4):                              # mlp.{0,4,8,12} Linear, {1,5,9,13} LN
        lin(f"encoder.mlp.{4 * li}", dims[li + 1], dims[li])
        ln(f"encoder.mlp.{4 * li + 1}", dims[li + 1])
    lin("encoder.mlp.16", dims[5], dims[4])
    lin("encoder.feature_fusion.0", 2048, 1024); ln("encoder.feature_fusion.1", 2048)
    lin("encoder.feature_fusion.3", 1024, 2048); ln("encoder.feature_fusion.4", 1024)
    lin("encoder.feature_fusion.6", 512, 1024)
    for k, (o, i) in enumerate([(4096, 512), (2048, 4096), (2048, 2048), (1024, 2048)], start=1):
        lin(f"vertex_predictor.vertex_mlp{k}.0", o, i)
        ln(f"vertex_predictor.vertex_mlp{k}.1", o)
    lin("vertex_predictor.final_layer", max_vertices * 4, 1024)
    lin("vertex_predictor.residual_proj1", 2048, 512)
    lin("vertex_predictor.residual_proj2", 1024, 512)
    lin("vertex_predictor.point_pool_proj", 512, 1024)   # lazy in the reference (Q1)
    lin("edge_predictor.vertex_proj.0", 256, 3); ln("edge_predictor.vertex_proj.1", 256)
    lin("edge_predictor.vertex_proj.3", 512, 256); ln("edge_predictor.vertex_proj.4", 512)
    out.append(("edge_predictor.attention.in_proj_weight", (1536, 512), "w"))
    out.append(("edge_predictor.attention.in_proj_bias", (1536,), "b"))
    lin("edge_predictor.attention.out_proj", 512, 512)
    lin("edge_predictor.spatial_proj.0", 128, 3)          # unused in forward (Q3)
    lin("edge_predictor.spatial_proj.2", 128, 128)
    lin("edge_predictor.edge_mlp.0", 512, 1031); ln("edge_predictor.edge_mlp.1", 512)
    lin("edge_predictor.edge_mlp.4", 256, 512); ln("edge_predictor.edge_mlp.5", 256)
    lin("edge_predictor.edge_mlp.8", 128, 256)
    lin("edge_predictor.edge_mlp.10", 1, 128)
    return out


def make_state_dict(seed: int = 0, max_vertices: int = 64, dtype=torch.float32,
                    input_dim: int = 8) -> Dict[str, torch.Tensor]:
    """Deterministic weights that do not depend on torch's RNG stream (numpy PCG64):
    Linear ~ U(+-1/sqrt(fan_in)), LayerNorm gain 1+-0.1, LayerNorm shift +-0.1.
    The first encoder layer's intensity column is scaled so un-normalised intensity
    (~5e4) does not swamp the LayerNorm input."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sd = {}
    for key, shape, kind in param_shapes(max_vertices, input_dim):
        if kind == "w":
            bound = 1.0 / math.sqrt(shape[1])
            a = rng.uniform(-bound, bound, size=shape)
        elif kind == "b":
            a = rng.uniform(-0.05, 0.05, size=shape)
        elif kind == "lnw":
            a = 1.0 + rng.uniform(-0.1, 0.1, size=shape)
        else:
            a = rng.uniform(-0.1, 0.1, size=shape)
        sd[key] = torch.from_numpy(a.astype(np.float32)).to(dtype).contiguous()
    return sd


# --------------------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------------------
def make_inputs(seed: int, B: int, N: int, V: int, *, pad_frac: float = 0.0,
                norm_intensity: bool = False, min_count: int = 2,
                max_count: Optional[int] = None, dtype=torch.float32):
    """Point clouds shaped like datasets/building3d.py:109-126 output plus GT targets shaped
    like train.py:48-88,112-115.  Returns (points[B,N,8], targets dict, counts[B] int64)."""
    rng = np.random.Generator(np.random.PCG64(1234 + seed))
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
    max_count = V if max_count is None else max_count
    counts = rng.integers(min_count, max_count + 1, size=(B,))
    tv = np.zeros((B, V, 3))
    te = np.zeros((B, V))
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
    targets = {"vertices": t(tv), "vertex_existence": t(te), "edge_labels": t(el),
               "vertex_counts": counts_t}
    return t(pts), targets, counts_t


# --------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------
def _lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def _ln(sd, name, x):
    w = sd[name + ".weight"]
    return F.layer_norm(x, (w.shape[0],), w, sd[name + ".bias"], 1e-5)


def _q(t: torch.Tensor) -> torch.Tensor:
    """Round to bf16 (nearest even) in the forward value, identity in the backward (straight-through)."""
    return t + (t.to(torch.bfloat16).to(t.dtype) - t).detach()


def encoder_point_features(sd, x: torch.Tensor, emulate_bf16: bool = False, return_hidden: bool = False):
    """models/PointNetEncoder.py:30-48,90-98 -- the per-point MLP (Linear+LN+ReLU x4, Linear).

    emulate_bf16: the SAME function with the roundings of the product's bf16 mode inserted where its kernels round
    (DESIGN.md 2.3): post-ReLU activations h1..h4 and the weights of layers 2-5 are rounded to bf16 before each product,
    the pre-LayerNorm z2..z4 are rounded to bf16 AFTER their row statistics were taken from the unrounded fp32 values,
    accumulation stays fp32.  Gradients are straight-through.  With it the oracle predicts the bf16 mode's outputs to
    ~1e-5 instead of ~3e-3, so ReLU masks, argmax and matchings coincide and gradients can be compared tightly."""
    B, N, D = x.shape
    h = x.reshape(B * N, D)
    if not emulate_bf16:
        for li in range(4):
            h = torch.relu(_ln(sd, f"encoder.mlp.{4 * li + 1}", _lin(sd, f"encoder.mlp.{4 * li}", h)))
        pf = _lin(sd, "encoder.mlp.16", h).reshape(B, N, -1)
        return (pf, h) if return_hidden else pf
    h = _q(torch.relu(_ln(sd, "encoder.mlp.1", _lin(sd, "encoder.mlp.0", h))))          # layer 1: fp32 math, bf16 store
    for li in range(1, 4):
        z = F.linear(h, _q(sd[f"encoder.mlp.{4 * li}.weight"]), sd[f"encoder.mlp.{4 * li}.bias"])
        mean = z.mean(dim=1, keepdim=True)
        rstd = torch.rsqrt(z.var(dim=1, unbiased=False, keepdim=True) + 1e-5)
        y = (_q(z) - mean) * rstd * sd[f"encoder.mlp.{4 * li + 1}.weight"] + sd[f"encoder.mlp.{4 * li + 1}.bias"]
        h = _q(torch.relu(y))
    pf = F.linear(h, _q(sd["encoder.mlp.16.weight"]), sd["encoder.mlp.16.bias"]).reshape(B, N, -1)
    return (pf, h) if return_hidden else pf


def encoder_pools(x: torch.Tensor, pf: torch.Tensor):
    """models/PointNetEncoder.py:85-86,103-111 -- mask, masked mean, masked max (+argmax)."""
    mask = x.detach().abs().sum(dim=-1) > 1e-9
    cnt = mask.sum(dim=1, keepdim=True).clamp(min=1).to(pf.dtype)
    avg = (pf * mask.unsqueeze(-1)).sum(dim=1) / cnt
    mx, arg = pf.masked_fill(~mask.unsqueeze(-1), float("-inf")).max(dim=1)
    mx = torch.where(torch.isfinite(mx), mx, torch.zeros_like(mx))
    return mask, avg, mx, arg


def encoder_forward(sd, x, emulate_bf16: bool = False):
    """models/PointNetEncoder.py:67-119 -> (global_features[B,512], point_features[B,N,512]).
    emulate_bf16: see encoder_point_features; the mean pool is taken through the affine map like the product does
    (mean of the bf16 h4, then the fp32 master weight of the final Linear -- DESIGN.md 2.2)."""
    if not emulate_bf16:
        pf = encoder_point_features(sd, x)
        _, avg, mx, _ = encoder_pools(x, pf)
    else:
        pf, h4 = encoder_point_features(sd, x, True, return_hidden=True)
        mask, _, mx, _ = encoder_pools(x, pf)
        B, N = x.shape[:2]
        cnt = mask.sum(dim=1, keepdim=True).clamp(min=1).to(pf.dtype)
        hbar = (h4.reshape(B, N, -1) * mask.unsqueeze(-1)).sum(dim=1) / cnt
        avg = _lin(sd, "encoder.mlp.16", hbar)
    g = torch.cat([mx, avg], dim=1)
    g = torch.relu(_ln(sd, "encoder.feature_fusion.1", _lin(sd, "encoder.feature_fusion.0", g)))
    g = torch.relu(_ln(sd, "encoder.feature_fusion.4", _lin(sd, "encoder.feature_fusion.3", g)))
    gf = _lin(sd, "encoder.feature_fusion.6", g)
    if emulate_bf16:
        mean_u = _lin(sd, "encoder.mlp.16", h4.reshape(B, N, -1).mean(dim=1))
        return gf, pf, mean_u
    return gf, pf


def vertex_forward(sd, gfeat, pf, max_vertices: int, mean_override=None):
    """models/VertexPredictor.py:63-133 (unmasked mean/max pool, projection, 4 LN-MLP blocks with
    two residuals added AFTER LN+ReLU, final layer, sigmoid, >0.5 count)."""
    pooled = torch.cat([pf.mean(dim=1) if mean_override is None else mean_override, pf.max(dim=1).values], dim=1)
    eg = gfeat + _lin(sd, "vertex_predictor.point_pool_proj", pooled)
    blk = lambda k, t: torch.relu(_ln(sd, f"vertex_predictor.vertex_mlp{k}.1",
                                      _lin(sd, f"vertex_predictor.vertex_mlp{k}.0", t)))
    h = blk(2, blk(1, eg))
    h = blk(3, h) + _lin(sd, "vertex_predictor.residual_proj1", eg)
    h = blk(4, h) + _lin(sd, "vertex_predictor.residual_proj2", eg)
    vf = _lin(sd, "vertex_predictor.final_layer", h).reshape(gfeat.shape[0], max_vertices, 4)
    prob = torch.sigmoid(vf[:, :, 3])
    return vf[:, :, :3], prob, (prob > 0.5).sum(dim=1)


def pair_index(c: int) -> torch.Tensor:
    """models/EdgePredictor.py:83-89 -- all (i,j), i<j, row-major."""
    iu = torch.triu_indices(c, c, offset=1)
    return iu.t().contiguous()


def edge_forward(sd, verts: torch.Tensor):
    """models/EdgePredictor.py:91-140 for ONE sample: verts[c,3] -> (probs[E], pairs[E,2]).
    Dropout sites disabled.  c<=1 raises IndexError like the reference."""
    c = verts.shape[0]
    if c <= 1:
        raise IndexError("too many indices for tensor of dimension 1")
    f = F.gelu(_ln(sd, "edge_predictor.vertex_proj.1", _lin(sd, "edge_predictor.vertex_proj.0", verts)))
    f = _ln(sd, "edge_predictor.vertex_proj.4", _lin(sd, "edge_predictor.vertex_proj.3", f))
    # nn.MultiheadAttention(512, 8, batch_first=True), self-attention, no masks
    qkv = F.linear(f, sd["edge_predictor.attention.in_proj_weight"],
                   sd["edge_predictor.attention.in_proj_bias"])
    q, k, v = qkv.split(HIDDEN, dim=1)
    hd = HIDDEN // HEADS
    q = q.reshape(c, HEADS, hd).transpose(0, 1)
    k = k.reshape(c, HEADS, hd).transpose(0, 1)
    v = v.reshape(c, HEADS, hd).transpose(0, 1)
    att = torch.softmax((q / math.sqrt(hd)) @ k.transpose(1, 2), dim=-1)
    o = (att @ v).transpose(0, 1).reshape(c, HIDDEN)
    f = f + _lin(sd, "edge_predictor.attention.out_proj", o)
    pairs = pair_index(c)
    vi, vj = verts[pairs[:, 0]], verts[pairs[:, 1]]
    dist = torch.norm(vi - vj, dim=-1, keepdim=True)
    e = torch.cat([f[pairs[:, 0]], f[pairs[:, 1]], vi, vj, dist], dim=-1)
    e = F.gelu(_ln(sd, "edge_predictor.edge_mlp.1", _lin(sd, "edge_predictor.edge_mlp.0", e)))
    e = F.gelu(_ln(sd, "edge_predictor.edge_mlp.5", _lin(sd, "edge_predictor.edge_mlp.4", e)))
    e = F.gelu(_lin(sd, "edge_predictor.edge_mlp.8", e))
    return torch.sigmoid(_lin(sd, "edge_predictor.edge_mlp.10", e)).reshape(-1), pairs


def model_forward(sd, x, target_counts=None, *, training: bool, max_vertices: int, emulate_bf16: bool = False):
    """models/PointCloudToWireframe.py:43-121.  Edge head runs per sample on the PREFIX
    vertices[:count] (GT count when training, #(p>0.5) otherwise -- SURVEY Q5), results are
    zero-padded to the batch maximum edge count."""
    if emulate_bf16:
        gfeat, pf, mean_u = encoder_forward(sd, x, True)
        verts, prob, dyn = vertex_forward(sd, gfeat, pf, max_vertices, mean_override=mean_u)
    else:
        gfeat, pf = encoder_forward(sd, x)
        verts, prob, dyn = vertex_forward(sd, gfeat, pf, max_vertices)
    B = x.shape[0]
    use = target_counts if (training and target_counts is not None) else dyn
    probs, idx = [], []
    for b in range(B):
        c = int(use[b])
        p, pr = edge_forward(sd, verts[b, :c])
        probs.append(p); idx.append(pr.tolist())
    max_e = max((len(p) for p in probs), default=0)
    padded = torch.zeros(B, max_e, dtype=verts.dtype)
    if max_e > 0:
        padded = torch.stack([F.pad(p, (0, max_e - p.shape[0])) for p in probs])
    return {"vertices": verts, "existence_probabilities": prob, "edge_probs": padded,
            "edge_indices": idx, "global_features": gfeat, "actual_vertex_counts": dyn}


# --------------------------------------------------------------------------------------
# LSAP (C restatement, oracle/lsap_oracle.c)
# --------------------------------------------------------------------------------------
_LIB = None


def build_c_oracle(force: bool = False) -> str:
    out_dir = os.path.join(_HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "liblsap_oracle.so")
    src = os.path.join(_HERE, "lsap_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-std=c11", "-ffp-contract=off", "-fPIC", "-shared",
                               "-o", so, src, "-lm"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_c_oracle())
        i64p = ctypes.POINTER(ctypes.c_int64)
        _LIB.wfo_lsap_f64.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.c_int64,
                                      ctypes.c_int64, i64p, i64p]
        _LIB.wfo_lsap_f32.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.c_int64,
                                      ctypes.c_int64, i64p, i64p]
        fp = ctypes.POINTER(ctypes.c_float)
        _LIB.wfo_loss_cost_f32.argtypes = [fp, fp, fp, ctypes.c_int64, ctypes.c_int64, fp]
        _LIB.wfo_loss_cost_f32.restype = None
    return _LIB


def lsap(cost: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Same contract as scipy.optimize.linear_sum_assignment(cost) (minimise)."""
    cost = np.asarray(cost)
    if cost.ndim != 2:
        raise ValueError("expected a matrix (2-D array), got a %r array" % (cost.shape,))
    nr, nc = cost.shape
    n = min(nr, nc)
    rows = np.empty(n, dtype=np.int64); cols = np.empty(n, dtype=np.int64)
    i64p = ctypes.POINTER(ctypes.c_int64)
    if cost.dtype == np.float32:
        c = np.ascontiguousarray(cost)
        rc = _lib().wfo_lsap_f32(c.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), nr, nc,
                                 rows.ctypes.data_as(i64p), cols.ctypes.data_as(i64p))
    else:
        c = np.ascontiguousarray(cost, dtype=np.float64)
        rc = _lib().wfo_lsap_f64(c.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), nr, nc,
                                 rows.ctypes.data_as(i64p), cols.ctypes.data_as(i64p))
    if rc == -1:
        raise ValueError("cost matrix is infeasible")
    if rc == -2:
        raise ValueError("matrix contains invalid numeric entries")
    if rc != 0:
        raise MemoryError("lsap oracle")
    return rows, cols


def loss_cost_matrix(pred_v: torch.Tensor, pred_e: torch.Tensor, tgt_v: torch.Tensor, count: int):
    """losses/WireframeLoss.py:142,206-232 -- the LIVE matrix (`final_cost_matrix`, Q9)."""
    V = pred_v.shape[0]
    real = torch.cdist(pred_v, tgt_v[:count], p=1) + (pred_e.unsqueeze(1) - 1.0).abs().expand(-1, count)
    if V - count > 0:
        real = torch.cat([real, pred_e.unsqueeze(1).expand(-1, V - count)], dim=1)
    if count > V:
        real = torch.cat([real, torch.full((count - V, real.shape[1]), float("inf"),
                                           dtype=real.dtype)], dim=0)
    return real


def loss_matching(pred, targets):
    """losses/WireframeLoss.py:106-246 -> list of (pred_idx, tgt_idx) int64 arrays, dummy
    columns filtered out (`col < count`, :240-244)."""
    out = []
    pv, pe = pred["vertices"].detach(), pred["existence_probabilities"].detach()
    for b in range(pv.shape[0]):
        c = int(targets["vertex_counts"][b])
        C = loss_cost_matrix(pv[b].float(), pe[b].float(), targets["vertices"][b].float(), c)
        r, k = lsap(C.cpu().numpy())
        keep = k < c
        out.append((r[keep], k[keep]))
    return out


def loss_forward(pred, targets, vertex_weight=1.0, edge_weight=1.0, existence_weight=1.0,
                 matches=None):
    """losses/WireframeLoss.py:38-104,248-283.  Vertex term: per-sample mean SmoothL1(beta=1)
    times its match count, summed, divided by the batch-total match count.  Existence: BCE mean
    over B*V in slot order.  Edges: BCE mean over the zero-padded (B, min_E) block (Q10)."""
    pv, tv = pred["vertices"], targets["vertices"]
    if matches is None:
        matches = loss_matching(pred, targets)
    tot, n = 0.0, 0
    for b, (pi, ti) in enumerate(matches):
        if len(pi):
            l = F.smooth_l1_loss(pv[b, torch.as_tensor(pi)], tv[b, torch.as_tensor(ti)].to(pv.dtype))
            tot = tot + l * len(pi); n += len(pi)
    vloss = tot / n if n > 0 else torch.zeros((), dtype=pv.dtype)
    eloss = F.binary_cross_entropy(pred["existence_probabilities"],
                                   targets["vertex_existence"].to(pv.dtype))
    pe, te = pred["edge_probs"], targets["edge_labels"].to(pv.dtype)
    dloss = torch.zeros((), dtype=pv.dtype)
    if pe.numel() > 0 and te.numel() > 0:
        m = min(pe.shape[1], te.shape[1])
        if m > 0:
            dloss = F.binary_cross_entropy(pe[:, :m], te[:, :m])
    total = vertex_weight * vloss + existence_weight * eloss + edge_weight * dloss
    return {"total_loss": total, "vertex_loss": vloss, "existence_loss": eloss, "edge_loss": dloss}


# --------------------------------------------------------------------------------------
# stand-alone matchers
# --------------------------------------------------------------------------------------
def wireframe_matcher(outputs, targets: Sequence[dict], cost_vertex=1.0, cost_existence=1.0):
    """models/WireframeHungarianMatcher.py:29-72."""
    pv = outputs["vertices"].detach().float()
    pe = outputs["existence_probabilities"].detach().float()
    res = []
    for b, t in enumerate(targets):
        C = cost_vertex * torch.cdist(pv[b], t["vertices"].float(), p=1) + \
            cost_existence * (pe[b].unsqueeze(1) - t["existence"].float().unsqueeze(0)).abs()
        r, c = lsap(C.numpy())
        res.append((torch.as_tensor(r, dtype=torch.int64), torch.as_tensor(c, dtype=torch.int64)))
    return res


def _xyxy(b):
    cx, cy, w, h = b.unbind(-1)
    return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], dim=-1)


def detr_cost(logits_b, boxes_b, labels, tboxes, cost_class=1.0, cost_bbox=1.0, cost_giou=1.0):
    """models/HungarianMatcher.py:20-55,101-123 for one sample: [Q,T] cost."""
    prob = logits_b.softmax(-1)
    cc = -prob[:, labels]
    cb = torch.cdist(boxes_b, tboxes, p=1)
    a, b = _xyxy(boxes_b), _xyxy(tboxes)
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    lt = torch.max(a[:, None, :2], b[:, :2]); rb = torch.min(a[:, None, 2:], b[:, 2:])
    wh = (rb - lt).clamp(min=0); inter = wh[..., 0] * wh[..., 1]
    union = area_a[:, None] + area_b - inter
    iou = inter / union
    lt2 = torch.min(a[:, None, :2], b[:, :2]); rb2 = torch.max(a[:, None, 2:], b[:, 2:])
    wh2 = (rb2 - lt2).clamp(min=0); hull = wh2[..., 0] * wh2[..., 1]
    giou = iou - (hull - union) / hull
    return cost_bbox * cb + cost_class * cc + cost_giou * (-giou)


def detr_matcher(outputs, targets: Sequence[dict], cost_class=1.0, cost_bbox=1.0, cost_giou=1.0):
    """models/HungarianMatcher.py:80-128."""
    res = []
    for b, t in enumerate(targets):
        C = detr_cost(outputs["pred_logits"][b].float(), outputs["pred_boxes"][b].float(),
                      t["labels"], t["boxes"].float(), cost_class, cost_bbox, cost_giou)
        r, c = lsap(C.numpy())
        res.append((torch.as_tensor(r, dtype=torch.int64), torch.as_tensor(c, dtype=torch.int64)))
    return res


# --------------------------------------------------------------------------------------
# one full training step on the oracle (used by bench.py's cpu_baseline leg)
# --------------------------------------------------------------------------------------
def train_step(sd_params: Dict[str, torch.Tensor], x, targets, *, max_vertices: int,
               weights=(3.0, 1.0, 1.5), emulate_bf16: bool = False, matches=None):
    """fwd + loss + bwd as train.py:127-140 does (weights from train.py:90-94).
    sd_params must hold leaf tensors with requires_grad=True.  Returns the loss dict.
    emulate_bf16: the product's bf16-mode roundings inserted (encoder_point_features); matches: a fixed assignment."""
    pred = model_forward(sd_params, x, targets["vertex_counts"], training=True,
                         max_vertices=max_vertices, emulate_bf16=emulate_bf16)
    ld = loss_forward(pred, targets, vertex_weight=weights[0], edge_weight=weights[1],
                      existence_weight=weights[2], matches=matches)
    ld["total_loss"].backward()
    return ld, pred
