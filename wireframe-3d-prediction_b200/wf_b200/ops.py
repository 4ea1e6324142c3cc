"""torch.autograd.Function wrappers over the C ABI (include/wf_b200.h).

PyTorch owns every tensor (inputs, outputs, saved activations, workspaces) and the autograd graph;
every FLOP of the hot path is done by a kernel of libwf_b200.so.  Each Function's forward/backward
is a short sequence of `_lib.call(...)`s on raw pointers and the current CUDA stream."""
from __future__ import annotations

import ctypes
import os
from ctypes import c_void_p
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import ACT_GELU, ACT_NONE, ACT_RELU, BF16, F32, call

_PRECISION = os.environ.get("WF_B200_PRECISION", "bf16").lower()


def set_precision(p: str) -> None:
    """'bf16': wide encoder layers on tcgen05 tensor cores (production).  'fp32': everything in the
    fp32 SIMT kernels (parity mode: 1e-5-level agreement with the reference's fp32 path)."""
    global _PRECISION
    if p not in ("bf16", "fp32"):
        raise ValueError("precision must be 'bf16' or 'fp32'")
    _PRECISION = p


def get_precision() -> str:
    return _PRECISION


_SMS = {}


def _sm_count() -> int:
    d = torch.cuda.current_device()
    if d not in _SMS:
        _SMS[d] = torch.cuda.get_device_properties(d).multi_processor_count
    return _SMS[d]


LAUNCHES = 0     # kernels launched through this module (bench.py reports it as gpu_launches)


def _count(n: int = 1) -> None:
    global LAUNCHES
    LAUNCHES += n


def _p(t: Optional[torch.Tensor]):
    return None if t is None else c_void_p(t.data_ptr())


def _s():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


_SIDE_STREAMS = {}


def side_stream(device) -> torch.cuda.Stream:
    """One auxiliary stream per device for work that is independent of what the main stream is running
    (the loss matching runs there, beside the edge head)."""
    idx = torch.device(device).index
    idx = torch.cuda.current_device() if idx is None else idx
    if idx not in _SIDE_STREAMS:
        _SIDE_STREAMS[idx] = torch.cuda.Stream(idx)
    return _SIDE_STREAMS[idx]


def mark_ready(t: torch.Tensor, event=None) -> torch.Tensor:
    """Tag a tensor with the CUDA event after which its contents are complete (recorded now on the current stream
    unless given).  Consumers that want to read it from another stream wait on `t._wf_ready` instead of on the whole
    main stream."""
    if event is None:
        event = torch.cuda.Event()
        event.record()
    t._wf_ready = event
    return t

# ---- zero-initialised fp32 buffers (accumulators of the atomically reduced gradients) --------------------------------
# A training step asks for ~80 of them (weight-gradient accumulators of split-K products, LayerNorm gain/shift/bias sums, ...),
# each a separate fill launch.  Between two `new_step()` marks the requests are served as 256-byte-aligned views of ONE
# buffer zeroed by a single fill, sized by what the previous step asked for; a step that asks for more than that (or code that
# never marks steps) falls back to individual torch.zeros.  A buffer is never zeroed twice: views handed out (they become
# parameter .grad tensors) keep their storage alive, the next step simply gets a new buffer.
class _ZeroArena:
    __slots__ = ("buf", "off", "need", "gen")

    def __init__(self):
        self.buf, self.off, self.need, self.gen = None, 0, 0, -1


_ZERO_ARENAS = {}
_STEP_GEN = 0


def new_step() -> None:
    """Marks the start of a training/inference step (PointCloudToWireframe.forward calls it)."""
    global _STEP_GEN
    _STEP_GEN += 1


def zeros_f32(*shape, device) -> torch.Tensor:
    dev = torch.device(device)
    n = 1
    for d in shape:
        n *= int(d)
    a = _ZERO_ARENAS.get(dev.index)
    if a is None:
        a = _ZERO_ARENAS[dev.index] = _ZeroArena()
    if a.gen != _STEP_GEN:
        a.buf = torch.zeros(a.need, device=dev, dtype=torch.float32) if (a.need and a.gen >= 0) else None
        a.off, a.need, a.gen = 0, 0, _STEP_GEN
    n_al = (n + 63) & ~63
    a.need += n_al
    if a.buf is not None and n > 0 and a.off + n <= a.buf.numel():
        v = a.buf[a.off:a.off + n].view(*shape)
        a.off += n_al
        return v
    return torch.zeros(*shape, device=dev, dtype=torch.float32)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.WfError("wf_b200 kernels run on CUDA tensors only; there is no CPU path "
                               "(move the module and its inputs to a B200 device)")


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def _rowmajor(t: torch.Tensor) -> torch.Tensor:
    """2-D fp32 with unit inner stride (views with a row stride are passed through via ld)."""
    if t.dtype != torch.float32:
        t = t.float()
    if t.dim() != 2 or t.stride(1) != 1 or t.stride(0) < t.shape[1]:
        t = t.contiguous()
    return t


# ----------------------------------------------------------------------------------------------
# raw kernels (no autograd)
# ----------------------------------------------------------------------------------------------
USE_TF32_HEADS = os.environ.get("WF_B200_TF32_HEADS", "1") == "1"
_TC_MIN_MACS = 1 << 22
USE_ROWMLP = os.environ.get("WF_B200_ROWMLP", "1") == "1"      # 0: the 64-row heads go back to the split-K kind::tf32 launches
_ROWMLP_MAX_ROWS = 128
_ROWMLP_MIN_WEIGHTS = 1 << 16


def _tc_operand_ok(t: torch.Tensor) -> bool:
    return (t.dtype == torch.float32 and t.dim() == 2 and t.stride(1) == 1 and t.stride(0) % 4 == 0
            and t.data_ptr() % 16 == 0)


def gemm_f32(A, B, *, transA=False, transB=False, out=None, beta=0.0, alpha=1.0, bias=None, tc=True, colsum_out=None):
    """C = alpha * op(A) op(B) + beta * C (+ bias), fp32 storage.
    Production precision ('bf16' mode): large, 16-byte-aligned products run on the tensor cores as TF32
    (wf_gemm_tf32, tcgen05 kind::tf32, fp32 accumulate); everything else, and the whole 'fp32' parity mode,
    runs on the fp32 SIMT kernel (wf_gemm_f32).  tc=False pins a product to the SIMT kernel."""
    A, B = _rowmajor(A), _rowmajor(B)
    M, K = (A.shape[1], A.shape[0]) if transA else A.shape
    N = B.shape[0] if transB else B.shape[1]
    kb = B.shape[1] if transB else B.shape[0]
    if kb != K:
        raise ValueError(f"gemm shape mismatch {tuple(A.shape)} {tuple(B.shape)} tA={transA} tB={transB}")
    if (tc and _PRECISION == "bf16" and USE_TF32_HEADS and USE_ROWMLP and alpha == 1.0 and beta == 0.0 and out is None
            and _tc_operand_ok(A) and _tc_operand_ok(B) and (bias is None or bias.data_ptr() % 16 == 0)):
        # the 64-row heads (csrc/rowmlp.cu): weight-streaming kernels instead of split-K tensor-core launches
        rows = A.shape[0]                      # batch rows in every orientation (A is x, dz or dz)
        if rows <= _ROWMLP_MAX_ROWS:
            if transA and not transB and bias is None and M % 4 == 0 and N % 4 == 0 and M * N >= _ROWMLP_MIN_WEIGHTS:
                out = torch.empty(M, N, device=A.device, dtype=torch.float32)        # dW[Nr, Kc] = A^T B, reduction over the rows
                # colsum_out (a list): the bias gradient db = column sums of A comes out of the same launch
                db = torch.empty(M, device=A.device, dtype=torch.float32) if colsum_out is not None else None
                call("wf_rowmlp_dw", _p(A), A.stride(0), _p(B), B.stride(0), rows, M, N, _p(out), out.stride(0), _p(db), _s())
                if colsum_out is not None:
                    colsum_out.append(db)
                _count()
                return out
            if not transA and N % 4 == 0 and K % 4 == 0 and N * K >= _ROWMLP_MIN_WEIGHTS:
                out = torch.empty(M, N, device=A.device, dtype=torch.float32)        # forward (W [N,K]) or dX (W [K,N], read in place)
                call("wf_rowmlp_linear", _p(A), A.stride(0), _p(B), B.stride(0), int(not transB), _p(bias), M, N, K, _p(out),
                     out.stride(0), _s())
                _count()
                return out
    if (tc and _PRECISION == "bf16" and USE_TF32_HEADS and alpha == 1.0 and beta in (0.0, 1.0) and M * N * K >= _TC_MIN_MACS
            and _tc_operand_ok(A) and _tc_operand_ok(B) and (out is None or _tc_operand_ok(out))
            and (bias is None or bias.data_ptr() % 16 == 0)):
        tiles = ((M + 127) // 128) * ((N + 255) // 256)
        split = 1
        if K >= 512:
            # few output tiles (the 64-row heads stream a whole weight matrix through 2..16 CTAs otherwise): split the
            # reduction over K
            sms = _sm_count()
            if tiles < sms:
                split = max(1, min((2 * sms + tiles - 1) // tiles, K // 128))
        if split > 1 and not transA and N % 4 == 0:
            # forward / dX products: deterministic split-K (partials in a workspace, added in order) -- the forward pass stays
            # bit-reproducible; only weight gradients (transA) use the atomic variant below
            acc = out is not None and beta == 1.0
            if out is None:
                out = torch.empty(M, N, device=A.device, dtype=torch.float32)
            work = torch.empty(split * M * N, device=A.device, dtype=torch.float32)
            call("wf_gemm_tf32_splitk", _p(A), A.stride(0), int(not transA), _p(B), B.stride(0), int(transB), M, N, K, _p(bias),
                 _p(out), out.stride(0), int(acc), int(split), _p(work), work.numel(), _s())
            _count(2)
            return out
        if split > 1 and bias is not None:
            split = 1
        accumulate = split > 1 or (out is not None and beta == 1.0)
        if out is None:
            out = zeros_f32(M, N, device=A.device) if accumulate else torch.empty(M, N, device=A.device, dtype=torch.float32)
        elif beta == 0.0 and accumulate:
            out.zero_()
        call("wf_gemm_tf32", _p(A), A.stride(0), int(not transA), _p(B), B.stride(0), int(transB), M, N, K, _p(bias),
             _p(out), out.stride(0), int(accumulate), int(split), _s())
        _count()
        return out
    if out is None:
        out = torch.empty(M, N, device=A.device, dtype=torch.float32)
        beta = 0.0
    call("wf_gemm_f32", int(transA), int(transB), M, N, K, float(alpha), _p(A), A.stride(0), _p(B),
         B.stride(0), float(beta), _p(out), out.stride(0), _p(bias), _s())
    _count()
    return out


def colsum(x: torch.Tensor) -> torch.Tensor:
    x2 = x.reshape(-1, x.shape[-1])
    out = zeros_f32(x2.shape[1], device=x.device)
    dt = BF16 if x2.dtype == torch.bfloat16 else F32
    call("wf_colsum", _p(x2), dt, x2.shape[0], x2.shape[1], x2.stride(0), _p(out), _s())
    _count()
    return out


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"unsupported dtype {t.dtype}")


# ----------------------------------------------------------------------------------------------
# Linear (+ LayerNorm + activation + dropout + residual)
# ----------------------------------------------------------------------------------------------
class LinearLNAct(torch.autograd.Function):
    """out = dropout(act(LN(x W^T + b))) + residual.   gamma None -> no LayerNorm.
    models/PointNetEncoder.py:37-40,58-65; models/VertexPredictor.py:28-61,105-117;
    models/EdgePredictor.py:31-38,57-68."""

    @staticmethod
    def forward(ctx, x, W, b, gamma, beta, act, residual, keep, keep_scale):
        _need_cuda(x, W)
        x2 = _rowmajor(x.reshape(-1, x.shape[-1]))
        Wc = _rowmajor(W)
        z = gemm_f32(x2, Wc, transB=True, bias=None if b is None else _f32c(b))
        M, C = z.shape
        plain = gamma is None and act == ACT_NONE and keep is None
        mean = rstd = None
        if plain and residual is None:
            out = z
        else:
            out = torch.empty_like(z)
            if gamma is not None:
                mean = torch.empty(M, device=z.device, dtype=torch.float32)
                rstd = torch.empty(M, device=z.device, dtype=torch.float32)
            res = None if residual is None else _f32c(residual.reshape(M, C))
            call("wf_ln_act_fwd", _p(z), F32, _p(gamma), _p(beta), int(act), _p(res), _p(keep), float(keep_scale),
                 _p(out), F32, _p(mean), _p(rstd), 0, M, C, 1e-5, _s())
            _count()
        ctx.save_for_backward(x2, Wc, z if not plain else None, gamma, beta, mean, rstd, keep)
        ctx.meta = (act, keep_scale, plain, b is not None, residual is not None, x.shape)
        return out.reshape(*x.shape[:-1], C)

    @staticmethod
    def backward(ctx, dout):
        x2, Wc, z, gamma, beta, mean, rstd, keep = ctx.saved_tensors
        act, keep_scale, plain, has_b, has_res, xshape = ctx.meta
        C = Wc.shape[0]
        d2 = _f32c(dout.reshape(-1, C))
        M = d2.shape[0]
        dgamma = dbeta = db = None
        want_db_plain = plain and has_b and ctx.needs_input_grad[2]
        if plain:
            dz = d2
        else:
            dz = torch.empty_like(d2)
            if gamma is not None:
                dgamma = zeros_f32(C, device=d2.device)
                dbeta = zeros_f32(C, device=d2.device)
            db = zeros_f32(C, device=d2.device) if has_b else None
            call("wf_ln_act_bwd", _p(d2), F32, _p(z), F32, _p(gamma), _p(beta), _p(mean), _p(rstd), int(act),
                 _p(keep), float(keep_scale), _p(dz), F32, _p(dgamma), _p(dbeta), _p(db), M, C, _s())
            _count()
        dx = dW = None
        if ctx.needs_input_grad[0]:
            dx = gemm_f32(dz, Wc).reshape(xshape)
        if ctx.needs_input_grad[1]:
            got = [] if want_db_plain else None
            dW = gemm_f32(dz, x2, transA=True, colsum_out=got)
            if got:
                db = got[0]                                  # the row-MLP dW kernel also returned the column sums of dz
        if want_db_plain and db is None:
            db = colsum(dz)
        dres = dout if has_res else None
        return dx, dW, db, dgamma, dbeta, None, dres, None, None


def linear_ln_act(x, W, b=None, gamma=None, beta=None, act=ACT_NONE, residual=None, keep=None, keep_scale=1.0):
    return LinearLNAct.apply(x, W, b, gamma, beta, act, residual, keep, keep_scale)


class LNAct(torch.autograd.Function):
    """out = dropout(act(LN(z)))  (no Linear in front: the all-pairs layer produces z itself)."""

    @staticmethod
    def forward(ctx, z, gamma, beta, act, keep, keep_scale):
        _need_cuda(z)
        z2 = _f32c(z.reshape(-1, z.shape[-1]))
        M, C = z2.shape
        out = torch.empty_like(z2)
        mean = torch.empty(M, device=z2.device, dtype=torch.float32)
        rstd = torch.empty(M, device=z2.device, dtype=torch.float32)
        call("wf_ln_act_fwd", _p(z2), F32, _p(gamma), _p(beta), int(act), None, _p(keep), float(keep_scale), _p(out), F32,
             _p(mean), _p(rstd), 0, M, C, 1e-5, _s())
        _count()
        ctx.save_for_backward(z2, gamma, beta, mean, rstd, keep)
        ctx.meta = (act, keep_scale, z.shape)
        return out.reshape(z.shape)

    @staticmethod
    def backward(ctx, dout):
        z2, gamma, beta, mean, rstd, keep = ctx.saved_tensors
        act, keep_scale, zshape = ctx.meta
        M, C = z2.shape
        d2 = _f32c(dout.reshape(M, C))
        dz = torch.empty_like(d2)
        dgamma = zeros_f32(C, device=d2.device)
        dbeta = zeros_f32(C, device=d2.device)
        call("wf_ln_act_bwd", _p(d2), F32, _p(z2), F32, _p(gamma), _p(beta), _p(mean), _p(rstd), int(act), _p(keep),
             float(keep_scale), _p(dz), F32, _p(dgamma), _p(dbeta), None, M, C, _s())
        _count()
        return dz.reshape(zshape), dgamma, dbeta, None, None, None


def dropout_keep(shape, p: float, training: bool, device) -> Tuple[Optional[torch.Tensor], float]:
    """Keep-mask for a dropout site.  The mask bits come from torch's CUDA generator (RNG state is
    plumbing); the masking itself is fused into the consuming kernel."""
    if not training or p <= 0.0:
        return None, 1.0
    keep = torch.empty(shape, device=device, dtype=torch.uint8).bernoulli_(1.0 - p)      # one kernel, bytes straight away
    return keep, 1.0 / (1.0 - p)


# ----------------------------------------------------------------------------------------------
# pooling of point features (fp32 path / public encoder API)
# ----------------------------------------------------------------------------------------------
def point_mask(x: torch.Tensor):
    B, N, D = x.shape
    mask = torch.empty(B, N, device=x.device, dtype=torch.uint8)
    valid = torch.empty(B, device=x.device, dtype=torch.float32)
    call("wf_point_mask", _p(x), B, N, D, _p(mask), _p(valid), _s())
    _count(2)
    return mask, valid


class PoolPoints(torch.autograd.Function):
    """(B,N,C) -> masked max, masked mean, unmasked max, unmasked mean.
    models/PointNetEncoder.py:103-111, models/VertexPredictor.py:86-87."""

    @staticmethod
    def forward(ctx, pf, mask, valid):
        _need_cuda(pf)
        pf = _f32c(pf)
        B, N, C = pf.shape
        mk = lambda dt: torch.empty(B, C, device=pf.device, dtype=dt)
        max_m, avg_m, max_u, mean_u = mk(torch.float32), mk(torch.float32), mk(torch.float32), mk(torch.float32)
        arg_m, arg_u = mk(torch.int32), mk(torch.int32)
        call("wf_pool_fwd", _p(pf), _p(mask), _p(valid), B, N, C, _p(max_m), _p(arg_m), _p(avg_m), _p(max_u),
             _p(arg_u), _p(mean_u), _s())
        _count()
        ctx.save_for_backward(arg_m, arg_u, mask, valid)
        ctx.shape = (B, N, C)
        ctx.mark_non_differentiable(arg_m, arg_u)
        return max_m, avg_m, max_u, mean_u, arg_m, arg_u

    @staticmethod
    def backward(ctx, g_max_m, g_avg_m, g_max_u, g_mean_u, _a, _b):
        arg_m, arg_u, mask, valid = ctx.saved_tensors
        B, N, C = ctx.shape
        d_pf = torch.empty(B, N, C, device=mask.device, dtype=torch.float32)
        # keep the (possibly re-laid-out) gradient tensors alive until the launch: a raw pointer taken from a
        # temporary would dangle once the caching allocator hands the block to the next temporary
        gs = [None if g is None else _f32c(g) for g in (g_max_m, g_avg_m, g_max_u, g_mean_u)]
        call("wf_pool_bwd", _p(gs[0]), _p(gs[1]), _p(gs[2]), _p(gs[3]), _p(arg_m), _p(arg_u),
             _p(mask), _p(valid), B, N, C, _p(d_pf), F32, None, _s())
        _count()
        return d_pf, None, None


# ----------------------------------------------------------------------------------------------
# encoder per-point MLP on tensor cores (bf16 operands, fp32 accumulate) fused with the pools
# ----------------------------------------------------------------------------------------------
GEMM_PROFILE = None   # bench.py sets this to a list: (start event, end event, algorithmic FLOPs) per wf_gemm_bf16 launch


# ---- side jobs: LayerNorm passes of another row chunk executed by spare warps INSIDE a tensor-core GEMM launch ----------
def _dp(t, row0=0):
    """Device address of row `row0` of a 2-D / 1-D tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr() + row0 * t.stride(0) * t.element_size()


def side_ln_fwd(z, mean, rstd, gamma, beta, h, r0, r1, *, colsum=None):
    """Segment: h[r0:r1] = relu(LN(z[r0:r1])) (wf_ln_relu_bf16_fwd).  colsum = (mask [M] u8, part, points_per_cloud,
    global row of row 0 of z): also the per-row-block column sums of h (wf_ln_relu_bf16_fwd_colsum)."""
    sg = _lib.SideSeg()
    sg.kind = _lib.SIDE_LN_FWD if colsum is None else _lib.SIDE_LN_FWD_COLSUM
    sg.C = z.shape[1]; sg.rows = r1 - r0
    sg.x0 = _dp(z, r0); sg.mean = _dp(mean, r0); sg.rstd = _dp(rstd, r0); sg.gamma = _dp(gamma); sg.beta = _dp(beta)
    sg.out = _dp(h, r0)
    if colsum is not None:
        mask, part, pool_n, row_base = colsum
        sg.mask = _dp(mask, r0); sg.part = _dp(part); sg.pool_n = int(pool_n); sg.row_off = int(row_base + r0)
    return sg


def side_ln_bwd(dh, z, mean, rstd, gamma, beta, dz, dgamma, dbeta, dbias, r0, r1):
    """Segment: rows [r0, r1) of wf_ln_relu_bf16_bwd (dgamma / dbeta / dbias are accumulated)."""
    sg = _lib.SideSeg()
    sg.kind = _lib.SIDE_LN_BWD; sg.C = z.shape[1]; sg.rows = r1 - r0
    sg.x0 = _dp(dh, r0); sg.x1 = _dp(z, r0); sg.mean = _dp(mean, r0); sg.rstd = _dp(rstd, r0)
    sg.gamma = _dp(gamma); sg.beta = _dp(beta); sg.out = _dp(dz, r0)
    sg.acc0 = _dp(dgamma); sg.acc1 = _dp(dbeta); sg.acc2 = _dp(dbias)
    return sg


def _side_bytes(side) -> float:
    """Algorithmic HBM bytes of the LayerNorm segments a GEMM launch carries: forward reads z and writes h (4 B / element),
    backward reads dh and z and writes dz (6 B / element)."""
    return float(sum(sg.rows * sg.C * (6 if sg.kind == _lib.SIDE_LN_BWD else 4) for sg in (side or ())))


def _seg_array(side):
    side = [s for s in side if s.rows > 0]
    if len(side) > _lib.SIDE_MAX:
        raise ValueError(f"at most {_lib.SIDE_MAX} side segments per launch")
    arr = (_lib.SideSeg * max(1, len(side)))(*side)
    return arr, len(side)


def gemm_bf16(A, B, *, M, N, K, kmajor=True, bias=None, out, accumulate=False, split_k=1, rowstats=None, side=None):
    lda = A.stride(0)
    ldb = B.stride(0)
    prof = GEMM_PROFILE
    if prof is not None:                      # CUDA events on the launching stream, around this kernel only
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
    if side:
        arr, n = _seg_array(side)
        call("wf_gemm_bf16_side", _p(A), lda, int(kmajor), _p(B), ldb, int(kmajor), M, N, K, _p(bias), _p(out), out.stride(0),
             _dt(out), int(accumulate), int(split_k), _p(rowstats), 0, 0, 0, None, None, None, ctypes.byref(arr), n, _s())
    else:
        call("wf_gemm_bf16", _p(A), lda, int(kmajor), _p(B), ldb, int(kmajor), M, N, K, _p(bias), _p(out), out.stride(0),
             _dt(out), int(accumulate), int(split_k), _p(rowstats), _s())
    if prof is not None:
        e1.record()
        prof.append((e0, e1, 2.0 * M * N * K, _side_bytes(side)))
    _count()
    return out


def gemm_bf16_ownln(A, Wb, *, M, N, K, bias, z, rowstats, gamma, beta, h, mean, rstd, side=None):
    """Linear + LayerNorm + ReLU in one launch (wf_gemm_bf16_ownln): z = A Wb^T + bias (bf16, kept for the backward) and
    h = relu(LN(z)) by the kernel's side warps from the freshly stored tiles (L2), mean / rstd written."""
    prof = GEMM_PROFILE
    if prof is not None:
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
    done = torch.zeros((M + 255) // 256, device=A.device, dtype=torch.int32)
    arr, n = _seg_array(side or [])
    call("wf_gemm_bf16_ownln", _p(A), A.stride(0), _p(Wb), Wb.stride(0), M, N, K, _p(bias), _p(z), _p(rowstats), _p(gamma), _p(beta),
         _p(h), _p(mean), _p(rstd), 1e-5, _p(done), ctypes.byref(arr), n, _s())
    if prof is not None:
        e1.record()
        prof.append((e0, e1, 2.0 * M * N * K, 4.0 * M * N + _side_bytes(side)))
    _count(2)


def gemm_bf16_pool(A, Wb, *, M, N, K, bias, points_per_cloud, row_offset, mask, packed, index_offset=0, side=None):
    """Final per-point Linear whose epilogue max-pools instead of storing (wf_gemm_bf16_pool).  packed: int64 [2, clouds, N]
    (zero-initialised by the caller; [0] = all rows, [1] = valid rows)."""
    prof = GEMM_PROFILE
    if prof is not None:
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
    if side:
        arr, n = _seg_array(side)
        call("wf_gemm_bf16_side", _p(A), A.stride(0), 1, _p(Wb), Wb.stride(0), 1, M, N, K, _p(bias), None, 8, BF16, 0, 1, None,
             int(points_per_cloud), int(row_offset), int(index_offset), _p(mask), _p(packed[0]), _p(packed[1]),
             ctypes.byref(arr), n, _s())
    else:
        call("wf_gemm_bf16_pool", _p(A), A.stride(0), _p(Wb), Wb.stride(0), M, N, K, _p(bias), int(points_per_cloud),
             int(row_offset), int(index_offset), _p(mask), _p(packed[0]), _p(packed[1]), _s())
    if prof is not None:
        e1.record()
        prof.append((e0, e1, 2.0 * M * N * K, _side_bytes(side)))
    _count()


# Data-parallel hook (wf_b200.parallel.GradAllReduce sets it): called with a list of gradient tensors as soon as they are FINAL
# inside the encoder's backward -- layer by layer, while the remaining layers' GEMMs are still to run -- instead of only
# when the whole autograd Function returns.  The tensors are the ones the Function later returns (autograd adopts them as
# .grad), so they can be all-reduced in place right away.
GRAD_READY_HOOK = None


def _grads_ready(tensors):
    hook = GRAD_READY_HOOK
    if hook is not None:
        hook([t for t in tensors if t is not None])


# Test hook: (argmax_masked, argmax_unmasked) int32 [B, 512] used INSTEAD of the computed max-pool argmax for the gradient
# routing of the next EncoderPointMLP_TC.forward.  The parity tests inject the reference's indices with it, so that a
# flipped argmax (a legitimate discontinuity under bf16-sized perturbations) cannot hide a gradient error.
ARGMAX_OVERRIDE = None

FUSED_POOL = os.environ.get("WF_B200_FUSED_POOL", "1") == "1"
_FUSED_MIN_POINTS = 128


def cast_bf16(w: torch.Tensor, transpose: bool = False) -> torch.Tensor:
    w = _f32c(w)
    R, C = w.shape
    out = torch.empty((C, R) if transpose else (R, C), device=w.device, dtype=torch.bfloat16)
    call("wf_cast_bf16", _p(w), R, C, _p(out), int(transpose), _s())
    _count()
    return out


# ---- the per-point MLP as a chunk pipeline with LayerNorm side jobs (sidesched.py, include/wf_b200.h wf_side_seg) ----------
SIDE_JOBS = os.environ.get("WF_B200_SIDE", "1") == "1"
SIDE_CHUNKS = int(os.environ.get("WF_B200_SIDE_CHUNKS", "2"))
SIDE_MIN_ROWS = int(os.environ.get("WF_B200_SIDE_MIN_ROWS", str(256 * 148)))   # below: one chunk, stand-alone LayerNorm passes
SIDE_BW = float(os.environ.get("WF_B200_SIDE_BW", "2.0e12"))                   # HBM bytes/s a side job sustains beside a GEMM
_GEMM_RATE = 1.3e15                                                            # FLOP/s used to estimate launch durations


def _use_side(m: int) -> bool:
    return SIDE_JOBS and m >= max(SIDE_MIN_ROWS, 256)


def _n_chunks(m: int) -> int:
    return max(1, SIDE_CHUNKS) if _use_side(m) else 1


def _ln_fwd_rows(z, mean, rstd, g, be, h, r0, r1, colsum=None):
    """Stand-alone LayerNorm+ReLU forward on rows [r0, r1)."""
    C = z.shape[1]
    if colsum is None:
        call("wf_ln_relu_bf16_fwd", _p(z[r0:r1]), _p(mean[r0:r1]), _p(rstd[r0:r1]), _p(g), _p(be), _p(h[r0:r1]), r1 - r0, C, _s())
    else:
        mask, part, pool_n, row_base = colsum
        call("wf_ln_relu_bf16_fwd_colsum", _p(z[r0:r1]), _p(mean[r0:r1]), _p(rstd[r0:r1]), _p(g), _p(be), _p(h[r0:r1]),
             _p(None if mask is None else mask[r0:r1]), r1 - r0, C, int(pool_n), int(row_base + r0), _p(part), _s())
    _count()


def _enc_mlp_forward(h1, layers, wbs, w5b, b5, *, zs, hs, means, rstds, mask, pool_n, row_base, packed, part, index_offset):
    """Layers 2-5 of the per-point MLP over the m rows of h1 (models/PointNetEncoder.py:37-45, i = 1..3 and the final Linear with
    the pooling epilogue), layer-major over row chunks: G[layer][chunk] launches in order, each carrying LayerNorm rows of
    chunks whose GEMM of the previous launch group is done (sidesched.plan).  zs/hs/means/rstds: per-layer tensors of m rows
    (written).  One chunk (small inputs, WF_B200_SIDE=0): every LayerNorm is a stand-alone pass, as before."""
    from . import sidesched as ss
    m = h1.shape[0]
    chunks = ss.split_rows(m, _n_chunks(m))
    nc, nl = len(chunks), len(layers)
    acts = [h1] + list(hs)                                   # acts[li] = input of layer li (0-based over layers 2..4), acts[3] = h4
    launches, idx = [], {}
    for li in range(nl + 1):
        for c in range(nc):
            idx[(li, c)] = len(launches)
            launches.append((li, c))
    C5, K5 = w5b.shape
    dims = [(W.shape[0], W.shape[1]) for (W, _, _, _) in layers] + [(C5, K5)]
    durs = [2.0 * (chunks[c][1] - chunks[c][0]) * dims[li][0] * dims[li][1] / _GEMM_RATE for (li, c) in launches]
    items = [ss.Item((li, c), chunks[c][0], chunks[c][1], dims[li][0] * 4.0, idx[(li, c)], idx[(li + 1, c)])
             for li in range(nl) for c in range(nc)]
    side, pre = ss.plan_cached(durs, items, SIDE_BW)
    colsum = (mask, part, pool_n, row_base)

    def seg_of(key, r0, r1):
        li = key[0]
        _, _, g, be = layers[li]
        return side_ln_fwd(zs[li], means[li], rstds[li], g, be, hs[li], r0, r1, colsum=colsum if li == nl - 1 else None)

    for j, (li, c) in enumerate(launches):
        for key, r0, r1 in pre.get(j, ()):
            l2 = key[0]
            _, _, g, be = layers[l2]
            _ln_fwd_rows(zs[l2], means[l2], rstds[l2], g, be, hs[l2], r0, r1, colsum if l2 == nl - 1 else None)
        segs = [seg_of(*t) for t in side[j]]
        r0, r1 = chunks[c]
        mc = r1 - r0
        if li < nl:
            W, b, _, _ = layers[li]
            Nn, K = W.shape
            parts = call("wf_gemm_rowstats_parts", Nn)
            stats = torch.empty(parts, mc, 2, device=h1.device, dtype=torch.float32)
            gemm_bf16(acts[li][r0:r1], wbs[li], M=mc, N=Nn, K=K, bias=b, out=zs[li][r0:r1], rowstats=stats, side=segs)
            call("wf_stats_finalize", _p(stats), mc, Nn, parts, 1e-5, _p(means[li][r0:r1]), _p(rstds[li][r0:r1]), _s())
            _count()
        else:
            gemm_bf16_pool(acts[nl][r0:r1], w5b, M=mc, N=C5, K=K5, bias=b5, points_per_cloud=pool_n, row_offset=row_base + r0,
                           mask=None if mask is None else mask[r0:r1], packed=packed, index_offset=index_offset, side=segs)


def _enc_mlp_backward(dh_top, hs, zs, means, rstds, layers, sms):
    """Backward of layers 2-4 of the per-point MLP: per layer (last first) LayerNorm+ReLU backward, dX = dz W and dW = dz^T h.
    dh_top: gradient w.r.t. h4 (m, 1024) bf16; hs = [h1..h4], zs = [z2..z4].  Row chunks: the dX launches of a layer come
    first (they make the next layer's dh available), then its dW launches; every launch carries LayerNorm-backward rows whose
    dh already exists (sidesched.plan).  Returns (grads dict keyed W/b/g/be + layer number 2..4, dh1)."""
    from . import sidesched as ss
    dev = dh_top.device
    m = dh_top.shape[0]
    chunks = ss.split_rows(m, _n_chunks(m))
    nc, nl = len(chunks), len(layers)
    order = list(range(nl - 1, -1, -1))                      # layer index li: nl-1 (layer 4) first
    launches, idx = [], {}
    for li in order:
        for kind in ("dX", "dW"):
            for c in range(nc):
                idx[(kind, li, c)] = len(launches)
                launches.append((kind, li, c))
    durs = [2.0 * (chunks[c][1] - chunks[c][0]) * layers[li][0].shape[0] * layers[li][0].shape[1] / _GEMM_RATE for (_, li, c) in launches]
    items = []
    for li in order:
        for c in range(nc):
            avail = -1 if li == nl - 1 else idx[("dX", li + 1, c)]
            if not _use_side(m):
                avail = idx[("dX", li, c)] - 1               # no window: a stand-alone pass right before its consumer
            items.append(ss.Item((li, c), chunks[c][0], chunks[c][1], layers[li][0].shape[0] * 6.0, avail, idx[("dX", li, c)]))
    side, pre = ss.plan_cached(durs, items, SIDE_BW)
    dhs = {nl - 1: dh_top}
    dzs, grads, dWs, wts = {}, {}, {}, {}
    for li in range(nl):
        W = layers[li][0]
        Nn, K = W.shape
        dzs[li] = torch.empty(m, Nn, device=dev, dtype=torch.bfloat16)
        k = li + 2
        grads[f"g{k}"], grads[f"be{k}"], grads[f"b{k}"] = zeros_f32(Nn, device=dev), zeros_f32(Nn, device=dev), zeros_f32(Nn, device=dev)
        dWs[li] = grads[f"W{k}"] = zeros_f32(Nn, K, device=dev)
        if li > 0:
            dhs[li - 1] = torch.empty(m, K, device=dev, dtype=torch.bfloat16)
    dh1 = torch.empty(m, layers[0][0].shape[1], device=dev, dtype=torch.bfloat16)

    def ln_args(li):
        _, g, be = layers[li][0], layers[li][1], layers[li][2]
        k = li + 2
        return (dhs[li], zs[li], means[li], rstds[li], g, be, dzs[li], grads[f"g{k}"], grads[f"be{k}"], grads[f"b{k}"])

    for j, (kind, li, c) in enumerate(launches):
        for key, r0, r1 in pre.get(j, ()):
            dh, z, mean, rstd, g, be, dz, dg, dbe, db = ln_args(key[0])
            call("wf_ln_relu_bf16_bwd", _p(dh[r0:r1]), _p(z[r0:r1]), _p(mean[r0:r1]), _p(rstd[r0:r1]), _p(g), _p(be), _p(dz[r0:r1]),
                 _p(dg), _p(dbe), _p(db), r1 - r0, z.shape[1], _s())
            _count()
        segs = [side_ln_bwd(*ln_args(key[0]), r0, r1) for key, r0, r1 in side[j]]
        r0, r1 = chunks[c]
        mc = r1 - r0
        W = layers[li][0]
        Nn, K = W.shape
        if kind == "dX":
            if li not in wts:
                wts[li] = cast_bf16(W, transpose=True)
            out = dh1 if li == 0 else dhs[li - 1]
            gemm_bf16(dzs[li][r0:r1], wts[li], M=mc, N=K, K=Nn, out=out[r0:r1], side=segs)
        else:
            tiles = ((Nn + 127) // 128) * ((K + 255) // 256)
            split = max(1, min((mc + 63) // 64, (2 * sms + tiles - 1) // tiles))
            gemm_bf16(dzs[li][r0:r1], hs[li][r0:r1], M=Nn, N=K, K=mc, kmajor=False, out=dWs[li], accumulate=True, split_k=split,
                      side=segs)
            if c == nc - 1:                                  # every chunk of this layer's LayerNorm backward and dW is enqueued
                k = li + 2
                _grads_ready([grads[f"W{k}"], grads[f"b{k}"], grads[f"g{k}"], grads[f"be{k}"]])
    return grads, dh1


INFER_CHUNK_ROWS = int(os.environ.get("WF_B200_INFER_CHUNK_ROWS", str(1 << 19)))
# L2-resident form: chunks small enough that every N x C activation lives and dies in the 126 MB L2.  37 x 256 rows make the
# 256 x 256 cluster tiles of the 1024 / 2048 / 512-wide layers 148 / 296 / 74 per chunk: whole waves of the 74 CTA pairs.
INFER_L2 = os.environ.get("WF_B200_INFER_L2", "0") == "1"
INFER_L2_ROWS = int(os.environ.get("WF_B200_INFER_L2_ROWS", str(37 * 256)))


def encoder_pooled_infer(x, params, *, chunk_rows=None, index_offset=0, points_total=None, reduce_fn=None, l2_resident=None):
    """Inference form of EncoderPointMLP_TC (no autograd, nothing saved): the per-point MLP runs over row chunks of at most
    `chunk_rows` points through three reused bf16 buffers, so a batch of million-point scans (BASELINE.json configs[3]:
    8 x 1M points = 139 GB of training activations) needs ~2.5 KB/point of the largest chunk instead.  The pools accumulate
    across chunks: packed maxima by atomicMax, column sums by row block (deterministic).

    index_offset / points_total / reduce_fn: clouds sharded by POINTS across ranks (SURVEY 8e, config 4 with B < world):
    this rank holds points [index_offset, index_offset + N) of every cloud out of points_total; reduce_fn(packed int64
    [2,B,C], hsum fp32 [2,B,K], cnt fp32 [B]) combines the ranks' partial pools in place (wf_b200.parallel.reduce_pool_shards:
    integer MAX and SUM all-reduces) before they are finalised.
    Returns (max_m, avg_m, max_u, mean_u, arg_m, arg_u) like models/PointNetEncoder.py:103-111 + VertexPredictor.py:86-87."""
    (W1, b1, g1, be1, W2, b2, g2, be2, W3, b3, g3, be3, W4, b4, g4, be4, W5, b5) = [t.detach() for t in params]
    _need_cuda(x, W1)
    x = _f32c(x.detach())
    B, N, D = x.shape
    M = B * N
    dev = x.device
    if N < _FUSED_MIN_POINTS:
        raise _lib.WfError(f"encoder_pooled_infer needs >= {_FUSED_MIN_POINTS} points per cloud (got {N})")
    l2 = INFER_L2 if l2_resident is None else bool(l2_resident)
    if l2 and chunk_rows is None:
        chunk_rows = INFER_L2_ROWS
    chunk = INFER_CHUNK_ROWS if chunk_rows is None else int(chunk_rows)
    chunk = max(128, (min(chunk, M) + 127) // 128 * 128)
    # rows are processed in super-chunks (one set of activation buffers, 17.4 KB per row, reused); inside a super-chunk the
    # layers run as a chunk pipeline whose GEMM launches carry the LayerNorm passes (_enc_mlp_forward)
    sup = chunk if chunk_rows is not None else min(M, max(1, SIDE_CHUNKS if SIDE_JOBS else 1) * chunk)
    if chunk_rows is None and M > sup:
        # equal super-chunks instead of full ones plus a short remainder; rounded to the 256-row cluster tile
        n_sup = -(-M // sup)
        sup = min(sup, (-(-M // n_sup) + 255) // 256 * 256)
    mask, cnt = point_mask(x)                                  # cnt = max(#valid, 1)
    layers = ((W2, b2, g2, be2), (W3, b3, g3, be3), (W4, b4, g4, be4))
    wbs = [cast_bf16(W) for (W, _, _, _) in layers]
    w5b = cast_bf16(W5)
    C5, K5 = W5.shape
    rows = min(sup, M)
    bf = lambda c: torch.empty(rows, c, device=dev, dtype=torch.bfloat16)
    if l2:
        # two buffers, ping-pong: X (widest layer) holds h1, then z3 -> h3 in place; Y holds z2 -> h2, then z4 -> h4.  The
        # footprint (rows x (Cmax + C2) x 2 B = 58 MB at 9472 rows) is rewritten chunk after chunk while still dirty in L2,
        # so in steady state neither the GEMM stores nor the LayerNorm passes move DRAM bytes.
        widths = [W.shape[0] for (W, _, _, _) in layers]
        if not (len(widths) == 3 and widths[0] == widths[2] and W1.shape[0] <= widths[1]):
            raise _lib.WfError("l2_resident: unexpected layer widths")
        bx, by = bf(widths[1]), bf(widths[0])
        h1 = bx.view(-1)[: rows * W1.shape[0]].view(rows, W1.shape[0])
        zs = [by, bx, by]
        hs = [by, bx, by]
    else:
        h1 = bf(W1.shape[0])
        zs = [bf(W.shape[0]) for (W, _, _, _) in layers]
        hs = [bf(W.shape[0]) for (W, _, _, _) in layers]
    means = [torch.empty(rows, device=dev, dtype=torch.float32) for _ in layers]
    rstds = [torch.empty(rows, device=dev, dtype=torch.float32) for _ in layers]
    part = torch.empty(call("wf_seg_part_floats", M, K5), device=dev, dtype=torch.float32)
    packed = torch.zeros(2, B, C5, device=dev, dtype=torch.int64)
    xf = x.view(M, D)
    mflat = mask.view(M)
    W1c = _f32c(W1)
    for r0 in range(0, M, sup):
        m = min(sup, M - r0)
        call("wf_enc_l1_fwd", _p(xf[r0:]), _p(W1c), _p(b1), _p(g1), _p(be1), _p(h1), BF16, m, D, W1.shape[0], 1e-5, _s())
        _count()
        _enc_mlp_forward(h1[:m], layers, wbs, w5b, b5, zs=[t[:m] for t in zs], hs=[t[:m] for t in hs],
                         means=[t[:m] for t in means], rstds=[t[:m] for t in rstds], mask=mflat[r0:r0 + m], pool_n=N,
                         row_base=r0, packed=packed, part=part, index_offset=index_offset)
    hbar = torch.empty(2 * B, K5, device=dev, dtype=torch.float32)
    ntot = N if points_total is None else int(points_total)
    if reduce_fn is None:
        call("wf_seg_mean", _p(part), _p(cnt), B, N, K5, _p(hbar), _s())
        _count()
    else:
        ones = torch.ones(B, device=dev, dtype=torch.float32)
        call("wf_seg_mean", _p(part), _p(ones), B, N, K5, _p(hbar), _s())     # hbar = [sum / N ; masked sum / 1]
        _count()
        hsum = hbar.view(2, B, K5)
        hsum[0].mul_(float(N))                                 # back to plain sums; the division happens after the reduce
        mk = mask.view(B, N).sum(dim=1, dtype=torch.float32)   # true valid counts of this shard (cnt is clamped to >= 1)
        reduce_fn(packed, hsum, mk)
        hsum[0].div_(float(ntot))
        hsum[1].div_(mk.clamp_min(1.0).unsqueeze(1))
    lin = gemm_f32(hbar, _f32c(W5), transB=True, tc=False)
    mkt = lambda dt: torch.empty(B, C5, device=dev, dtype=dt)
    max_m, avg_m, max_u, mean_u = mkt(torch.float32), mkt(torch.float32), mkt(torch.float32), mkt(torch.float32)
    arg_m, arg_u = mkt(torch.int32), mkt(torch.int32)
    call("wf_pool_finalize", _p(packed[0]), _p(packed[1]), _p(lin), _p(b5), B, C5, _p(max_m), _p(arg_m), _p(avg_m),
         _p(max_u), _p(arg_u), _p(mean_u), _s())
    _count()
    return max_m, avg_m, max_u, mean_u, arg_m, arg_u


class EncoderPointMLP_TC(torch.autograd.Function):
    """x (B,N,8) -> the four pooled reductions of the (B,N,512) point features, without keeping the
    fp32 point-feature tensor.  models/PointNetEncoder.py:85-111 + models/VertexPredictor.py:86-87.

    Layer 1 (K=8) runs in fp32 SIMT (un-normalised intensity); layers 2..5 are
    wf_gemm_bf16 (tcgen05).  LayerNorm statistics come out of the GEMM epilogue (fp32 accumulators).
    Saved for backward: bf16 pre-LN z2..z4 and post-ReLU h1..h4 (17.4 KB/point), row statistics,
    argmax indices."""

    @staticmethod
    def forward(ctx, x, want_pf, *params):
        (W1, b1, g1, be1, W2, b2, g2, be2, W3, b3, g3, be3, W4, b4, g4, be4, W5, b5) = params
        _need_cuda(x, W1)
        x = _f32c(x)
        B, N, D = x.shape
        M = B * N
        dev = x.device
        mask, valid = point_mask(x)
        h = torch.empty(M, W1.shape[0], device=dev, dtype=torch.bfloat16)
        W1c = _f32c(W1)
        call("wf_enc_l1_fwd", _p(x), _p(W1c), _p(b1), _p(g1), _p(be1), _p(h), BF16, M, D, W1.shape[0], 1e-5, _s())
        _count()
        C5, K5 = W5.shape
        fused = FUSED_POOL and not want_pf and N >= _FUSED_MIN_POINTS and 2 * C5 <= 1024
        hs, zs, means, rstds = [h], [], [], []
        part = None
        layers = ((W2, b2, g2, be2), (W3, b3, g3, be3), (W4, b4, g4, be4))
        if fused:
            # layers 2-5 as a chunk pipeline: GEMM launches carry the LayerNorm passes of other chunks (_enc_mlp_forward)
            bf = lambda c: torch.empty(M, c, device=dev, dtype=torch.bfloat16)
            zs = [bf(W.shape[0]) for (W, _, _, _) in layers]
            hn = [bf(W.shape[0]) for (W, _, _, _) in layers]
            means = [torch.empty(M, device=dev, dtype=torch.float32) for _ in layers]
            rstds = [torch.empty(M, device=dev, dtype=torch.float32) for _ in layers]
            part = torch.empty(call("wf_seg_part_floats", M, layers[-1][0].shape[0]), device=dev, dtype=torch.float32)
            packed = torch.zeros(2, B, C5, device=dev, dtype=torch.int64)
            w5b = cast_bf16(W5)
            _enc_mlp_forward(h, layers, [cast_bf16(W) for (W, _, _, _) in layers], w5b, b5, zs=zs, hs=hn, means=means,
                             rstds=rstds, mask=mask.view(M), pool_n=N, row_base=0, packed=packed, part=part, index_offset=0)
            hs = [h] + hn
        for li, (W, b, g, be) in enumerate(layers if not fused else ()):
            Nn, K = W.shape
            wb = cast_bf16(W)
            z = torch.empty(M, Nn, device=dev, dtype=torch.bfloat16)
            parts = call("wf_gemm_rowstats_parts", Nn)
            stats = torch.empty(parts, M, 2, device=dev, dtype=torch.float32)
            gemm_bf16(hs[-1], wb, M=M, N=Nn, K=K, bias=b, out=z, rowstats=stats)
            mean = torch.empty(M, device=dev, dtype=torch.float32)
            rstd = torch.empty(M, device=dev, dtype=torch.float32)
            call("wf_stats_finalize", _p(stats), M, Nn, parts, 1e-5, _p(mean), _p(rstd), _s())
            hn = torch.empty(M, Nn, device=dev, dtype=torch.bfloat16)
            call("wf_ln_relu_bf16_fwd", _p(z), _p(mean), _p(rstd), _p(g), _p(be), _p(hn), M, Nn, _s())
            _count(2)
            hs.append(hn); zs.append(z); means.append(mean); rstds.append(rstd)
        hbar = None
        if fused:
            # max pools came out of the final GEMM's epilogue, mean pools go through the affine map: the (B,N,512) tensor never exists
            hbar = torch.empty(2 * B, K5, device=dev, dtype=torch.float32)
            call("wf_seg_mean", _p(part), _p(valid), B, N, K5, _p(hbar), _s())
            lin = gemm_f32(hbar, _f32c(W5), transB=True, tc=False)
            mk = lambda dt: torch.empty(B, C5, device=dev, dtype=dt)
            max_m, avg_m, max_u, mean_u = mk(torch.float32), mk(torch.float32), mk(torch.float32), mk(torch.float32)
            arg_m, arg_u = mk(torch.int32), mk(torch.int32)
            call("wf_pool_finalize", _p(packed[0]), _p(packed[1]), _p(lin), _p(b5), B, C5, _p(max_m), _p(arg_m), _p(avg_m),
                 _p(max_u), _p(arg_u), _p(mean_u), _s())
            _count(2)
            pf = None
        else:
            pf = torch.empty(M, C5, device=dev, dtype=torch.float32)
            gemm_bf16(hs[-1], cast_bf16(W5), M=M, N=C5, K=K5, bias=b5, out=pf)
            pooled = PoolPoints.forward(_Scratch(), pf.view(B, N, -1), mask, valid)
            max_m, avg_m, max_u, mean_u, arg_m, arg_u = pooled
        sv_m, sv_u = (arg_m, arg_u) if ARGMAX_OVERRIDE is None else ARGMAX_OVERRIDE
        ctx.save_for_backward(x, mask, valid, sv_m, sv_u, *hs, *zs, *means, *rstds, *params)
        ctx.hbar = hbar
        ctx.dims = (B, N, D)
        ctx.mark_non_differentiable(arg_m, arg_u)
        pf_out = pf.view(B, N, -1) if want_pf else x.new_empty(0)
        return max_m, avg_m, max_u, mean_u, arg_m, arg_u, pf_out

    @staticmethod
    def backward(ctx, g_max_m, g_avg_m, g_max_u, g_mean_u, _a, _b, g_pf):
        sv = ctx.saved_tensors
        x, mask, valid, arg_m, arg_u = sv[:5]
        hs = sv[5:9]; zs = sv[9:12]; means = sv[12:15]; rstds = sv[15:18]
        (W1, b1, g1, be1, W2, b2, g2, be2, W3, b3, g3, be3, W4, b4, g4, be4, W5, b5) = sv[18:]
        B, N, D = ctx.dims
        M = B * N
        dev = x.device
        gs = [None if g is None else _f32c(g) for g in (g_max_m, g_avg_m, g_max_u, g_mean_u)]   # held until the launch
        C5, K5 = W5.shape
        grads = {}
        sms = _sm_count()

        def weight_grad(dzl, hin, Nn, K):
            dW = zeros_f32(Nn, K, device=dev)
            tiles = ((Nn + 127) // 128) * ((K + 255) // 256)
            split = max(1, min((M + 63) // 64, (2 * sms + tiles - 1) // tiles))
            gemm_bf16(dzl, hin, M=Nn, N=K, K=M, kmajor=False, out=dW, accumulate=True, split_k=split)
            return dW

        hbar = ctx.hbar
        if hbar is not None:
            # pools -> final Linear, analytically (wf_pool_fused_bwd): no dense (M,512) gradient, no dX/dW GEMM for layer 5
            G = zeros_f32(2 * B, C5, device=dev)
            if gs[3] is not None:
                G[:B].copy_(gs[3])
            if gs[1] is not None:
                G[B:].copy_(gs[1])
            W5c = _f32c(W5)
            dbar = gemm_f32(G, W5c, tc=False)                          # (2B, K5)
            dW5 = gemm_f32(G, hbar, transA=True, tc=False)             # (C5, K5), argmax rows added below
            db5 = torch.empty(C5, device=dev, dtype=torch.float32)
            work = torch.empty(call("wf_pool_fused_bwd_work_ints", B, C5), device=dev, dtype=torch.int32)
            dh = torch.empty(M, K5, device=dev, dtype=torch.bfloat16)
            call("wf_pool_fused_bwd", _p(gs[0]), _p(gs[1]), _p(gs[2]), _p(gs[3]), _p(arg_m), _p(arg_u), _p(mask), _p(valid),
                 _p(dbar), _p(W5c), _p(hs[3]), B, N, C5, K5, _p(work), _p(dh), _p(dW5), _p(db5), _s())
            _count(4)
            grads["W5"], grads["b5"] = dW5, db5
            _grads_ready([dW5, db5])
        else:
            dz = torch.empty(M, C5, device=dev, dtype=torch.bfloat16)
            db5 = zeros_f32(C5, device=dev)
            call("wf_pool_bwd", _p(gs[0]), _p(gs[1]), _p(gs[2]), _p(gs[3]), _p(arg_m), _p(arg_u),
                 _p(mask), _p(valid), B, N, C5, _p(dz), BF16, _p(db5), _s())
            _count()
            if g_pf is not None and g_pf.numel() > 0:
                dz = (dz.float() + g_pf.reshape(M, C5)).to(torch.bfloat16)      # only when a caller used point_features
                db5 = db5 + g_pf.reshape(M, C5).sum(0)
            # layer 5 (no LayerNorm)
            grads["W5"] = weight_grad(dz, hs[3], C5, K5)
            grads["b5"] = db5
            dh = torch.empty(M, K5, device=dev, dtype=torch.bfloat16)
            gemm_bf16(dz, cast_bf16(W5, transpose=True), M=M, N=K5, K=C5, out=dh)
        g2to4, dh = _enc_mlp_backward(dh, hs, zs, means, rstds, ((W2, g2, be2), (W3, g3, be3), (W4, g4, be4)), sms)
        grads.update(g2to4)
        C1 = W1.shape[0]
        dW1 = zeros_f32(C1, D, device=dev)
        db1 = zeros_f32(C1, device=dev)
        dg1 = zeros_f32(C1, device=dev)
        dbe1 = zeros_f32(C1, device=dev)
        dx = torch.empty(M, D, device=dev, dtype=torch.float32) if ctx.needs_input_grad[0] else None
        W1c = _f32c(W1)
        call("wf_enc_l1_bwd", _p(x), _p(W1c), _p(b1), _p(g1), _p(be1), _p(dh), BF16, _p(dW1), _p(db1), _p(dg1),
             _p(dbe1), _p(dx), M, D, C1, 1e-5, _s())
        _count()
        grads.update(W1=dW1, b1=db1, g1=dg1, be1=dbe1)
        order = ["W1", "b1", "g1", "be1", "W2", "b2", "g2", "be2", "W3", "b3", "g3", "be3", "W4", "b4", "g4", "be4",
                 "W5", "b5"]
        return (None if dx is None else dx.view(B, N, D), None, *[grads[k] for k in order])


class _Scratch:
    """Stand-in ctx for calling a Function's forward as a plain kernel sequence."""

    def save_for_backward(self, *a):
        pass

    def mark_non_differentiable(self, *a):
        pass


# ----------------------------------------------------------------------------------------------
# vertex head tail
# ----------------------------------------------------------------------------------------------
class VertexSplit(torch.autograd.Function):
    """models/VertexPredictor.py:117-127."""

    @staticmethod
    def forward(ctx, vf, V):
        _need_cuda(vf)
        vf = _f32c(vf)
        B = vf.shape[0]
        coords = torch.empty(B, V, 3, device=vf.device, dtype=torch.float32)
        prob = torch.empty(B, V, device=vf.device, dtype=torch.float32)
        count = torch.empty(B, device=vf.device, dtype=torch.int64)
        call("wf_vertex_split_fwd", _p(vf), B, V, _p(coords), _p(prob), _p(count), _s())
        _count()
        ctx.save_for_backward(prob)
        ctx.V = V
        ctx.mark_non_differentiable(count)
        return coords, prob, count

    @staticmethod
    def backward(ctx, d_coords, d_prob, _c):
        (prob,) = ctx.saved_tensors
        B, V = prob.shape
        d_vf = torch.empty(B, V * 4, device=prob.device, dtype=torch.float32)
        dc = None if d_coords is None else _f32c(d_coords)
        dp = None if d_prob is None else _f32c(d_prob)
        call("wf_vertex_split_bwd", _p(dc), _p(dp), _p(prob), B, V, _p(d_vf), _s())
        _count()
        return d_vf, None


# ----------------------------------------------------------------------------------------------
# edge head (ragged)
# ----------------------------------------------------------------------------------------------
class Ragged:
    """Host-known per-sample vertex counts -> CSR offsets on the device."""

    def __init__(self, counts: Sequence[int], device):
        self.counts = [int(c) for c in counts]
        B = len(self.counts)
        v_off = [0] * (B + 1); e_off = [0] * (B + 1); p_off = [0] * (B + 1)
        for b, c in enumerate(self.counts):
            v_off[b + 1] = v_off[b] + c
            e_off[b + 1] = e_off[b] + c * (c - 1) // 2
            p_off[b + 1] = p_off[b] + 8 * c * c
        self.B, self.T, self.E, self.Ptot = B, v_off[B], e_off[B], p_off[B]
        self.max_c = max(self.counts) if self.counts else 0
        self.max_e = self.max_c * (self.max_c - 1) // 2
        self.v_off = torch.tensor(v_off, dtype=torch.int32).to(device, non_blocking=True)
        self.e_off = torch.tensor(e_off, dtype=torch.int64).to(device, non_blocking=True)
        self.p_off = torch.tensor(p_off, dtype=torch.int64).to(device, non_blocking=True)


class GatherPrefix(torch.autograd.Function):
    """verts[b, :count_b] for all b, packed -- models/PointCloudToWireframe.py:81,91."""

    @staticmethod
    def forward(ctx, verts, rg: Ragged):
        _need_cuda(verts)
        verts = _f32c(verts)
        B, V, _ = verts.shape
        out = torch.empty(rg.T, 3, device=verts.device, dtype=torch.float32)
        call("wf_gather_prefix", _p(verts), B, V, _p(rg.v_off), rg.T, _p(out), _s())
        _count()
        ctx.rg = rg; ctx.shape = (B, V)
        return out

    @staticmethod
    def backward(ctx, d_packed):
        B, V = ctx.shape
        d = zeros_f32(B, V, 3, device=d_packed.device)
        dpk = _f32c(d_packed)
        call("wf_scatter_prefix_add", _p(dpk), B, V, _p(ctx.rg.v_off), ctx.rg.T, _p(d), _s())
        _count()
        return d, None


class AttentionCore(torch.autograd.Function):
    """softmax(q k^T / sqrt(d)) v for `heads` heads, per sample -- the core of nn.MultiheadAttention
    (models/EdgePredictor.py:109-111); in_proj / out_proj are LinearLNAct calls around it."""

    @staticmethod
    def forward(ctx, qkv, rg: Ragged, keep, keep_scale, heads=8):
        _need_cuda(qkv)
        qkv = _f32c(qkv)
        E = qkv.shape[1] // 3
        ctx.heads = heads = int(heads)
        out = torch.empty(rg.T, E, device=qkv.device, dtype=torch.float32)
        probs = torch.empty(rg.Ptot, device=qkv.device, dtype=torch.float32)
        call("wf_attn_fwd", _p(qkv), _p(rg.v_off), _p(rg.p_off), rg.B, heads, E // heads, rg.max_c, _p(out), _p(probs), _p(keep),
             float(keep_scale), _s())
        _count()
        ctx.save_for_backward(qkv, probs, keep)
        ctx.rg = rg; ctx.keep_scale = keep_scale
        return out

    @staticmethod
    def backward(ctx, d_out):
        qkv, probs, keep = ctx.saved_tensors
        rg = ctx.rg
        E = qkv.shape[1] // 3
        d_qkv = torch.empty_like(qkv)
        d_out = _f32c(d_out)
        call("wf_attn_bwd", _p(d_out), _p(qkv), _p(probs), _p(rg.v_off), _p(rg.p_off), rg.B, ctx.heads, E // ctx.heads, rg.max_c,
             _p(d_qkv), _p(keep), float(ctx.keep_scale), _s())
        _count()
        return d_qkv, None, None, None, None


class EdgePairLayer(torch.autograd.Function):
    """z1[e=(i,j)] = P[i] + Q[j] + wd * |v_i - v_j| + b  (models/EdgePredictor.py:117-134 + edge_mlp.0
    without the (E,1031) concat).  PQ = [P | Q] is the ONE stacked [T, 2C] product of the vertex features with both feature
    blocks of the layer: the kernels read the halves in place (row stride 2C) and write their gradients into one [T, 2C]
    tensor, so neither the halves nor their gradients are copied apart or added back together by autograd."""

    @staticmethod
    def forward(ctx, PQ, verts, wd, bias, rg: Ragged):
        _need_cuda(PQ)
        PQ, verts, wd, bias = _f32c(PQ), _f32c(verts), _f32c(wd), _f32c(bias)
        C = PQ.shape[1] // 2
        z1 = torch.empty(rg.E, C, device=PQ.device, dtype=torch.float32)
        dist = torch.empty(rg.E, device=PQ.device, dtype=torch.float32)
        call("wf_edge_pair_fwd", _p(PQ), c_void_p(PQ.data_ptr() + 4 * C), _p(verts), _p(wd), _p(bias), _p(rg.v_off), _p(rg.e_off),
             rg.B, rg.T, C, 2 * C, _p(z1), _p(dist), _s())
        _count()
        ctx.save_for_backward(dist, verts, wd)
        ctx.rg = rg
        return z1

    @staticmethod
    def backward(ctx, dz1):
        dist, verts, wd = ctx.saved_tensors
        rg = ctx.rg
        C = wd.shape[0]
        dz1 = _f32c(dz1)
        dPQ = torch.empty(rg.T, 2 * C, device=dz1.device, dtype=torch.float32)
        dv = zeros_f32(rg.T, 3, device=dz1.device)
        dwd = zeros_f32(C, device=dz1.device)
        call("wf_edge_pair_bwd", _p(dz1), _p(dist), _p(verts), _p(wd), _p(rg.v_off), _p(rg.e_off), rg.B, rg.T, C, 2 * C, _p(dPQ),
             c_void_p(dPQ.data_ptr() + 4 * C), _p(dv), _p(dwd), _s())
        _count()
        dbias = colsum(dPQ[:, :C])                           # every pair contributes once to exactly one dP row
        return dPQ, dv, dwd, dbias, None


class SplitPairWeight(torch.autograd.Function):
    """edge_mlp.0.weight [H, 2H + 7] -> (Wfq [2H, H]: the two feature blocks stacked, Wvq [2H, 3]: the two coordinate blocks
    stacked, wd [H]: the distance column) as contiguous operands, and back: ONE assembly of the weight gradient instead of
    autograd's slice / cat backward per use (a zero-filled [H, 2H + 7] tensor, a copy and an add for each of the five slices)."""

    @staticmethod
    def forward(ctx, W1):
        H = W1.shape[0]
        ctx.H = H
        Wfq = torch.cat([W1[:, :H], W1[:, H:2 * H]], dim=0)
        Wvq = torch.cat([W1[:, 2 * H:2 * H + 3], W1[:, 2 * H + 3:2 * H + 6]], dim=0)
        return Wfq, Wvq, W1[:, 2 * H + 6].contiguous()

    @staticmethod
    def backward(ctx, dWfq, dWvq, dwd):
        H = ctx.H
        ref = next(g for g in (dWfq, dWvq, dwd) if g is not None)
        dW1 = torch.empty(H, 2 * H + 7, device=ref.device, dtype=ref.dtype)
        if dWfq is not None:
            dW1[:, :H] = dWfq[:H]; dW1[:, H:2 * H] = dWfq[H:]
        else:
            dW1[:, :2 * H] = 0
        if dWvq is not None:
            dW1[:, 2 * H:2 * H + 3] = dWvq[:H]; dW1[:, 2 * H + 3:2 * H + 6] = dWvq[H:]
        else:
            dW1[:, 2 * H:2 * H + 6] = 0
        if dwd is not None:
            dW1[:, 2 * H + 6] = dwd
        else:
            dW1[:, 2 * H + 6] = 0
        return dW1


class EdgeOut(torch.autograd.Function):
    """edge_mlp.10 (128 -> 1) + sigmoid, written zero-padded as (B, max_e):
    models/EdgePredictor.py:137-138, models/PointCloudToWireframe.py:103-112."""

    @staticmethod
    def forward(ctx, h, w, bias, rg: Ragged):
        _need_cuda(h)
        h, w, bias = _f32c(h), _f32c(w).reshape(-1), _f32c(bias)
        probs = torch.empty(rg.B, rg.max_e, device=h.device, dtype=torch.float32)
        call("wf_edge_out_fwd", _p(h), _p(w), _p(bias), _p(rg.e_off), rg.B, w.shape[0], rg.max_e, _p(probs), _s())
        _count()
        ctx.save_for_backward(h, w, probs)
        ctx.rg = rg
        return probs

    @staticmethod
    def backward(ctx, d_probs):
        h, w, probs = ctx.saved_tensors
        rg = ctx.rg
        dh = torch.empty_like(h)
        dw = zeros_f32(*w.shape, device=w.device)
        db = zeros_f32(1, device=h.device)
        d_probs = _f32c(d_probs)
        call("wf_edge_out_bwd", _p(d_probs), _p(probs), _p(h), _p(w), _p(rg.e_off), rg.B, w.shape[0], rg.max_e,
             _p(dh), _p(dw), _p(db), _s())
        _count()
        return dh, dw.reshape(1, -1), db, None


# ----------------------------------------------------------------------------------------------
# matching + loss
# ----------------------------------------------------------------------------------------------
def loss_match(pred_v, pred_e, tgt_v, counts, want_cost: bool = False):
    """losses/WireframeLoss.py:106-246 on the device.  Returns (col_of_row[B,V] int32, status[B] int32, cost|None)."""
    _need_cuda(pred_v)
    pv, pe, tv = _f32c(pred_v.detach()), _f32c(pred_e.detach()), _f32c(tgt_v)
    B, V, _ = pv.shape
    Vt = tv.shape[1]
    cnt = counts.to(device=pv.device, dtype=torch.int64).contiguous()
    col = torch.empty(B, V, device=pv.device, dtype=torch.int32)
    status = torch.empty(B, device=pv.device, dtype=torch.int32)
    cost = torch.empty(B, V, V, device=pv.device, dtype=torch.float32) if want_cost else None
    call("wf_loss_match", _p(pv), _p(pe), _p(tv), _p(cnt), B, V, Vt, _p(col), _p(status), _p(cost), _s())
    _count()
    return col, status, cost


def lsap_batched(cost: torch.Tensor, nr: torch.Tensor, nc: torch.Tensor):
    """cost [B, max_nr, ld] fp32 on the device; nr/nc int32 [B].  Returns (col_of_row [B,max_nr], status [B])."""
    _need_cuda(cost)
    cost = _f32c(cost)
    B, max_nr, ld = cost.shape
    max_nc = int(ld)
    col = torch.empty(B, max_nr, device=cost.device, dtype=torch.int32)
    status = torch.empty(B, device=cost.device, dtype=torch.int32)
    call("wf_lsap_batched", _p(cost), max_nr * ld, ld, _p(nr), _p(nc), B, max_nr, max_nc, _p(col), _p(status), _s())
    _count()
    return col, status


class WireframeLossFn(torch.autograd.Function):
    """losses/WireframeLoss.py:38-104,248-283: returns (total, vertex, existence, edge)."""

    @staticmethod
    def forward(ctx, pred_v, pred_e, edge_p, tgt_v, tgt_e, edge_l, col_of_row, counts, wv, we, wx):
        _need_cuda(pred_v)
        pv, pe, ep = _f32c(pred_v), _f32c(pred_e), _f32c(edge_p)
        tv, te, el = _f32c(tgt_v), _f32c(tgt_e), _f32c(edge_l)
        B, V, _ = pv.shape
        Ep = ep.shape[1] if ep.dim() == 2 else 0
        El = el.shape[1] if el.dim() == 2 else 0
        if ep.numel() == 0 or el.numel() == 0:
            Ep_eff, El_eff = (Ep, 0) if el.numel() == 0 else (0, El)
        else:
            Ep_eff, El_eff = Ep, El
        out = torch.empty(call("wf_loss_out_floats"), device=pv.device, dtype=torch.float32)
        cnt = counts.to(device=pv.device, dtype=torch.int64).contiguous()
        call("wf_loss_fwd", _p(pv), _p(pe), _p(ep), _p(tv), _p(te), _p(el), _p(col_of_row), _p(cnt), B, V, tv.shape[1],
             Ep_eff if Ep_eff else 0, El_eff if El_eff else 0, float(wv), float(we), float(wx), _p(out), _s())
        _count(2)
        ctx.save_for_backward(pv, pe, ep, tv, te, el, col_of_row, cnt, out)
        ctx.meta = (B, V, tv.shape[1], Ep, El, Ep_eff, El_eff, wv, we, wx)
        return out[0], out[1], out[2], out[3]

    @staticmethod
    def backward(ctx, g_t, g_v, g_x, g_e):
        pv, pe, ep, tv, te, el, col, cnt, out = ctx.saved_tensors
        B, V, Vt, Ep, El, Ep_eff, El_eff, wv, we, wx = ctx.meta
        z = lambda g: torch.zeros((), device=pv.device) if g is None else g.reshape(()).float()
        g_out = torch.stack([z(g_t), z(g_v), z(g_x), z(g_e)]).contiguous()
        d_pv = torch.empty_like(pv); d_pe = torch.empty_like(pe); d_ep = torch.empty_like(ep)
        # Ep_eff/El_eff are 0 when either side is empty (edge term off); d_edge_p must still be zero-filled
        if Ep_eff == 0 or El_eff == 0:
            d_ep.zero_()
            call("wf_loss_bwd", _p(g_out), _p(pv), _p(pe), _p(ep), _p(tv), _p(te), _p(el), _p(col), _p(cnt), _p(out), B, V,
                 Vt, 0, 0, float(wv), float(we), float(wx), _p(d_pv), _p(d_pe), _p(d_ep), _s())
        else:
            call("wf_loss_bwd", _p(g_out), _p(pv), _p(pe), _p(ep), _p(tv), _p(te), _p(el), _p(col), _p(cnt), _p(out), B, V,
                 Vt, Ep, El, float(wv), float(we), float(wx), _p(d_pv), _p(d_pe), _p(d_ep), _s())
        _count()
        return d_pv, d_pe, d_ep, None, None, None, None, None, None, None, None
