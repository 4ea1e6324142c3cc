"""ctypes binding of libwf_b200.so (C ABI declared in include/wf_b200.h).

The library is the ONLY compute path of this package.  If it is missing or a call fails the
error is raised -- there is no PyTorch/CPU fallback anywhere (BASELINE.json north_star)."""
from __future__ import annotations

import ctypes
import os
from ctypes import c_float, c_int, c_int64, c_void_p

_PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_PKG_DIR, "libwf_b200.so")

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
LSAP_OK, LSAP_INFEASIBLE, LSAP_INVALID = 0, 1, 2

P = c_void_p
I = c_int
L = c_int64
F = c_float

# name -> argtypes (return type is int unless listed in _RESTYPE)
SIGNATURES = {
    "wf_version": [],
    "wf_last_error": [],
    "wf_device_info": [P, P, P],
    "wf_loss_match": [P, P, P, P, I, I, I, P, P, P, P],
    "wf_lsap_batched": [P, L, I, P, P, I, I, I, P, P, P],
    "wf_wireframe_matcher_cost": [P, P, P, P, P, I, I, F, F, P, I, P],
    "wf_detr_matcher_cost": [P, P, P, P, P, I, I, I, F, F, F, P, I, P],
    "wf_gemm_f32": [I, I, I, I, I, F, P, I, P, I, F, P, I, P, P],
    "wf_ln_act_fwd": [P, I, P, P, I, P, P, F, P, I, P, P, I, I, I, F, P],
    "wf_ln_act_bwd": [P, I, P, I, P, P, P, P, I, P, F, P, I, P, P, P, I, I, P],
    "wf_colsum": [P, I, I, I, I, P, P],
    "wf_vertex_split_fwd": [P, I, I, P, P, P, P],
    "wf_vertex_split_bwd": [P, P, P, I, I, P, P],
    "wf_point_mask": [P, I, I, I, P, P, P],
    "wf_enc_l1_fwd": [P, P, P, P, P, P, I, I, I, I, F, P],
    "wf_enc_l1_bwd": [P, P, P, P, P, P, I, P, P, P, P, P, I, I, I, F, P],
    "wf_gemm_bf16": [P, I, I, P, I, I, I, I, I, P, P, I, I, I, I, P, P],
    "wf_gemm_tf32": [P, I, I, P, I, I, I, I, I, P, P, I, I, I, P],
    "wf_gemm_tf32_splitk": [P, I, I, P, I, I, I, I, I, P, P, I, I, I, P, L, P],
    "wf_rowmlp_linear": [P, I, P, I, I, P, I, I, I, P, I, P],
    "wf_rowmlp_dw": [P, I, P, I, I, I, I, P, I, P, P],
    "wf_gemm_bf16_side": [P, I, I, P, I, I, I, I, I, P, P, I, I, I, I, P, I, I, I, P, P, P, P, I, P],
    "wf_gemm_bf16_ownln": [P, I, P, I, I, I, I, P, P, P, P, P, P, P, P, F, P, P, I, P],
    "wf_ln_relu_bf16_fwd": [P, P, P, P, P, P, I, I, P],
    "wf_ln_relu_bf16_bwd": [P, P, P, P, P, P, P, P, P, P, I, I, P],
    "wf_stats_finalize": [P, I, I, I, F, P, P, P],
    "wf_gemm_rowstats_parts": [I],
    "wf_cast_bf16": [P, I, I, P, I, P],
    "wf_pool_fwd": [P, P, P, I, I, I, P, P, P, P, P, P, P],
    "wf_pool_bwd": [P, P, P, P, P, P, P, P, I, I, I, P, I, P, P],
    "wf_gemm_bf16_pool": [P, I, P, I, I, I, I, P, I, I, I, P, P, P, P],
    "wf_ln_relu_bf16_fwd_colsum": [P, P, P, P, P, P, P, I, I, I, I, P, P],
    "wf_seg_part_floats": [I, I],
    "wf_seg_mean": [P, P, I, I, I, P, P],
    "wf_pool_finalize": [P, P, P, P, I, I, P, P, P, P, P, P, P],
    "wf_pool_fused_bwd": [P, P, P, P, P, P, P, P, P, P, P, I, I, I, I, P, P, P, P, P],
    "wf_pool_fused_bwd_work_ints": [I, I],
    "wf_pack_targets": [P, P, P, P, I, I, I, I, P, P, P, P, P],
    "wf_gather_prefix": [P, I, I, P, I, P, P],
    "wf_scatter_prefix_add": [P, I, I, P, I, P, P],
    "wf_attn_fwd": [P, P, P, I, I, I, I, P, P, P, F, P],
    "wf_attn_bwd": [P, P, P, P, P, I, I, I, I, P, P, F, P],
    "wf_edge_pair_fwd": [P, P, P, P, P, P, P, I, I, I, I, P, P, P],
    "wf_edge_pair_bwd": [P, P, P, P, P, P, I, I, I, I, P, P, P, P, P],
    "wf_edge_out_fwd": [P, P, P, P, I, I, I, P, P],
    "wf_edge_out_bwd": [P, P, P, P, P, I, I, I, P, P, P, P],
    "wf_hausdorff_lines": [P, P, P, P, P, I, I, P, I, P, P],
    "wf_cdist_f64": [P, P, P, P, P, I, L, I, P, P],
    "wf_lsap_f64": [P, P, P, P, P, I, I, I, P, P, P, P, P],
    "wf_loss_out_floats": [],
    "wf_loss_fwd": [P, P, P, P, P, P, P, P, I, I, I, I, I, F, F, F, P, P],
    "wf_loss_bwd": [P, P, P, P, P, P, P, P, P, P, I, I, I, I, I, F, F, F, P, P, P, P],
}
SIDE_LN_FWD, SIDE_LN_FWD_COLSUM, SIDE_LN_BWD, SIDE_MAX = 1, 2, 3, 4


class SideSeg(ctypes.Structure):
    """include/wf_b200.h: wf_side_seg -- one side-job segment of wf_gemm_bf16_side."""
    _fields_ = [("kind", ctypes.c_int32), ("C", ctypes.c_int32), ("rows", ctypes.c_int64),
                ("x0", c_void_p), ("x1", c_void_p), ("mean", c_void_p), ("rstd", c_void_p), ("gamma", c_void_p),
                ("beta", c_void_p), ("out", c_void_p), ("acc0", c_void_p), ("acc1", c_void_p), ("acc2", c_void_p),
                ("mask", c_void_p), ("part", c_void_p), ("pool_n", ctypes.c_int32), ("row_off", ctypes.c_int32)]


_RESTYPE = {"wf_last_error": ctypes.c_char_p}
_NO_STATUS = {"wf_version", "wf_last_error", "wf_loss_out_floats", "wf_seg_part_floats", "wf_pool_fused_bwd_work_ints",
              "wf_gemm_rowstats_parts"}

_lib = None


class WfError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Works without a GPU -- only symbol binding happens here."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise WfError(
            f"{LIB_PATH} not found: build it with `python __graft_entry__.py build` "
            "(or `make -C wireframe-3d-prediction_b200/csrc`). There is no fallback path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export it
        fn.argtypes = args
        fn.restype = _RESTYPE.get(name, c_int)
    _lib = lib
    return lib


def call(name: str, *args):
    """Invoke an entry point; non-zero status raises WfError with the library's message."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if name in _NO_STATUS:
        return rc
    if rc != 0:
        msg = lib.wf_last_error()
        raise WfError(f"{name} failed (status {rc}): {msg.decode() if msg else '?'}")
    return rc
