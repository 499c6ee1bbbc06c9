"""ctypes binding of librrt_b200.so -- one-to-one with include/rrt_b200.h.

The library is the product's only compute path.  It is loaded lazily on first use and there is no
fallback: a missing or stale library raises ``RuntimeError`` telling the user to build it.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librrt_b200.so")

RRT_ABI_VERSION = 8
RRT_DROP_STREAM_CRMSA = 64
RRT_DROP_STREAM_PATCH = 65
RRT_DROP_STREAM_POOL = 66
RRT_POS_NONE, RRT_POS_PEG, RRT_POS_PPEG = 0, 1, 2
RRT_EPEG_ATTN, RRT_EPEG_VALUE_BF, RRT_EPEG_VALUE_AF = 0, 1, 2
RRT_MAX_RMSA_LAYERS = 8
RRT_MAX_CRMSA_K = 16
RRT_MAX_EPEG_K = 63
RRT_MAX_LANES = 8
RRT_OK, RRT_E_INVALID, RRT_E_WORKSPACE, RRT_E_CUDA = 0, -1, -2, -3
RRT_MATH_F16 = 0
RRT_ACT_NONE, RRT_ACT_RELU, RRT_ACT_GELU, RRT_ACT_TANH = 0, 1, 2, 3
RRT_ACT_GATED = 0x100

c_float_p = C.c_void_p  # device pointers travel as integers


class RrtConfig(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("n_rmsa_layers", C.c_int32), ("n_heads", C.c_int32),
        ("region_num", C.c_int32), ("region_size", C.c_int32), ("min_region_num", C.c_int32),
        ("min_region_ratio", C.c_double), ("epeg", C.c_int32), ("epeg_k", C.c_int32),
        ("qkv_bias", C.c_int32), ("cr_msa", C.c_int32), ("crmsa_k", C.c_int32),
        ("crmsa_heads", C.c_int32), ("crmsa_mlp", C.c_int32), ("all_shortcut", C.c_int32),
        ("math_mode", C.c_int32),
        ("pos", C.c_int32), ("pos_pos", C.c_int32), ("peg_k", C.c_int32), ("peg_1d", C.c_int32),
        ("ffn", C.c_int32), ("ffn_act", C.c_int32), ("ffn_hidden", C.c_int32),
        ("epeg_type", C.c_int32), ("epeg_2d", C.c_int32),
    ]


class RrtAttnWeights(C.Structure):
    _fields_ = [("qkv_w", c_float_p), ("qkv_b", c_float_p), ("proj_w", c_float_p),
                ("proj_b", c_float_p), ("pe_w", c_float_p), ("qkv_w_f16", c_float_p),
                ("proj_w_f16", c_float_p), ("pe_b", c_float_p)]


class RrtFfnWeights(C.Structure):
    _fields_ = [(n, c_float_p) for n in ("norm_w", "norm_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b",
                                         "fc1_w_f16", "fc2_w_f16")]


class RrtWeights(C.Structure):
    _fields_ = [
        ("norm_w", c_float_p), ("norm_b", c_float_p),
        ("layer_norm_w", c_float_p * RRT_MAX_RMSA_LAYERS),
        ("layer_norm_b", c_float_p * RRT_MAX_RMSA_LAYERS),
        ("layer_attn", RrtAttnWeights * RRT_MAX_RMSA_LAYERS),
        ("cr_norm_w", c_float_p), ("cr_norm_b", c_float_p), ("cr_phi", c_float_p),
        ("cr_phi_w1", c_float_p), ("cr_phi_w2", c_float_p), ("cr_phi_w1_f16", c_float_p),
        ("cr_attn", RrtAttnWeights),
        ("pos_w", c_float_p * 3), ("pos_b", c_float_p * 3),
        ("layer_ffn", RrtFfnWeights * RRT_MAX_RMSA_LAYERS), ("cr_ffn", RrtFfnWeights),
    ]


class RrtAdamTensor(C.Structure):
    _fields_ = [("param", c_float_p), ("grad", c_float_p), ("exp_avg", c_float_p),
                ("exp_avg_sq", c_float_p), ("n", C.c_int64)]


class RrtAttnGrads(C.Structure):
    _fields_ = [("qkv_w", c_float_p), ("qkv_b", c_float_p), ("proj_w", c_float_p),
                ("proj_b", c_float_p), ("pe_w", c_float_p)]


class RrtGrads(C.Structure):
    _fields_ = [
        ("norm_w", c_float_p), ("norm_b", c_float_p),
        ("layer_norm_w", c_float_p * RRT_MAX_RMSA_LAYERS),
        ("layer_norm_b", c_float_p * RRT_MAX_RMSA_LAYERS),
        ("layer_attn", RrtAttnGrads * RRT_MAX_RMSA_LAYERS),
        ("cr_norm_w", c_float_p), ("cr_norm_b", c_float_p), ("cr_phi", c_float_p),
        ("cr_attn", RrtAttnGrads),
        ("cr_phi_w1", c_float_p), ("cr_phi_w2", c_float_p),
        ("pos_w", c_float_p * 3), ("pos_b", c_float_p * 3),
    ]


# name -> (restype, argtypes); must list every RRT_API symbol of include/rrt_b200.h
_P = C.c_void_p
SIGNATURES = {
    "rrt_abi_version": (C.c_int, []),
    "rrt_last_error": (C.c_char_p, []),
    "rrt_grid_geometry": (C.c_int, [C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_double,
                                    C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "rrt_workspace_bytes": (C.c_int, [C.POINTER(RrtConfig), C.c_int64, C.POINTER(C.c_size_t)]),
    "rrt_encoder_forward": (C.c_int, [C.POINTER(RrtConfig), C.POINTER(RrtWeights), _P, _P,
                                      C.c_int64, _P, C.c_size_t, _P]),
    "rrt_encoder_forward_batch": (C.c_int, [C.POINTER(RrtConfig), C.POINTER(RrtWeights),
                                            C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                            C.POINTER(C.c_int64), C.c_int32, _P, C.c_size_t, _P]),
    "rrt_encoder_forward_host": (C.c_int, [C.POINTER(RrtConfig), C.POINTER(RrtWeights), _P, _P, _P,
                                           _P, C.c_int64, _P, C.c_size_t, _P]),
    "rrt_rmsa_block_forward": (C.c_int, [C.POINTER(RrtConfig), _P, _P, C.POINTER(RrtAttnWeights),
                                         _P, _P, C.c_int64, _P, C.c_size_t, _P]),
    "rrt_crmsa_block_forward": (C.c_int, [C.POINTER(RrtConfig), C.POINTER(RrtWeights), _P, _P, _P,
                                          C.c_int64, C.c_int32, _P, C.c_size_t, _P]),
    "rrt_train_tape_bytes": (C.c_int, [C.POINTER(RrtConfig), C.c_int64, C.POINTER(C.c_size_t)]),
    "rrt_encoder_forward_train": (C.c_int, [C.POINTER(RrtConfig), C.POINTER(RrtWeights), _P, _P,
                                            C.c_int64, _P, C.c_size_t, C.c_float, C.c_uint64,
                                            C.POINTER(C.c_float), _P]),
    "rrt_peg_forward": (C.c_int, [_P, _P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                  C.POINTER(c_float_p), C.POINTER(c_float_p), _P]),
    "rrt_linear_wgrad_f16": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int32, C.c_int32, _P]),
    "rrt_adam_step": (C.c_int, [C.POINTER(RrtAdamTensor), C.c_int32, C.c_float, C.c_float, C.c_float,
                                C.c_float, C.c_float, C.c_int32, C.c_int64, C.c_float, _P]),
    "rrt_dropout_mask": (C.c_int, [_P, C.c_int64, C.c_float, C.c_uint64, C.c_uint32, _P]),
    "rrt_backward_supported": (C.c_int, [C.POINTER(RrtConfig), C.c_int64]),
    "rrt_backward_workspace_bytes": (C.c_int, [C.POINTER(RrtConfig), C.c_int64, C.POINTER(C.c_size_t)]),
    "rrt_encoder_backward": (C.c_int, [C.POINTER(RrtConfig), C.POINTER(RrtWeights), _P, _P, C.c_int64,
                                       _P, C.c_size_t, C.POINTER(RrtGrads), _P, _P, C.c_size_t,
                                       C.c_float, C.c_uint64, C.POINTER(C.c_float), _P]),
    "rrt_attention_backward": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32,
                                         C.c_int32, C.c_int32, _P]),
    "rrt_layernorm_backward": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int64, C.c_int32, _P]),
    "rrt_mil_head_workspace_bytes": (C.c_int, [C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]),
    "rrt_patch_embed_forward": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, _P, _P, _P, C.c_int32, _P, _P,
                                          C.c_size_t, C.c_float, C.c_uint64, _P, _P]),
    "rrt_patch_embed_backward": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_float,
                                           C.c_uint64, _P, C.c_size_t, _P, _P, _P, C.c_size_t, _P]),
    "rrt_mil_head_backward_workspace_bytes": (C.c_int, [C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]),
    "rrt_attn_pool_backward": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, _P, C.c_int32, _P, _P, C.c_int32,
                                         _P, _P, C.c_float, C.c_uint64, _P, _P, C.c_size_t, _P, _P, _P, _P, _P, _P,
                                         _P, _P, C.c_size_t, _P]),
    "rrt_attn_pool_forward": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, _P, _P, _P, C.c_int32, _P, _P, _P, _P,
                                        C.c_int32, _P, _P, _P, C.c_int32, C.c_float, C.c_uint64, _P, _P, C.c_size_t,
                                        _P]),
    "rrt_set_step_state": (C.c_int, [_P]),
    "rrt_launch_count": (C.c_int64, []),
    "rrt_stage_timing_enable": (C.c_int, [C.c_int32]),
    "rrt_stage_count": (C.c_int32, []),
    "rrt_stage_name": (C.c_char_p, [C.c_int32]),
    "rrt_stage_timing_read": (C.c_int, [C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "rrt_debug_set_gemm_trace": (C.c_int, [_P]),
    "rrt_debug_set_attn_trace": (C.c_int, [_P]),
    "rrt_debug_set_attention_kernel": (C.c_int, [C.c_int32]),
    "rrt_debug_set_gemm_cluster": (C.c_int, [C.c_int32]),
    "rrt_debug_skip_stages": (C.c_int, [C.c_uint32]),
    "rrt_convert_f16": (C.c_int, [_P, _P, C.c_int64, _P]),
    "rrt_widen_f32": (C.c_int, [_P, C.c_int32, _P, C.c_int64, _P]),
    "rrt_linear_f16_forward": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int32, C.c_int32, _P]),
    "rrt_linear_forward": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int32, C.c_int32, _P]),
    "rrt_layernorm_forward": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int32, _P]),
}

_lib = None
_lock = threading.Lock()


class RrtError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load (once) and return the C-ABI library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: build it with `python -m rrt_mil_b200.build` "
                    "(there is no CPU or PyTorch fallback for the RRTEncoder hot path)")
            handle = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(handle, name)  # AttributeError if the .so is stale
                fn.restype, fn.argtypes = res, args
            if handle.rrt_abi_version() != RRT_ABI_VERSION:
                raise RuntimeError("librrt_b200.so ABI version mismatch: rebuild it")
            if os.environ.get("RRT_ATTN"):  # tuning knob: auto (default) | tc05 (wherever supported) | mma
                handle.rrt_debug_set_attention_kernel({"mma": 0, "auto": 1, "tc05": 2}[os.environ["RRT_ATTN"]])
            if os.environ.get("RRT_GEMM_CLUSTER"):  # tuning knob: 11 (default), 2, 21, 22, 128, 256
                handle.rrt_debug_set_gemm_cluster(int(os.environ["RRT_GEMM_CLUSTER"]))
            _lib = handle
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != RRT_OK:
        msg = lib().rrt_last_error().decode("utf-8", "replace")
        exc = {RRT_E_INVALID: ValueError, RRT_E_WORKSPACE: RrtError, RRT_E_CUDA: RrtError}.get(rc, RrtError)
        raise exc(f"{what or 'librrt_b200'} failed ({rc}): {msg}")


def grid_geometry(L: int, region_num: int, region_size: int = 0, min_region_num: int = 0,
                  min_region_ratio: float = 0.0):
    H, rs = C.c_int32(), C.c_int32()
    check(lib().rrt_grid_geometry(L, region_num, region_size, min_region_num, min_region_ratio,
                                  C.byref(H), C.byref(rs)), "rrt_grid_geometry")
    return H.value, rs.value


def workspace_bytes(cfg: RrtConfig, L: int) -> int:
    n = C.c_size_t()
    check(lib().rrt_workspace_bytes(C.byref(cfg), L, C.byref(n)), "rrt_workspace_bytes")
    return n.value


def stage_timing(enable: bool) -> None:
    check(lib().rrt_stage_timing_enable(int(enable)), "rrt_stage_timing_enable")


def read_stage_timing() -> dict:
    """{stage name: (total ms, intervals)} accumulated since ``stage_timing(True)``; synchronise the
    stream(s) first."""
    out = {}
    L = lib()
    for i in range(L.rrt_stage_count()):
        ms, n = C.c_double(), C.c_int64()
        check(L.rrt_stage_timing_read(i, C.byref(ms), C.byref(n)), "rrt_stage_timing_read")
        if n.value:
            out[L.rrt_stage_name(i).decode()] = (ms.value, n.value)
    return out


def launch_count() -> int:
    return int(lib().rrt_launch_count())
