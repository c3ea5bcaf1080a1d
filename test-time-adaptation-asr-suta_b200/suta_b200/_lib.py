"""ctypes binding of libsuta_b200.so (the C ABI declared in include/suta_b200.h).

There is no fallback: if the shared library is missing or a CUDA device is absent the product path raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SUTA_B200_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libsuta_b200.so")   # override: instrumented builds

MAX_LAYERS = 48
MAX_CONV = 8
ABI_VERSION = 4

c_void_p, c_int, c_int32, c_int64, c_float = C.c_void_p, C.c_int, C.c_int32, C.c_int64, C.c_float


class ModelCfg(C.Structure):
    _fields_ = [("hidden", c_int32), ("layers", c_int32), ("heads", c_int32), ("intermediate", c_int32),
                ("vocab", c_int32), ("n_conv", c_int32), ("conv_dim", c_int32 * MAX_CONV),
                ("conv_kernel", c_int32 * MAX_CONV), ("conv_stride", c_int32 * MAX_CONV),
                ("pos_k", c_int32), ("pos_groups", c_int32), ("ln_eps", c_float),
                ("feat_norm_layer", c_int32), ("stable_layer_norm", c_int32)]


class LayerWeights(C.Structure):
    _fields_ = [(n, c_void_p) for n in ("wqkv", "wqkv_t", "wo", "wo_t", "w1", "w1_t", "w2", "w2_t",
                                         "bqkv", "bo", "b1", "b2")]


class Weights(C.Structure):
    _fields_ = [("conv0_w", c_void_p), ("gn_g", c_void_p), ("gn_b", c_void_p), ("conv_b", c_void_p * MAX_CONV),
                ("conv_w", c_void_p * MAX_CONV), ("conv_w_t", c_void_p * MAX_CONV),
                ("proj_w", c_void_p), ("proj_w_t", c_void_p), ("proj_b", c_void_p),
                ("pos_w", c_void_p), ("pos_w_t", c_void_p), ("pos_b", c_void_p),
                ("layer", LayerWeights * MAX_LAYERS),
                ("lm_w", c_void_p), ("lm_w_t", c_void_p), ("lm_b", c_void_p),
                ("params0", c_void_p), ("mult", c_void_p)]


class ParamSeg(C.Structure):
    _fields_ = [("kind", c_int32), ("module", c_int32), ("index", c_int32), ("offset", c_int64), ("size", c_int64)]


class Hyper(C.Structure):
    _fields_ = [("em_coef", c_float), ("temp", c_float), ("reweight", c_int32), ("not_blank", c_int32),
                ("opt_kind", c_int32), ("lr", c_float), ("beta1", c_float), ("beta2", c_float), ("eps", c_float),
                ("weight_decay", c_float), ("div_coef", c_float), ("pl_coef", c_float)]


# name -> (restype, argtypes); must list every symbol include/suta_b200.h declares (tests check this)
SIGNATURES = {
    "suta_last_error": (C.c_char_p, []),
    "suta_abi_version": (c_int, []),
    "suta_device_sm_count": (c_int, []),
    "suta_engine_create": (c_int, [C.POINTER(ModelCfg), c_int, C.POINTER(c_void_p)]),
    "suta_engine_destroy": (None, [c_void_p]),
    "suta_engine_param_count": (c_int64, [c_void_p]),
    "suta_engine_param_layout": (c_int, [c_void_p, C.POINTER(ParamSeg), c_int, C.POINTER(c_int)]),
    "suta_engine_set_weights": (c_int, [c_void_p, C.POINTER(Weights)]),
    "suta_batch_workspace_bytes": (c_int64, [c_void_p, c_int, C.POINTER(c_int32)]),
    "suta_batch_begin": (c_int, [c_void_p, c_int, C.POINTER(c_int32), c_void_p, c_int64, c_void_p]),
    "suta_batch_info": (c_int, [c_void_p, C.POINTER(c_int64), C.POINTER(c_int32), C.POINTER(c_int64),
                                C.POINTER(c_int64), C.POINTER(c_int64)]),
    "suta_batch_set_audio": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "suta_batch_add_noise": (c_int, [c_void_p, c_float, C.c_uint64, C.POINTER(c_int32), c_void_p]),
    "suta_reset": (c_int, [c_void_p, c_void_p]),
    "suta_frontend": (c_int, [c_void_p, c_void_p]),
    "suta_forward": (c_int, [c_void_p, c_void_p]),
    "suta_loss_backward": (c_int, [c_void_p, C.POINTER(Hyper), c_void_p]),
    "suta_optimizer_step": (c_int, [c_void_p, C.POINTER(Hyper), c_void_p]),
    "suta_decode": (c_int, [c_void_p, c_void_p]),
    "suta_adapt_step": (c_int, [c_void_p, C.POINTER(Hyper), c_void_p]),
    "suta_logits": (c_void_p, [c_void_p]),
    "suta_dlogits": (c_void_p, [c_void_p]),
    "suta_losses": (c_void_p, [c_void_p]),
    "suta_params": (c_void_p, [c_void_p]),
    "suta_grads": (c_void_p, [c_void_p]),
    "suta_argmax_ids": (c_void_p, [c_void_p]),
    "suta_collapsed_ids": (c_void_p, [c_void_p]),
    "suta_collapsed_len": (c_void_p, [c_void_p]),
    "suta_params_written": (c_int, [c_void_p, c_void_p]),
    "suta_adam_exp_avg": (c_void_p, [c_void_p]),
    "suta_adam_exp_avg_sq": (c_void_p, [c_void_p]),
    "suta_opt_steps": (c_int, [c_void_p]),
    "suta_set_opt_steps": (c_int, [c_void_p, c_int]),
    "suta_debug_buffer": (c_void_p, [c_void_p, C.c_char_p, C.POINTER(c_int64), C.POINTER(c_int64), C.POINTER(c_int)]),
    "suta_launch_count": (c_int64, [c_void_p]),
    "suta_graph_replays": (c_int64, [c_void_p]),
    "suta_profile": (c_int, [c_void_p, c_int, C.POINTER(C.c_double), C.POINTER(c_int64), C.POINTER(C.c_double)]),
    "suta_profile_report": (C.c_char_p, [c_void_p]),
    "suta_debug_set_gemm_trace": (None, [c_void_p, c_int]),
    "suta_op_gemm": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_void_p,
                             c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "suta_op_gemm_mn": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_int,
                                c_void_p, c_int, c_void_p]),
    "suta_op_layernorm_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_int64, c_int, c_float, c_void_p]),
    "suta_op_layernorm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                      c_int, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int,
                                      c_void_p, c_void_p]),
    "suta_op_layernorm_bwd_scratch_floats": (c_int64, [c_int, c_int]),
    "suta_op_layernorm_fwd_mode": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p,
                                           c_void_p, c_void_p, c_int64, c_int, c_float, c_void_p, c_int, c_void_p]),
    "suta_op_layernorm_bwd_mode": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                           c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p,
                                           c_void_p, c_int, c_void_p, c_void_p]),
    "suta_op_attention_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int64, c_void_p]),
    "suta_op_attention_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                      c_int, c_int64, c_void_p]),
    "suta_op_loss": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_float, c_float, c_int, c_int, c_float, c_void_p,
                             c_void_p, c_void_p, c_void_p]),
    "suta_op_ctc_pseudo_label": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_void_p]),
    "suta_op_softmax_entropy": (c_int, [c_void_p, c_int64, c_float, c_void_p, c_void_p]),
    "suta_op_adam": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, C.POINTER(Hyper),
                             c_void_p, c_void_p]),
    "suta_op_decode": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
}

_lib = None


class SutaError(RuntimeError):
    pass


def load():
    """Load the CUDA extension; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SutaError(f"{LIB_PATH} not found: build it with `python __graft_entry__.py` "
                        f"(or test-time-adaptation-asr-suta_b200/build.py); suta_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.suta_abi_version() != ABI_VERSION:
        raise SutaError("libsuta_b200.so ABI version mismatch: rebuild")
    _lib = lib
    return lib


def check(status: int):
    if status != 0:
        raise SutaError(f"suta_b200 error {status}: {load().suta_last_error().decode(errors='replace')}")
