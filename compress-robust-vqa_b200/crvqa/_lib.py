"""Loader and prototypes for the C ABI declared in include/crvqa.h."""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_longlong, c_size_t, c_void_p

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libcrvqa.so")


class CrvqaError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise CrvqaError(
            f"{LIB_PATH} not found: build it with compress-robust-vqa_b200/csrc/build.sh "
            "(or __graft_entry__.build()); there is no CPU fallback for the masked kernels")
    return ctypes.CDLL(LIB_PATH)


lib = _load()

_P = c_void_p
_PROTOS = {
    "crv_version": (c_int, []),
    "crv_error_string": (c_char_p, [c_int]),
    "crv_last_cuda_error": (c_int, []),
    "crv_launch_count": (ctypes.c_ulonglong, []),
    "crv_cast_f32_to_bf16": (c_int, [_P, _P, c_int64, _P]),
    "crv_binarize": (c_int, [_P, _P, _P, _P, _P, c_int64, _P]),
    "crv_apply_mask_bf16": (c_int, [_P, _P, _P, _P, c_int64, _P]),
    "crv_apply_mask_segmented": (c_int, [_P, _P, _P, _P, c_int, _P, _P]),
    "crv_masked_linear_fwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "crv_gemm_debug_timestamps": (c_int, [_P]),
    "crv_masked_linear_bwd_dx": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "crv_masked_linear_bwd_ds": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "crv_masked_gemm_grouped": (c_int, [_P, c_int, _P]),
    "crv_masked_linear_small_k_fwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P]),
    "crv_masked_linear_small_k_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "crv_masked_embedding_fwd": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int64, c_int, _P]),
    "crv_masked_embedding_bwd": (c_int, [_P, _P, _P, _P, c_int64, c_int64, c_int, c_longlong, _P]),
    "crv_mul_cast_bf16": (c_int, [_P, _P, _P, c_int64, _P]),
    "crv_kth_value_workspace_bytes": (c_size_t, [c_int]),
    "crv_kth_value_workspace_bytes_for": (c_size_t, [_P, c_int]),
    "crv_kth_value_batched": (c_int, [_P, _P, _P, c_int, c_int, _P, _P, c_size_t, _P]),
    "crv_magnitude_init": (c_int, [_P, _P, c_float, c_float, _P, c_int64, _P]),
    "crv_vqa_loss_workspace_bytes": (c_size_t, [c_int]),
    "crv_vqa_loss_bce": (c_int, [_P, _P, _P, _P, c_int, c_int, _P, _P]),
    "crv_vqa_loss_lpf": (c_int, [_P, _P, _P, c_float, _P, _P, _P, c_int, c_int, _P, _P]),
    "crv_vqa_loss_rubi": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, _P, _P]),
    "crv_vqa_loss_lmh": (c_int, [_P, _P, _P, _P, c_float, c_float, _P, _P, _P, c_int, c_int, _P, _P]),
    "crv_ln_fwd": (c_int, [_P, c_int, _P, _P, _P, c_float, c_float, _P, c_int, _P, _P, _P, _P, c_int, c_int, _P]),
    "crv_ln_bwd_partials_bytes": (c_size_t, [c_int, c_int]),
    "crv_ln_bwd": (c_int, [_P, _P, _P, c_int, _P, _P, _P, _P, c_float, _P, c_int, _P, c_int, _P, _P, c_int, c_int, _P]),
    "crv_partial_reduce": (c_int, [_P, c_int, c_int, _P, c_int, _P]),
    "crv_colsum_workspace_bytes": (c_size_t, [c_int]),
    "crv_colsum_bf16": (c_int, [_P, c_int, c_int, _P, c_int, _P, _P]),
    "crv_attention_fwd": (c_int, [_P, c_longlong, c_longlong, _P, c_longlong, c_longlong, _P, c_longlong, c_longlong, _P,
                                  _P, c_int, c_int, c_int, c_int, c_float, c_float, _P, c_int, _P]),
    "crv_ln_avg_drop_fwd": (c_int, [_P, _P, _P, _P, _P, _P, c_float, c_float, _P, c_int, _P, _P, _P, c_int, c_int, _P]),
    "crv_ln_avg_drop_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_float, _P, c_int, _P, _P, c_int, c_int, _P]),
    "crv_attention_probs_pitch": (c_int, [c_int]),
    "crv_attention_fwd_p": (c_int, [_P, c_longlong, c_longlong, _P, c_longlong, c_longlong, _P, c_longlong, c_longlong, _P,
                                    _P, _P, c_int, c_int, c_int, c_int, c_float, c_float, _P, c_int, _P]),
    "crv_attention_bwd_p": (c_int, [_P, c_longlong, c_longlong, _P, c_longlong, c_longlong, _P, c_longlong, c_longlong, _P,
                                    _P, _P, c_longlong, c_longlong, _P, c_longlong, c_longlong, _P, c_longlong, c_longlong,
                                    c_int, c_int, c_int, c_int, c_float, c_float, _P]),
    "crv_attention_bwd": (c_int, [_P, c_longlong, c_longlong, _P, c_longlong, c_longlong, _P, c_longlong, c_longlong, _P,
                                  _P, _P, c_longlong, c_longlong, _P, c_longlong, c_longlong, _P, c_longlong, c_longlong,
                                  c_int, c_int, c_int, c_int, c_float, c_float, _P, c_int, _P]),
    "crv_gelu_fwd": (c_int, [_P, _P, c_int64, _P]),
    "crv_gelu_bwd": (c_int, [_P, _P, _P, c_int64, _P]),
    "crv_momentum_update": (c_int, [_P, _P, _P, c_int, c_float, c_float, _P]),
    "crv_fq_attention_fwd": (c_int, [_P, _P, _P, _P, c_longlong, c_longlong, _P, _P, c_int, c_int, c_int, c_int, c_float,
                                     c_float, _P, c_int, _P]),
    "crv_fq_attention_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_float, c_float, _P]),
    "crv_sumsq_multi": (c_int, [_P, _P, c_int, _P, _P, _P]),
    "crv_adamw_multi": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int, c_double, c_int, c_double, c_double, c_double,
                                c_double, _P, c_float, _P]),
    "crv_quick_gelu_fwd": (c_int, [_P, _P, c_int64, _P]),
    "crv_quick_gelu_bwd": (c_int, [_P, _P, _P, c_int64, _P]),
    "crv_rng_advance": (c_int, [_P, _P]),
    "crv_sumsq_workspace_bytes": (c_size_t, []),
    "crv_sumsq": (c_int, [_P, c_int64, _P, _P, _P]),
    "crv_adamw_step": (c_int, [_P, _P, _P, _P, _P, c_int64, c_float, c_float, c_double, c_double, c_float,
                               c_float, _P, c_float, _P, c_int, c_float, _P]),
    "crv_sumsq_segmented": (c_int, [_P, _P, c_int, _P, _P, _P]),
    "crv_adamw_segmented": (c_int, [_P, _P, _P, _P, _P, _P, c_int, _P, _P, _P, c_float, c_float, c_double, c_double, c_float,
                                    c_float, _P, c_float, _P, c_int, c_int, c_float, _P]),
}
EXPORTED = tuple(_PROTOS)
for _name, (_res, _args) in _PROTOS.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args


class GemmProblem(ctypes.Structure):
    """crv_gemm_problem of include/crvqa.h."""
    _fields_ = [("kind", c_int), ("act", c_int), ("out_dtype", c_int), ("accumulate", c_int),
                ("M", c_int), ("N", c_int), ("K", c_int), ("reserved", c_int),
                ("a", c_void_p), ("b", c_void_p), ("bias", c_void_p), ("w_f32", c_void_p),
                ("out", c_void_p), ("aux", c_void_p)]


def check(rc, what=""):
    """Raise CrvqaError for a non-zero return code of a crv_* call."""
    if rc != 0:
        msg = lib.crv_error_string(int(rc))
        msg = msg.decode() if msg else "?"
        raise CrvqaError(f"{what or 'crvqa call'} failed: rc={rc} ({msg})")
