"""ctypes binding of libaqualora_b200.so (the C ABI in include/aqualora_b200.h).

The library is the product: there is no Python / CPU fallback.  If it is missing or a call fails, callers get
an exception that names the failing entry point and `aq_last_error()`.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p
from pathlib import Path

_PKG = Path(__file__).resolve().parent
# AQUALORA_B200_LIB: developer override (an instrumented build of the same sources, tools/gemm_trace.py)
LIB_PATH = Path(os.environ.get("AQUALORA_B200_LIB") or (_PKG / "libaqualora_b200.so"))


class AqualoraError(RuntimeError):
    pass


_SIGNATURES = {
    "aq_version": ([], c_int),
    "aq_arch": ([], c_int),
    "aq_last_error": ([], c_char_p),
    "aq_sm_count": ([], c_int),
    "aq_launch_count": ([], ctypes.c_longlong),
    "aq_lora_set_tuning": ([c_int, c_int], c_int),
    "aq_lora_linear_fwd": (
        [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64,
         c_int, c_int, c_int, c_void_p],
        c_int,
    ),
    "aq_lora_linear_fwd_residual": (
        [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64,
         c_int64, c_int, c_int, c_int, c_void_p],
        c_int,
    ),
    "aq_lora_linear_fwd_grouped": ([c_void_p, c_int64, c_void_p, c_int, c_void_p, c_int64, c_int64, c_int, c_int, c_void_p], c_int),
    "aq_lora_linear_bwd_workspace_bytes": ([c_int64, c_int], c_size_t),
    "aq_lora_linear_bwd": (
        [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
         c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p],
        c_int,
    ),
    "aq_lora_linear_bwd_dx": (
        [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_int,
         c_int, c_void_p, c_size_t, c_void_p],
        c_int,
    ),
    "aq_lora_wgrad_batch": ([c_void_p, c_int, c_void_p], c_int),
    "aq_wgrad_tn": ([c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_void_p], c_int),
    "aq_secret_encoder_workspace_bytes": ([c_int, c_int], c_size_t),
    "aq_secret_encoder_fwd": ([c_void_p] * 8 + [c_int] * 6 + [c_void_p, c_void_p], c_int),
    "aq_secret_encoder_bwd_workspace_bytes": ([c_int, c_int, c_int], c_size_t),
    "aq_secret_encoder_bwd": ([c_void_p] * 9 + [c_int] * 6 + [c_void_p, c_size_t, c_void_p], c_int),
    "aq_noise_jpeg_bwd": ([c_void_p, c_void_p, c_int, c_int, c_int, c_void_p], c_int),
    "aq_noise_crop_resize_bwd": ([c_void_p, c_void_p] + [c_int] * 11 + [c_void_p], c_int),
    "aq_noise_gauss_blur_bwd": ([c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p], c_int),
    "aq_noise_color_jiggle_bwd": ([c_void_p, c_void_p, c_void_p, c_void_p, ctypes.POINTER(c_int), c_int, c_int, c_int, c_void_p], c_int),
    "aq_prvl_workspace_bytes": ([c_int, c_int, c_int], c_size_t),
    "aq_prvl_loss_fwd": ([c_void_p] * 4 + [c_int] * 3 + [c_void_p, c_size_t, c_void_p], c_int),
    "aq_prvl_loss_bwd": ([c_void_p] * 6 + [c_int] * 3 + [c_void_p], c_int),
    "aq_bce_logits": ([c_void_p] * 4 + [c_int64, c_void_p], c_int),
    "aq_bn_train_workspace_bytes": ([c_int], c_size_t),
    "aq_bn_train_fwd": ([c_void_p] * 7 + [c_int64, c_int, c_float, c_float, c_int, c_void_p, c_size_t, c_void_p], c_int),
    "aq_bn_train_bwd": ([c_void_p] * 8 + [c_int64, c_int, c_int, c_void_p, c_size_t, c_void_p], c_int),
    "aq_dwconv_fwd": ([c_void_p] * 3 + [c_int] * 6 + [c_void_p], c_int),
    "aq_dwconv_bwd": ([c_void_p] * 5 + [c_int] * 6 + [c_void_p], c_int),
    "aq_conv1x1_wgrad": ([c_void_p] * 4 + [c_int64, c_int, c_int, c_int, c_void_p], c_int),
    "aq_stem_conv_fwd": ([c_void_p] * 3 + [c_int] * 3 + [c_void_p], c_int),
    "aq_stem_conv_bwd": ([c_void_p] * 5 + [c_int] * 3 + [c_void_p], c_int),
    "aq_mapper_fwd": ([c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p], c_int),
    "aq_mapper_bwd": ([c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p], c_int),
    "aq_cast_transpose_bf16": ([c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p], c_int),
    "aq_cast_transpose_bf16_batched": ([c_void_p, c_int, c_int64, c_void_p], c_int),
    "aq_transpose_bf16": ([c_void_p, c_void_p, c_int, c_int, c_void_p], c_int),
    "aq_noise_jpeg": ([c_void_p, c_void_p, c_int, c_int, c_int, c_void_p], c_int),
    "aq_noise_crop_resize": ([c_void_p, c_void_p] + [c_int] * 11 + [c_void_p], c_int),
    "aq_noise_gauss_blur": ([c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p], c_int),
    "aq_noise_gauss_noise": ([c_void_p, c_void_p, c_int64, c_float, ctypes.c_uint64, ctypes.c_uint64, c_void_p], c_int),
    "aq_noise_color_jiggle": ([c_void_p, c_void_p, c_void_p, ctypes.POINTER(c_int), c_int, c_int, c_int, c_void_p], c_int),
    "aq_effnetb1_packed_floats": ([c_int], c_size_t),
    "aq_effnetb1_workspace_bytes": ([c_int], c_size_t),
    "aq_effnetb1_fwd": ([c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p], c_int),
    "aq_conv1x1_tf32x3": ([c_void_p] * 7 + [c_int64, c_int, c_int, c_int, c_int, c_void_p], c_int),
    "aq_depthwise_silu": ([c_void_p] * 5 + [c_int] * 5 + [c_void_p], c_int),
    "aq_expand_dw_fused": ([c_void_p] * 8 + [c_int] * 6 + [c_void_p], c_int),
    "aq_lora_fold_down": ([c_void_p, c_void_p, c_void_p, c_int, c_int64, c_float, c_void_p], c_int),
    "aq_lora_merge": ([c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p], c_int),
    "aq_group_norm_workspace_bytes": ([c_int, c_int], c_size_t),
    "aq_group_norm_nhwc_fwd": ([c_void_p] * 6 + [c_int] * 4 + [c_float, c_int, c_void_p, c_size_t, c_void_p], c_int),
    "aq_group_norm_nhwc_bwd": ([c_void_p] * 7 + [c_int] * 4 + [c_float, c_int, c_void_p, c_size_t, c_void_p], c_int),
    "aq_geglu_fwd": ([c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p], c_int),
    "aq_geglu_bwd": ([c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int, c_void_p], c_int),
    "aq_layer_norm_fwd": ([c_void_p] * 5 + [c_int64, c_int, c_float, c_void_p], c_int),
    "aq_layer_norm_bwd": ([c_void_p] * 5 + [c_int64, c_int, c_void_p], c_int),
    "aq_add_bias_rows": ([c_void_p] * 4 + [c_int64, c_int, c_void_p], c_int),
    "aq_flat_sumsq": ([c_void_p, c_int64, c_void_p, c_void_p], c_int),
    "aq_flat_clip_adamw": (
        [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_float, c_float, c_float, c_float, c_float, c_float,
         c_float, c_int, c_void_p],
        c_int,
    ),
}



class LoraProjection(ctypes.Structure):
    """`aq_lora_projection` of include/aqualora_b200.h."""
    _fields_ = [("w", c_void_p), ("bias", c_void_p), ("down", c_void_p), ("up", c_void_p), ("y", c_void_p), ("ldy", c_int64),
                ("h_save", c_void_p), ("dout", c_int)]


class WgradJob(ctypes.Structure):
    """`aq_wgrad_job` of include/aqualora_b200.h."""
    _fields_ = [("gy", c_void_p), ("ldgy", c_int64), ("x", c_void_p), ("ldx", c_int64), ("ws", c_void_p), ("g_down", c_void_p),
                ("g_up", c_void_p), ("M", c_int64), ("din", c_int), ("dout", c_int), ("r", c_int)]


_lib = None


def exported_symbols() -> list[str]:
    """Every symbol include/aqualora_b200.h declares (kept in sync by tests/test_abi.py)."""
    return sorted(_SIGNATURES)


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise AqualoraError(
            f"{LIB_PATH} is missing: build it with `python -m aqualora_b200.build` (needs nvcc). "
            "aqualora_b200 has no CPU or PyTorch fallback for its kernels."
        )
    lib = ctypes.CDLL(os.fspath(LIB_PATH))
    for name, (argtypes, restype) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def last_error() -> str:
    return load().aq_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise AqualoraError(f"{what} failed with code {rc}: {last_error()}")


def call(name: str, *args) -> None:
    check(getattr(load(), name)(*args), name)
