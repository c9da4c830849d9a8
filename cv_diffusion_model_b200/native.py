"""ctypes binding of ``liblcmunet.so`` (C ABI: include/lcm_unet.h).

Loading fails loudly when the library is missing: the B200 path has no CPU or eager fallback.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liblcmunet.so")

LCM_MAX_LEVELS = 8
PREC_FP32, PREC_BF16 = 0, 1
ACT_F16 = 2   # single-kernel entry points: fp16 hidden tensors of the tensor-core plan
FLAG_SIMT_GEMM, FLAG_TAPS, FLAG_TRAIN, FLAG_DRY = 1, 2, 4, 8
ERR_INVALID, ERR_CUDA, ERR_UNKNOWN_WEIGHT, ERR_MISSING_WEIGHT, ERR_WORKSPACE = -1, -2, -3, -4, -5


class UNetConfigC(C.Structure):
    _fields_ = [
        ("in_channels", C.c_int32), ("out_channels", C.c_int32), ("base_channels", C.c_int32),
        ("num_levels", C.c_int32), ("channel_multipliers", C.c_int32 * LCM_MAX_LEVELS),
        ("num_attention_resolutions", C.c_int32), ("attention_resolutions", C.c_int32 * LCM_MAX_LEVELS),
        ("num_attention_heads", C.c_int32), ("num_res_blocks", C.c_int32), ("expansion_ratio", C.c_int32),
        ("se_ratio", C.c_float), ("time_embed_dim", C.c_int32), ("image_size", C.c_int32),
        ("groupnorm_gcd", C.c_int32), ("standard_attention", C.c_int32),
    ]


class GemmSegC(C.Structure):
    _fields_ = [("A", C.c_void_p), ("coef", C.c_void_p), ("K", C.c_int32), ("mode", C.c_int32), ("f16", C.c_int32),
                ("reserved", C.c_int32)]


class OpProfileC(C.Structure):
    _fields_ = [("name", C.c_char * 96), ("kernel", C.c_char * 32), ("ms", C.c_float), ("bytes", C.c_double),
                ("flops", C.c_double), ("ref_bytes", C.c_double)]


# name -> (restype, argtypes); also the list tests check against include/lcm_unet.h
SIGNATURES = {
    "lcm_last_error": (C.c_char_p, []),
    "lcm_version": (C.c_int, []),
    "lcm_plan_create": (C.c_int, [C.POINTER(UNetConfigC), C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int,
                                  C.POINTER(C.c_void_p)]),
    "lcm_plan_destroy": (None, [C.c_void_p]),
    "lcm_plan_workspace_bytes": (C.c_size_t, [C.c_void_p]),
    "lcm_plan_num_weights": (C.c_int, [C.c_void_p]),
    "lcm_plan_weight_info": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int64)]),
    "lcm_plan_set_weight": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "lcm_unet_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int, C.c_int64,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lcm_enhance": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int64),
                              C.POINTER(C.c_float), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lcm_condition_encode_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "lcm_condition_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                       C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "lcm_enhance_add": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int64),
                                  C.POINTER(C.c_float), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lcm_scheduler_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                     C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p]),
    "lcm_image_preprocess_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "lcm_image_postprocess_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "lcm_image_resize_u8": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "lcm_scheduler_mix": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64,
                                    C.c_int, C.c_void_p]),
    "lcm_train_num_backward_ops": (C.c_int, [C.c_void_p]),
    "lcm_train_backward_op_info": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p)]),
    "lcm_train_grad_elems": (C.c_int64, [C.c_void_p]),
    "lcm_train_grad_offset_bytes": (C.c_size_t, [C.c_void_p]),
    "lcm_train_grad_info": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                      C.POINTER(C.c_int)]),
    "lcm_train_loss": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "lcm_train_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int, C.c_int64, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                     C.c_void_p]),
    "lcm_plan_set_weights_flat": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "lcm_grad_sumsq": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "lcm_adamw_ema_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float,
                                     C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_float, C.c_void_p, C.c_float,
                                     C.c_float, C.c_void_p]),
    "lcm_train_read_grad_tap": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lcm_ddim_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p]),
    "lcm_consistency_loss": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_int, C.c_int64, C.c_void_p]),
    "lcm_plan_num_taps": (C.c_int, [C.c_void_p]),
    "lcm_plan_tap_info": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                    C.POINTER(C.c_int)]),
    "lcm_plan_read_tap": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lcm_plan_launches_per_forward": (C.c_int, [C.c_void_p]),
    "lcm_plan_algorithmic_bytes": (C.c_double, [C.c_void_p]),
    "lcm_plan_algorithmic_flops": (C.c_double, [C.c_void_p]),
    "lcm_plan_fused_bytes": (C.c_double, [C.c_void_p]),
    "lcm_plan_profile_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_int, C.c_int64,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(OpProfileC),
                                           C.c_int]),
    "lcm_debug_timeline": (C.c_int, [C.POINTER(C.c_longlong), C.c_int]),
    "lcm_op_gemm": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                              C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_void_p]),
    "lcm_op_conv3x3": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float),
                                 C.c_void_p]),
    "lcm_op_dwconv": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_void_p]),
    "lcm_op_xdw": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                             C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_void_p]),
}

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m cv_diffusion_model_b200.build` "
                "(nvcc, sm_100a). There is no CPU / eager fallback for this path.")
        _lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.restype = res
            fn.argtypes = args
    return _lib


def check(rc: int) -> int:
    if rc >= 0:
        return rc
    msg = (lib().lcm_last_error() or b"").decode()
    if rc == ERR_INVALID or rc == ERR_UNKNOWN_WEIGHT:
        raise ValueError(msg)
    raise RuntimeError(msg)


def config_struct(cfg, groupnorm: str) -> UNetConfigC:
    c = UNetConfigC()
    c.in_channels, c.out_channels, c.base_channels = cfg.in_channels, cfg.out_channels, cfg.base_channels
    mult = tuple(cfg.channel_multipliers)
    if len(mult) > LCM_MAX_LEVELS or len(cfg.attention_resolutions) > LCM_MAX_LEVELS:
        raise ValueError("too many levels / attention resolutions")
    c.num_levels = len(mult)
    for i, m in enumerate(mult):
        c.channel_multipliers[i] = m
    c.num_attention_resolutions = len(cfg.attention_resolutions)
    for i, r in enumerate(cfg.attention_resolutions):
        c.attention_resolutions[i] = r
    c.num_attention_heads, c.num_res_blocks = cfg.num_attention_heads, cfg.num_res_blocks
    c.expansion_ratio, c.se_ratio = cfg.expansion_ratio, cfg.se_ratio
    c.time_embed_dim, c.image_size = cfg.time_embed_dim, cfg.image_size
    c.groupnorm_gcd = 1 if groupnorm == "gcd" else 0
    c.standard_attention = 0 if getattr(cfg, "use_linear_attention", True) else 1
    return c
