"""Python wrappers of the single-kernel C entry points (unit parity tests, micro-benchmarks).

Activations are NHWC tensors in the plan precision (torch.float32 or torch.bfloat16); weights are
given in the reference's fp32 layouts.  ``impl=1`` selects the tcgen05 / tuned kernel (bf16 only).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import native
from .engine import _stream_ptr

NONE, AFFINE, AFFINE_RELU6, AFFINE_SILU = 0, 1, 2, 3


def _prec(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return native.PREC_FP32
    if t.dtype == torch.bfloat16:
        return native.PREC_BF16
    if t.dtype == torch.float16:
        return native.ACT_F16
    raise ValueError(f"unsupported activation dtype {t.dtype}")


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def gemm(segs: Sequence[Tuple[torch.Tensor, Optional[torch.Tensor], int]], weight: torch.Tensor, pixels_per_image: int,
         impl: int = 0, want_stats: bool = True, repeat: int = 1, timing: bool = False, out_f16: bool = False):
    """segs: [(A [M,K] activations, coef [images,K,2] fp32 or None, mode)], weight fp32 [Nc, sum K].
    With impl=1 a segment may be torch.float16 (a block's hidden tensor) among bfloat16 ones, and out_f16 stores
    the result as fp16.  Returns (out [M,Nc], stats [images,Nc,2] float64 or None[, ms])."""
    a0 = next((a for a, _, _ in segs if a.dtype != torch.float16), segs[0][0])
    plan_dtype = torch.bfloat16 if a0.dtype == torch.float16 else a0.dtype
    M, Nc = a0.shape[0], weight.shape[0]
    images = M // pixels_per_image
    arr = (native.GemmSegC * len(segs))()
    keep = []
    for i, (a, coef, mode) in enumerate(segs):
        a = a.contiguous()
        keep.append(a)
        arr[i].A = a.data_ptr()
        if coef is not None:
            coef = coef.to(torch.float32).contiguous()
            keep.append(coef)
            arr[i].coef = coef.data_ptr()
        else:
            arr[i].coef = None
        arr[i].K = a.shape[1]
        arr[i].mode = mode
        arr[i].f16 = 1 if a.dtype == torch.float16 else 0
    w = weight.to(torch.float32).contiguous()
    out = torch.empty(M, Nc, dtype=torch.float16 if out_f16 else plan_dtype, device=a0.device)
    stats = torch.zeros(images, Nc, 2, dtype=torch.float64, device=a0.device) if want_stats else None
    ms = C.c_float(0)
    with torch.cuda.device(a0.device):
        prec = native.PREC_BF16 if plan_dtype == torch.bfloat16 else native.PREC_FP32
        native.check(native.lib().lcm_op_gemm(arr, len(segs), _p(w), _p(out), _p(stats), M, pixels_per_image, Nc,
                                              prec, impl | (0x100 if out_f16 else 0), repeat,
                                              C.byref(ms) if timing else None, _stream_ptr()))
    return (out, stats, ms.value) if timing else (out, stats)


def conv3x3(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, mode: int, impl: int = 0, repeat: int = 1,
            timing: bool = False):
    """x NHWC [N,H,W,Ci]; weight fp32 [Co,Ci,3,3]; mode 0 s1, 1 s2, 2 bilinear-x2 + s1."""
    n, h, w_, ci = x.shape
    co = weight.shape[0]
    ho, wo = (h // 2, w_ // 2) if mode == 1 else ((2 * h, 2 * w_) if mode == 2 else (h, w_))
    x = x.contiguous()
    wt, bs = weight.to(torch.float32).contiguous(), bias.to(torch.float32).contiguous()
    out = torch.empty(n, ho, wo, co, dtype=x.dtype, device=x.device)
    stats = torch.zeros(n, co, 2, dtype=torch.float64, device=x.device)
    ms = C.c_float(0)
    with torch.cuda.device(x.device):
        native.check(native.lib().lcm_op_conv3x3(_p(x), _p(wt), _p(bs), _p(out), _p(stats), n, h, w_, ci, co, mode,
                                                 _prec(x), impl, repeat, C.byref(ms) if timing else None, _stream_ptr()))
    return (out, stats, ms.value) if timing else (out, stats)


def dwconv(x: torch.Tensor, coef: torch.Tensor, weight: torch.Tensor, impl: int = 0, repeat: int = 1,
           timing: bool = False):
    """x NHWC [N,H,W,C]; coef fp32 [N,C,2]; weight fp32 [C,1,3,3].  Returns (out, pooled sums [N,C])."""
    n, h, w_, c = x.shape
    x, coef, wt = x.contiguous(), coef.to(torch.float32).contiguous(), weight.to(torch.float32).contiguous()
    out = torch.empty_like(x)
    pool = torch.zeros(n, c, dtype=torch.float64, device=x.device)
    ms = C.c_float(0)
    with torch.cuda.device(x.device):
        native.check(native.lib().lcm_op_dwconv(_p(x), _p(coef), _p(wt), _p(out), _p(pool), n, h, w_, c, _prec(x), impl,
                                                repeat, C.byref(ms) if timing else None, _stream_ptr()))
    return (out, pool, ms.value) if timing else (out, pool)


def xdw(segs: Sequence[Tuple[torch.Tensor, torch.Tensor]], weight: torch.Tensor, coef2: torch.Tensor, dw_weight: torch.Tensor,
        repeat: int = 1, timing: bool = False):
    """Fused expand -> GN2/FiLM/ReLU6 -> depthwise 3x3 + SE pool (csrc/xstats.cu + csrc/xdw_fused.cu).
    segs: [(x NHWC [N,H,W,K] bfloat16, coef1 [N,K,2] fp32)] (1-2 concat parts, GroupNorm1 + ReLU6 prologue);
    weight fp32 [Nc, sum K]; coef2 fp32 [N,Nc,2]; dw_weight fp32 [Nc,1,3,3].
    Returns (h2 fp16 [N,H,W,Nc], pool [N,Nc] f64, stats of the expand output [N,Nc,2] f64, t bf16 [N,H,W,sum K][, ms])."""
    x0 = segs[0][0]
    n, h, w_, _ = x0.shape
    nc = weight.shape[0]
    arr = (native.GemmSegC * len(segs))()
    keep = []
    kt = 0
    for i, (a, coef) in enumerate(segs):
        a = a.contiguous(); coef = coef.to(torch.float32).contiguous()
        keep += [a, coef]
        arr[i].A = a.data_ptr(); arr[i].coef = coef.data_ptr(); arr[i].K = a.shape[-1]; arr[i].mode = 2; arr[i].f16 = 0
        kt += a.shape[-1]
    wt = weight.to(torch.float32).contiguous()
    c2 = coef2.to(torch.float32).contiguous()
    dw = dw_weight.to(torch.float32).contiguous()
    t = torch.empty(n, h, w_, kt, dtype=torch.bfloat16, device=x0.device)
    out = torch.empty(n, h, w_, nc, dtype=torch.float16, device=x0.device)
    pool = torch.zeros(n, nc, dtype=torch.float64, device=x0.device)
    stats = torch.zeros(n, nc, 2, dtype=torch.float64, device=x0.device)
    ms = C.c_float(0)
    with torch.cuda.device(x0.device):
        native.check(native.lib().lcm_op_xdw(arr, len(segs), _p(wt), _p(c2), _p(dw), _p(t), _p(out), _p(pool), _p(stats),
                                             n, h, w_, nc, repeat, C.byref(ms) if timing else None, _stream_ptr()))
    return (out, pool, stats, t, ms.value) if timing else (out, pool, stats, t)


def image_preprocess_u8(images: torch.Tensor) -> torch.Tensor:
    """uint8 RGB ``[N,H,W,3]`` on the device -> fp32 ``[N,3,H,W]`` in [-1, 1]: ``x / 127.5 - 1``
    (the reference's ``preprocess_image`` after its resize, scripts/inference.py:111-116; bit-identical)."""
    if images.dtype != torch.uint8 or images.dim() != 4 or images.shape[-1] != 3 or not images.is_cuda:
        raise ValueError("expected a CUDA uint8 tensor [N, H, W, 3]")
    images = images.contiguous()
    n, h, w, _ = images.shape
    out = torch.empty(n, 3, h, w, dtype=torch.float32, device=images.device)
    native.check(native.lib().lcm_image_preprocess_u8(_p(images), _p(out), n, h, w, _stream_ptr()))
    return out


def image_postprocess_u8(images: torch.Tensor) -> torch.Tensor:
    """fp32 ``[N,3,H,W]`` on the device -> uint8 RGB ``[N,H,W,3]``: ``clip((y + 1) * 127.5, 0, 255)`` truncated
    (the reference's ``postprocess_image`` before its resize, scripts/inference.py:121-127; bit-identical)."""
    if images.dtype != torch.float32 or images.dim() != 4 or images.shape[1] != 3 or not images.is_cuda:
        raise ValueError("expected a CUDA fp32 tensor [N, 3, H, W]")
    images = images.contiguous()
    n, _, h, w = images.shape
    out = torch.empty(n, h, w, 3, dtype=torch.uint8, device=images.device)
    native.check(native.lib().lcm_image_postprocess_u8(_p(images), _p(out), n, h, w, _stream_ptr()))
    return out


def image_resize_u8(images: torch.Tensor, height: int, width: int) -> torch.Tensor:
    """``cv2.resize(img, (width, height))`` (default INTER_LINEAR) on uint8 RGB ``[N,H,W,3]`` device tensors,
    bit-identical to OpenCV's fixed-point bilinear (scripts/inference.py:109,130)."""
    if images.dtype != torch.uint8 or images.dim() != 4 or images.shape[-1] != 3 or not images.is_cuda:
        raise ValueError("expected a CUDA uint8 tensor [N, H, W, 3]")
    images = images.contiguous()
    n, h, w, _ = images.shape
    out = torch.empty(n, int(height), int(width), 3, dtype=torch.uint8, device=images.device)
    native.check(native.lib().lcm_image_resize_u8(_p(images), n, h, w, _p(out), int(height), int(width), _stream_ptr()))
    return out
