#!/usr/bin/env python
"""GPU diagnostic (not a test): tcgen05 GEMM / conv kernels against torch on a few shapes, with timings."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cv_diffusion_model_b200 import ops  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def one(images, P, Ks, Nc, modes, impl=1, repeat=1):
    g = torch.Generator(device="cuda").manual_seed(7)
    M = images * P
    segs, cols = [], []
    for K, mode in zip(Ks, modes):
        a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
        coef = None
        af = a.float().view(images, P, K)
        if mode:
            coef = torch.stack([torch.rand(images, K, device="cuda", generator=g) + 0.5,
                                torch.randn(images, K, device="cuda", generator=g) * 0.3], -1)
            af = af * coef[:, None, :, 0] + coef[:, None, :, 1]
            if mode == 2:
                af = af.clamp(0, 6)
        segs.append((a, coef, mode))
        cols.append(af.reshape(M, K).bfloat16().float())
    w = (torch.randn(Nc, sum(Ks), device="cuda", generator=g) / sum(Ks) ** 0.5).bfloat16().float()
    ref = torch.cat(cols, 1) @ w.t()
    if repeat > 1:
        ops.gemm(segs, w, P, impl=impl)   # warm-up (lazy module load, attribute set)
    out, stats, ms = ops.gemm(segs, w, P, impl=impl, repeat=repeat, timing=True)
    torch.cuda.synchronize()
    err = (out.float() - ref).abs().max().item()
    o = out.float().double().view(images, P, Nc)
    sref = torch.stack([o.sum(1), (o * o).sum(1)], -1)
    serr = (stats - sref).abs().max().item()
    byt = (sum(Ks) + Nc) * M * 2
    print(f"gemm impl={impl} img={images} P={P} K={Ks} N={Nc} modes={modes}: max|err|={err:.4f} (ref max {ref.abs().max().item():.2f}) "
          f"stats err={serr:.3e}  {ms*1e3:.1f} us  {byt/ms/1e6:.0f} GB/s {2*M*sum(Ks)*Nc/ms/1e9:.1f} TFLOP/s", flush=True)


def conv(N, H, W, C, mode, impl=1, repeat=1):
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(N, H, W, C, device="cuda", generator=g).bfloat16()
    w = (torch.randn(C, C, 3, 3, device="cuda", generator=g) / (9 * C) ** 0.5).bfloat16().float()
    b = torch.randn(C, device="cuda", generator=g) * 0.1
    xin = x.float().permute(0, 3, 1, 2)
    if mode == 2:
        xin = F.interpolate(xin, scale_factor=2, mode="bilinear", align_corners=False)
    ref = F.conv2d(xin, w, b, stride=2 if mode == 1 else 1, padding=1).permute(0, 2, 3, 1)
    if repeat > 1:
        ops.conv3x3(x, w, b, mode, impl=impl)
    out, stats, ms = ops.conv3x3(x, w, b, mode, impl=impl, repeat=repeat, timing=True)
    err = (out.float() - ref).abs().max().item()
    fl = 18.0 * ref.numel() * C
    print(f"conv impl={impl} N={N} {H}x{W} C={C} mode={mode}: max|err|={err:.4f} (ref max {ref.abs().max().item():.2f}) {ms*1e3:.1f} us "
          f"{fl/ms/1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    if which == "small":
        one(2, 128, [64], 128, [0])
        one(2, 128, [64], 128, [2])
        one(2, 256, [32], 128, [2])
        one(2, 256, [128, 32], 32, [1, 0])
        one(3, 64, [64, 32], 384, [2, 2])
        one(1, 16, [512], 2048, [2])
        one(2, 16, [2048, 512], 256, [1, 0])
        one(2, 100, [48], 192, [2])
        conv(2, 16, 16, 32, 0)
        conv(1, 8, 16, 64, 1)
        conv(2, 8, 8, 128, 2)
    else:
        for impl in (1,):
            one(64, 65536, [32], 128, [2], impl, 5)
            one(64, 65536, [128, 32], 32, [1, 0], impl, 5)
            one(64, 65536, [64, 32], 384, [2, 2], impl, 5)
            one(64, 65536, [384, 64, 32], 32, [1, 0, 0], impl, 5)
            one(64, 16384, [64], 256, [2], impl, 5)
            one(64, 16384, [256, 64], 64, [1, 0], impl, 5)
            one(64, 4096, [128], 512, [2], impl, 5)
            one(64, 1024, [256], 1024, [2], impl, 5)
            one(64, 1024, [512], 2048, [2], impl, 5)
            one(64, 1024, [2048, 512], 256, [1, 0], impl, 5)
            conv(64, 128, 128, 64, 0, impl, 3)
            conv(64, 32, 32, 256, 2, impl, 3)
            conv(64, 64, 64, 128, 2, impl, 3)
            conv(64, 256, 256, 32, 1, impl, 3)
