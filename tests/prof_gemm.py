#!/usr/bin/env python
"""GPU: one production-size GEMM shape, a few launches (target of `ncu -k regex:gemm_tc`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cv_diffusion_model_b200 import ops  # noqa: E402

# (images, pixels per image, [K per segment], Nc, [prologue mode per segment], [segment is fp16], fp16 output)
CASES = {
    "expand0": (32, 65536, [32], 128, [2], [0], 1),
    "project0": (32, 65536, [128, 32], 32, [4, 0], [1, 0], 0),
    "expand_d3": (32, 65536, [64, 32], 384, [2, 2], [0, 0], 1),
    "project_d3": (32, 65536, [384, 64, 32], 32, [4, 0, 0], [1, 0, 0], 0),
    "expand_d2": (32, 16384, [128, 64], 768, [2, 2], [0, 0], 1),
    "project_d2": (32, 16384, [768, 128, 64], 64, [4, 0, 0], [1, 0, 0], 0),
    "expand_e1": (32, 16384, [64], 256, [2], [0], 1),
    "expand_m": (64, 1024, [256], 1024, [2], [0], 1),
    "expand_d1": (64, 4096, [256, 128], 1536, [2, 2], [0, 0], 1),
    "expand_d0": (64, 1024, [256, 256], 2048, [2, 2], [0, 0], 1),
    "project_d1": (64, 4096, [1536, 256, 128], 128, [4, 0, 0], [1, 0, 0], 0),
    "project_d0": (64, 1024, [2048, 256, 256], 256, [4, 0, 0], [1, 0, 0], 0),
    "expand_e2": (64, 4096, [128], 512, [2], [0], 1),
    "project_e2": (64, 4096, [512, 128], 128, [4, 0], [1, 0], 0),
    "project_m": (64, 1024, [1024, 256], 256, [4, 0], [1, 0], 0),
}
images, P, Ks, Nc, modes, h16, o16 = CASES[sys.argv[1] if len(sys.argv) > 1 else "expand0"]
g = torch.Generator(device="cuda").manual_seed(7)
M = images * P
segs = []
for K, mode, is16 in zip(Ks, modes, h16):
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.float16 if is16 else torch.bfloat16)
    coef = torch.stack([torch.rand(images, K, device="cuda", generator=g) + 0.5,
                        torch.randn(images, K, device="cuda", generator=g) * (0.0 if mode == 4 else 0.3)], -1) if mode else None
    segs.append((a, coef, mode))
w = torch.randn(Nc, sum(Ks), device="cuda", generator=g) / sum(Ks) ** 0.5
ops.gemm(segs, w, P, impl=1, repeat=2, out_f16=bool(o16))   # warm-up: module load, tensor maps
out, stats, ms = ops.gemm(segs, w, P, impl=1, repeat=int(sys.argv[2]) if len(sys.argv) > 2 else 5, timing=True, out_f16=bool(o16))
by = (sum(Ks) + Nc) * M * 2 + sum(Ks) * Nc * 2
print(sys.argv[1:], f"{ms*1e3:.1f} us/launch  {by/ms/1e6:.0f} GB/s  {2*M*sum(Ks)*Nc/ms/1e9:.0f} TFLOP/s")
