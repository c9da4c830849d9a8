#!/usr/bin/env python
"""GPU: one production-size GEMM shape, a few launches (target of `ncu -k regex:gemm_tc`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cv_diffusion_model_b200 import ops  # noqa: E402

CASES = {
    "expand0": (32, 65536, [32], 128, [2]),
    "project0": (32, 65536, [128, 32], 32, [1, 0]),
    "expand_d3": (32, 65536, [64, 32], 384, [2, 2]),
    "project_d3": (32, 65536, [384, 64, 32], 32, [1, 0, 0]),
}
images, P, Ks, Nc, modes = CASES[sys.argv[1] if len(sys.argv) > 1 else "expand0"]
g = torch.Generator(device="cuda").manual_seed(7)
M = images * P
segs = []
for K, mode in zip(Ks, modes):
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    coef = torch.stack([torch.rand(images, K, device="cuda", generator=g) + 0.5,
                        torch.randn(images, K, device="cuda", generator=g) * 0.3], -1) if mode else None
    segs.append((a, coef, mode))
w = torch.randn(Nc, sum(Ks), device="cuda", generator=g) / sum(Ks) ** 0.5
out, stats, ms = ops.gemm(segs, w, P, impl=1, repeat=3, timing=True)
print(sys.argv[1:], f"{ms*1e3:.1f} us/launch")
