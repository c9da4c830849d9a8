#!/usr/bin/env python
"""GPU diagnostic: the fused expand -> depthwise path at production shapes (run under ncu for per-kernel figures).

    python tests/prof_xdw.py [l0|d30|l1]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cv_diffusion_model_b200 import ops  # noqa: E402

CASES = {"l0": (64, 256, 256, [32], 128), "d30": (64, 256, 256, [64, 32], 384), "l1": (64, 128, 128, [64], 256)}
N, H, W, Ks, Nc = CASES[sys.argv[1] if len(sys.argv) > 1 else "l0"]
g = torch.Generator(device="cuda").manual_seed(3)
xs = [torch.randn(N, H, W, K, device="cuda", generator=g).bfloat16() for K in Ks]
c1 = [torch.stack([torch.rand(N, K, device="cuda", generator=g) + 0.5, torch.randn(N, K, device="cuda", generator=g) * 0.5 + 1.0], -1) for K in Ks]
Kt = sum(Ks)
w = torch.randn(Nc, Kt, device="cuda", generator=g) / Kt ** 0.5
c2 = torch.stack([torch.rand(N, Nc, device="cuda", generator=g) + 0.5, torch.randn(N, Nc, device="cuda", generator=g)], -1)
wd = torch.randn(Nc, 1, 3, 3, device="cuda", generator=g) / 3
ops.xdw(list(zip(xs, c1)), w, c2, wd)
*_, ms = ops.xdw(list(zip(xs, c1)), w, c2, wd, repeat=5, timing=True)
print(f"{sys.argv[1] if len(sys.argv) > 1 else 'l0'}: {ms:.3f} ms per (xstats + finalise + fused + 3 memsets)")
