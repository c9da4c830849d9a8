#!/usr/bin/env python
"""GPU: one production-size depthwise shape, a few launches (target of `ncu -k regex:dwconv`).

    python tests/prof_dwconv.py [images] [size] [channels] [repeat]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cv_diffusion_model_b200 import ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
s = int(sys.argv[2]) if len(sys.argv) > 2 else 256
c = int(sys.argv[3]) if len(sys.argv) > 3 else 128
rep = int(sys.argv[4]) if len(sys.argv) > 4 else 5
dt = torch.float16 if (sys.argv[5] if len(sys.argv) > 5 else "f16") == "f16" else torch.bfloat16
g = torch.Generator(device="cuda").manual_seed(7)
x = torch.randn(n, s, s, c, device="cuda", generator=g).to(dt)
coef = torch.stack([torch.rand(n, c, device="cuda", generator=g) + 0.5, torch.randn(n, c, device="cuda", generator=g) * 0.3 + 1.0], -1)
w = torch.randn(c, 1, 3, 3, device="cuda", generator=g) / 3
ops.dwconv(x, coef, w, impl=1, repeat=2)   # warm-up: module load, tensor-map encode
out, pool, ms = ops.dwconv(x, coef, w, impl=1, repeat=rep, timing=True)
by = 2 * x.numel() * 2
print(f"dwconv {dt} N={n} {s}x{s} C={c}: {ms*1e3:.1f} us/launch, {by/ms/1e6:.0f} GB/s algorithmic")
