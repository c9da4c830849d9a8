#!/usr/bin/env python
"""GPU diagnostic: per-tile timeline of the expand GEMM roles (run with LCM_X_TIMELINE=1).

    LCM_X_TIMELINE=1 python tests/diag_timeline_expand.py [expand0|expand_d3]
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cv_diffusion_model_b200 import native, ops  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "expand0"
images, P = 32, 65536
Ks, Nc = ([32], 128) if case == "expand0" else ([64, 32], 384)
g = torch.Generator(device="cuda").manual_seed(7)
M = images * P
segs = []
for K in Ks:
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    coef = torch.stack([torch.rand(images, K, device="cuda", generator=g) + 0.5, torch.randn(images, K, device="cuda", generator=g) * 0.3], -1)
    segs.append((a, coef, 2))
w = torch.randn(Nc, sum(Ks), device="cuda", generator=g) / sum(Ks) ** 0.5
ops.gemm(segs, w, P, impl=1, out_f16=True)
out, stats, ms = ops.gemm(segs, w, P, impl=1, repeat=1, timing=True, out_f16=True)
buf = (C.c_longlong * 1024)()
native.lib().lcm_debug_timeline(buf, -1024)
t0 = min(buf[i] for i in range(16) if buf[i] > 0)
names = ["tma_issue", "xf_raw", "xf_done", "mma_xf", "mma_done", "e_tfull0", "e_ld0", "e_st0", "mma_j0", "mma_j1", "e_tfull1", "e_ld1", "e_st1", "gram_start", "gram_done"]
print(f"{case}: {ms*1e3:.1f} us;  cycles relative to first stamp")
print("tile " + " ".join(f"{n:>9s}" for n in names))
for it in range(0, 48):
    print(f"{it:4d} " + " ".join(f"{(buf[it*16+s]-t0) if buf[it*16+s] else -1:9d}" for s in range(len(names))))
