#!/usr/bin/env python
"""GPU diagnostic: per-tile timeline of the tcgen05 GEMM pipeline roles (LCM_TC_DEBUG must include 64)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cv_diffusion_model_b200 import native, ops  # noqa: E402

CASES = {
    "expand0": (64, 65536, [32], 128, [2], [0], 1),
    "project0": (64, 65536, [128, 32], 32, [4, 0], [1, 0], 0),
    "expand_d1": (64, 4096, [256, 128], 1536, [2, 2], [0, 0], 1),
    "expand_d2": (32, 16384, [128, 64], 768, [2, 2], [0, 0], 1),
    "expand_m": (64, 1024, [256], 1024, [2], [0], 1),
    "expand_l2": (64, 4096, [128], 512, [2], [0], 1),
    "expand_d2f": (64, 16384, [128, 64], 768, [2, 2], [0, 0], 1),
    "project_m": (64, 1024, [1024, 256], 256, [4, 0], [1, 0], 0),
    "project_m128": (64, 1024, [1024, 256], 128, [4, 0], [1, 0], 0),
    "project_m_raw": (64, 1024, [1024, 256], 256, [0, 0], [1, 0], 0),
    "project0_raw": (64, 65536, [128, 32], 32, [0, 0], [1, 0], 0),
    "project0_n64": (64, 65536, [128, 32], 64, [0, 0], [1, 0], 0),
    "project0_k64": (64, 65536, [64, 32], 32, [0, 0], [1, 0], 0),
}
images, P, Ks, Nc, modes, h16, o16 = CASES[sys.argv[1] if len(sys.argv) > 1 else "expand0"]
images = int(os.environ.get("LCM_DIAG_IMAGES", images))
g = torch.Generator(device="cuda").manual_seed(7)
M = images * P
segs = []
for K, mode, is16 in zip(Ks, modes, h16):
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.float16 if is16 else torch.bfloat16)
    coef = torch.stack([torch.rand(images, K, device="cuda", generator=g) + 0.5,
                        torch.randn(images, K, device="cuda", generator=g) * (0.0 if mode == 4 else 0.3)], -1) if mode else None
    segs.append((a, coef, mode))
w = torch.randn(Nc, sum(Ks), device="cuda", generator=g) / sum(Ks) ** 0.5
ops.gemm(segs, w, P, impl=1, out_f16=bool(o16))
out, stats, ms = ops.gemm(segs, w, P, impl=1, repeat=1, timing=True, out_f16=bool(o16))
if os.environ.get("LCM_TIME_ONLY"):
    by = images * P * (sum(K * 2 for K in Ks) + Nc * 2)
    print(f"{ms*1e3:.1f} us  images {images}: {by / ms / 1e6:.0f} GB/s algorithmic")
    if int(os.environ.get("LCM_W_DEBUG", "0")) & 16:
        b8 = (C.c_longlong * 8)()
        native.lib().lcm_debug_timeline(b8, 8)
        n = max(1, b8[5])
        print("  MMA warp of block 0: wait_acc %d wait_tile %d wait_w %d issue %d total %d cycles over %d n-blocks (%.0f / n-block)" % (b8[0], b8[1], b8[2], b8[3], b8[4], b8[5], b8[4] / n))
    sys.exit(0)
buf = (C.c_longlong * 1024)()
native.lib().lcm_debug_timeline(buf, 1024)
t0 = min(buf[i] for i in range(16) if buf[i] > 0)
names = ["tma_empty", "xf_raw", "xf_arrive", "mma_tempty", "mma_xf", "mma_commit", "e1_tfull", "e1_sempty", "e1_done", "e2_sfull", "e2_done", "xf_cu", "xf_sts", "xf_fence"]
print(f"{ms*1e3:.1f} us;  cycles relative to first stamp")

print("tile " + " ".join(f"{n:>10s}" for n in names))
for it in range(0, 40):
    print(f"{it:4d} " + " ".join(f"{(buf[it*16+s]-t0) if buf[it*16+s] else -1:10d}" for s in range(len(names))))
