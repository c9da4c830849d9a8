#!/usr/bin/env python
"""GPU diagnostic: run-to-run bitwise reproducibility of the big kernels at production size."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cv_diffusion_model_b200 import ops  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(3)


def same(a, b):
    return bool((a == b).all().item()) if a is not None else True


if os.environ.get("DW", "1") == "1":
    pass
def gemm_case(name, images, P, Ks, Nc, modes, h16, o16):
    M = images * P
    segs = []
    for K, mode, is16 in zip(Ks, modes, h16):
        a = torch.randn(M, K, device="cuda", generator=g).to(torch.float16 if is16 else torch.bfloat16)
        cf = torch.stack([torch.rand(images, K, device="cuda", generator=g) + 0.5,
                          torch.randn(images, K, device="cuda", generator=g) * (0.0 if mode == 4 else 0.3)], -1) if mode else None
        segs.append((a, cf, mode))
    wt = torch.randn(Nc, sum(Ks), device="cuda", generator=g) / sum(Ks) ** 0.5
    r = [ops.gemm(segs, wt, P, impl=1, out_f16=bool(o16)) for _ in range(3)]
    print(f"{name}: out", all(same(r[0][0], q[0]) for q in r), "stats", all(same(r[0][1], q[1]) for q in r),
          "max stats rel diff", max(((r[0][1] - q[1]).abs() / (r[0][1].abs() + 1e-9)).max().item() for q in r))


gemm_case("project0 f16", 64, 65536, [128, 32], 32, [4, 0], [1, 0], 0)
gemm_case("project0 bf16", 64, 65536, [128, 32], 32, [4, 0], [0, 0], 0)
gemm_case("project0 8 images", 8, 65536, [128, 32], 32, [4, 0], [1, 0], 0)
gemm_case("project_e1 f16", 64, 16384, [256, 64], 64, [4, 0], [1, 0], 0)
gemm_case("project_d3 f16", 32, 65536, [384, 64, 32], 32, [4, 0, 0], [1, 0, 0], 0)
gemm_case("project0 ungated", 64, 65536, [128, 32], 32, [0, 0], [1, 0], 0)
# activation-stationary expand kernel (gemm_wide.cu) and resident-weight expand kernel at production shapes
gemm_case("expand_d1 wide", 64, 4096, [256, 128], 1536, [2, 2], [0, 0], 1)
gemm_case("expand_d2 wide", 64, 16384, [128, 64], 768, [2, 2], [0, 0], 1)
gemm_case("expand_m wide", 64, 1024, [256], 1024, [2], [0], 1)
gemm_case("expand_l2 wide", 64, 4096, [128], 512, [2], [0], 1)
gemm_case("expand0 resident", 64, 65536, [32], 128, [2], [0], 1)


def conv_case(name, N, H, W, Ci, Co, mode):
    x = torch.randn(N, H, W, Ci, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Co, Ci, 3, 3, device="cuda", generator=g) / (9 * Ci) ** 0.5).bfloat16().float()
    b = torch.randn(Co, device="cuda", generator=g) * 0.1
    r = [ops.conv3x3(x, w, b, mode, impl=1) for _ in range(3)]
    print(f"{name}: out", all(same(r[0][0], q[0]) for q in r), "stats", all(same(r[0][1], q[1]) for q in r))


conv_case("up-conv halo 256x256 64ch", 16, 256, 256, 64, 32, 0)
conv_case("up-conv halo 128x128 128ch", 16, 128, 128, 128, 64, 0)
conv_case("down-conv 256x256 32ch", 16, 256, 256, 32, 32, 1)
