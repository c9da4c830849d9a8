#!/usr/bin/env python
"""GPU diagnostic ("the GPU bar to beat"): the reference's op sequence as PyTorch eager on the same B200.

The reference itself does not travel to the GPU box; its restatement (oracle/, the same torch ops in the same order,
pinned against the reference's outputs) does.  fp32 with TF32 off, fp32 with TF32 on, and bf16 autocast.
    python tests/diag_eager_gpu.py [batch]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cv_diffusion_model_b200 import LowLightDiffusion  # noqa: E402
from oracle import lcm_oracle  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
pipe = LowLightDiffusion(unet_variant="small", image_size=256, num_inference_steps=4, precision="bf16")
cfg = pipe.unet.config
sd = {k[5:]: v.clone().cuda() for k, v in pipe.state_dict().items() if k.startswith("unet.")}
g = torch.Generator().manual_seed(1234)
low = (torch.rand(B, 3, 256, 256, generator=g) * 0.4 - 1).cuda()
lat = torch.randn(B, 3, 256, 256, generator=torch.Generator().manual_seed(9)).cuda()
noises = [torch.randn(B, 3, 256, 256).cuda() for _ in range(3)]


def run(tag, tf32, autocast):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    def once():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            return lcm_oracle.enhance(sd, cfg, low, lat, noises, 4)
    for _ in range(2):
        once()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 3
    a.record()
    for _ in range(n):
        once()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    print(f"{tag:28s} {ms:9.1f} ms per 4-step enhance of {B} images  = {B / ms * 1e3:8.1f} images/s")


run("eager fp32 (TF32 off)", False, False)
run("eager fp32 (TF32 on)", True, False)
run("eager bf16 autocast", True, True)
