#!/usr/bin/env python
"""ncu target: ONE native training step (forward + backward + optimizer) at Small@256, batch from argv (default 16)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cv_diffusion_model_b200 import LowLightDiffusion  # noqa: E402
from cv_diffusion_model_b200.training import NativeTrainer  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
torch.manual_seed(0)
pipe = LowLightDiffusion(unet_variant="small", image_size=256, precision="bf16").cuda().train()
tr = NativeTrainer(pipe, batch=B, precision="bf16")
g = torch.Generator().manual_seed(1)
high = (torch.rand(B, 3, 256, 256, generator=g) * 2 - 1).cuda()
low = ((high + 1) / 2) ** 3 * 2 - 1
print("loss", tr.train_step(low, high).item())
torch.cuda.synchronize()
