#!/usr/bin/env python
"""GPU diagnostic: does capturing one `enhance` call in a CUDA graph shorten it? (inter-kernel launch gaps)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cv_diffusion_model_b200 import LowLightDiffusion  # noqa: E402

B, S = 64, 256
torch.manual_seed(0)
pipe = LowLightDiffusion(unet_variant="small", image_size=S, num_inference_steps=4, precision="bf16").cuda().eval()
low = torch.rand(B, 3, S, S, device="cuda") * 0.4 - 1
lat0 = torch.randn(B, 3, S, S, device="cuda")
noises = torch.randn(3, B, 3, S, S, device="cuda")


from cv_diffusion_model_b200.engine import get_engine  # noqa: E402
eng = get_engine(pipe.unet, B, S, S, low.device)
pipe.scheduler.set_timesteps(4, device=low.device)
ts = list(pipe.scheduler._host_timesteps)
coefs = [pipe.scheduler.step_coefficients(t) for t in ts]
lat = lat0.clone()


def run():   # the native call only (what a serving loop would capture)
    lat.copy_(lat0)
    return eng.enhance(low, lat, noises, ts, coefs)


def timeit(f, n=5):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        f()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for _ in range(3):
    ref = run()
print("stream launches: %.2f ms per enhance" % timeit(run))
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    run()
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    out = run()
g.replay()
torch.cuda.synchronize()
print("graph replay:    %.2f ms per enhance" % timeit(g.replay))
o1 = out.clone()
g.replay(); torch.cuda.synchronize()
print("max |graph - stream| =", (out - ref).abs().max().item(), " graph run-to-run", (out - o1).abs().max().item())
ref2 = run(); torch.cuda.synchronize()
print("stream run-to-run", (ref2 - ref).abs().max().item(), " frac differing px graph vs stream", ((out - ref).abs() > 1e-3).float().mean().item())
