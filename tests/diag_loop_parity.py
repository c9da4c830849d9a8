#!/usr/bin/env python
"""GPU diagnostic (not a test): where does the free-running bf16 loop lose PSNR against the fp32 oracle?

For each case: the 4/8-step loop in bf16 through (a) the tcgen05 plan (fp16 hidden tensors, packed-fp16 edge convs),
(b) the CUDA-core bf16 plan (one storage type), (c) fp32 plan, next to (d) the oracle under torch's CPU bf16 autocast.
Per step: teacher-forced eps rel-RMS and free-running pre-clamp PSNR (peak 2).

    python tests/diag_loop_parity.py > gpurun_out/loop_parity.txt
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cv_diffusion_model_b200 import LowLightDiffusion  # noqa: E402
from cv_diffusion_model_b200.engine import Engine  # noqa: E402
from oracle import lcm_oracle, unet_oracle  # noqa: E402
from tests.util import psnr, randomise_affine, rel_rms  # noqa: E402


def case(tag, size, b, steps, affine, low_scale):
    torch.manual_seed(0)
    pipe = LowLightDiffusion(unet_variant="small", image_size=size, num_inference_steps=steps, precision="bf16")
    if affine:
        randomise_affine(pipe.unet)
    sd = {k[5:]: v.clone() for k, v in pipe.state_dict().items()}
    cfg = pipe.unet.config
    low = torch.rand(b, 3, size, size, generator=torch.Generator().manual_seed(1234)) * low_scale - 1
    lat0 = torch.randn(b, 3, size, size, generator=torch.Generator().manual_seed(9))
    torch.manual_seed(5)
    noises = torch.stack([torch.randn(b, 3, size, size) for _ in range(steps - 1)])
    _, trace = lcm_oracle.enhance(sd, cfg, low, lat0, list(noises), steps, return_all=True)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        _, tr_ac = lcm_oracle.enhance(sd, cfg, low, lat0, list(noises), steps, return_all=True)
    sched = lcm_oracle.timesteps(steps)
    print(f"\n## {tag}: small@{size} b={b} steps={steps} affine={affine}")
    print("free-running pre-clamp PSNR per step (dB):")
    print(f"  {'torch cpu autocast bf16':28s}", " ".join(f"{psnr(tr_ac[i][1].float(), trace[i][1], 2.0):6.1f}" for i in range(steps)))
    pipe = pipe.cuda().eval()
    for label, prec, simt in (("native bf16 tcgen05", "bf16", False), ("native bf16 cuda-core", "bf16", True),
                              ("native fp32", "fp32", False)):
        eng = Engine(pipe.unet, b, size, size, precision=prec, simt_gemm=simt, device="cuda")
        coefs = []
        pipe.scheduler.set_timesteps(steps, device="cuda")
        coefs = [pipe.scheduler.step_coefficients(t) for t in sched]
        lat = lat0.cuda().clone()
        _, tr = eng.enhance(low.cuda(), lat, noises.cuda(), sched, coefs, trace=True)
        print(f"  {label:28s}", " ".join(f"{psnr(tr[i].cpu(), trace[i][1], 2.0):6.1f}" for i in range(steps)))
        # teacher-forced eps error per step
        errs = []
        for i, t in enumerate(sched):
            x = torch.cat([lat0 if i == 0 else trace[i - 1][1], low], dim=1)
            tt = torch.full((b,), t, dtype=torch.long)
            eps = eng.forward(x.cuda(), tt.cuda()).cpu()
            errs.append(rel_rms(eps, trace[i][0]))
        print(f"  {'  teacher-forced eps relRMS':28s}", " ".join(f"{e:6.4f}" for e in errs))
        eng.close()
    errs = []
    for i, t in enumerate(sched):
        x = torch.cat([lat0 if i == 0 else trace[i - 1][1], low], dim=1)
        tt = torch.full((b,), t, dtype=torch.long)
        with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
            e = unet_oracle.unet_forward(sd, cfg, x, tt).float()
        errs.append(rel_rms(e, trace[i][0]))
    print(f"  {'autocast teacher-forced eps':28s}", " ".join(f"{e:6.4f}" for e in errs))


def layers(size, b, affine, low_scale):
    """per-layer error at step 0 of a case, tcgen05 vs CUDA-core bf16"""
    torch.manual_seed(0)
    pipe = LowLightDiffusion(unet_variant="small", image_size=size, num_inference_steps=4, precision="bf16")
    if affine:
        randomise_affine(pipe.unet)
    sd = {k[5:]: v.clone() for k, v in pipe.state_dict().items()}
    low = torch.rand(b, 3, size, size, generator=torch.Generator().manual_seed(1234)) * low_scale - 1
    lat0 = torch.randn(b, 3, size, size, generator=torch.Generator().manual_seed(9))
    x = torch.cat([lat0, low], dim=1)
    t = torch.full((b,), 739, dtype=torch.long)
    taps = {}
    with torch.no_grad():
        want = unet_oracle.unet_forward(sd, pipe.unet.config, x, t, tap=lambda k, v: taps.__setitem__(k.rstrip("."), v))
    cols = {}
    names = None
    for label, simt in (("tc", False), ("simt", True)):
        eng = Engine(pipe.unet.cuda(), b, size, size, precision="bf16", simt_gemm=simt, taps=True, device="cuda")
        y = eng.forward(x.cuda(), t.cuda()).cpu()
        names = [n for n in eng.taps() if n in taps]
        cols[label] = {n: rel_rms(eng.read_tap(n).cpu(), taps[n]) for n in names}
        cols[label]["eps"] = rel_rms(y, want)
        eng.close()
    print(f"\n## per-layer rel-RMS at step 0 (t=739), small@{size} b={b} affine={affine}")
    print(f"{'tap':44s} {'tcgen05':>10s} {'cuda-core':>10s}")
    for n in names + ["eps"]:
        print(f"{n:44s} {cols['tc'][n]:10.3e} {cols['simt'][n]:10.3e}")


if __name__ == "__main__":
    case("smoke case", 64, 2, 4, False, 0.4)
    case("golden small64 recipe", 64, 2, 4, True, 0.4)
    case("8-step", 64, 2, 8, True, 0.4)
    case("small128 default init", 128, 2, 4, False, 0.4)
    layers(64, 2, False, 0.4)
