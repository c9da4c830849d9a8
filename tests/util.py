"""Shared helpers for the parity tests (weights, inputs and metrics)."""
import hashlib
import math

import torch

from cv_diffusion_model_b200.modules import create_efficient_unet


def randomise_affine(model, seed=1):
    """Same recipe as tests/golden/make_golden.py: GN gamma~U(0.5,1.5), all other 1-D params ~N(0,0.1^2)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.ndim == 1:
                if "norm" in name and name.endswith("weight") or name.endswith("to_out.1.weight"):
                    p.copy_(torch.rand(p.shape, generator=g) + 0.5)
                else:
                    p.copy_(torch.randn(p.shape, generator=g) * 0.1)


def seeded_unet(variant, image_size, patched=False, affine=False, seed=0, **overrides):
    torch.manual_seed(seed)
    m = create_efficient_unet(variant, image_size=image_size, in_channels=6, groupnorm="gcd" if patched else "strict", **overrides)
    if affine:
        randomise_affine(m)
    return m


def sd_digest(sd):
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def rel_rms(a, b):
    a, b = a.double(), b.double()
    return ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt().clamp_min(1e-30)).item()


def psnr(a, b, peak):
    mse = (a.double() - b.double()).pow(2).mean().item()
    return float("inf") if mse == 0 else 10 * math.log10(peak * peak / mse)
