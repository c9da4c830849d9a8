"""ORACLE — test infrastructure only; never imported by the product path.

CPU restatement of the reference's LCM sampler:
  * abar table            /root/reference/src/models/lcm_scheduler.py:76-100,116-129
  * inference schedule    lcm_scheduler.py:150-161, prev-timestep rule :169-174
  * one step              lcm_scheduler.py:204-242
  * the enhance loop      /root/reference/src/models/low_light_diffusion.py:204-240
The loop is *teacher-forced*: the initial latents and the per-step noises are
arguments, because CPU and CUDA generators produce different streams
and the reference ignores ``generator`` for step noise.  Feeding the tensors the
reference would have drawn reproduces ``LowLightDiffusion.enhance`` exactly
(checked against the unmodified reference by tests/golden/make_golden.py).

Parity status: PINNED (see unet_oracle.py header).  Importable only from tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from . import unet_oracle


def alphas_cumprod(num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                   rescale_zero_snr: bool = True) -> torch.Tensor:
    """scaled_linear schedule (:79-83), cumprod (:89-90), zero-SNR rescale (:116-129); fp32."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps) ** 2
    abar = torch.cumprod(1.0 - betas, dim=0)
    if rescale_zero_snr:
        s = abar.sqrt()
        s0, sT = s[0].clone(), s[-1].clone()
        s -= sT
        s *= s0 / (s0 - sT)
        abar = s ** 2
    return abar


def timesteps(num_inference_steps: int, num_train_timesteps: int = 1000, original_steps: int = 50) -> List[int]:
    c = num_train_timesteps // original_steps
    origin = torch.arange(1, original_steps + 1) * c - 1
    skip = len(origin) // num_inference_steps
    return origin[::skip][:num_inference_steps].flip(0).tolist()


def step(eps: torch.Tensor, t: int, sample: torch.Tensor, schedule: Sequence[int], abar: torch.Tensor,
         noise: Optional[torch.Tensor]):
    """Returns (prev_sample, pred_original_sample) — lcm_scheduler.py:204-242, epsilon prediction."""
    i = list(schedule).index(int(t))
    prev_t = schedule[i + 1] if i + 1 < len(schedule) else 0
    a_t = abar[t]
    a_prev = abar[prev_t] if prev_t > 0 else abar[0]          # final_alpha_cumprod = abar[0] (:100)
    x0 = (sample - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5
    if prev_t == 0:
        return x0, x0
    return a_prev ** 0.5 * x0 + (1 - a_prev) ** 0.5 * noise, x0


def condition_encoder(enc_sd: Dict[str, torch.Tensor], low_light: torch.Tensor) -> torch.Tensor:
    """condition_mode="add" (low_light_diffusion.py:108-113): Conv2d(3,32,3,p=1) -> SiLU -> Conv2d(32,3,3,p=1); keys
    `0.weight`, `0.bias`, `2.weight`, `2.bias` (the nn.Sequential's)."""
    import torch.nn.functional as F
    h = F.silu(F.conv2d(low_light, enc_sd["0.weight"], enc_sd["0.bias"], padding=1))
    return F.conv2d(h, enc_sd["2.weight"], enc_sd["2.bias"], padding=1)


@torch.no_grad()
def enhance(sd: Dict[str, torch.Tensor], cfg, low_light: torch.Tensor, latents0: torch.Tensor,
            noises: Sequence[torch.Tensor], num_inference_steps: int = 4, strict_groupnorm: bool = True,
            return_all: bool = False, condition_encoder_sd: Optional[Dict[str, torch.Tensor]] = None):
    """low_light_diffusion.py:204-240 with injected randomness.  `noises` has steps-1 entries.  With
    `condition_encoder_sd` the conditioning is "add" (:223-225): model_input = latents + condition_encoder(low_light)."""
    abar = alphas_cumprod()
    sched = timesteps(num_inference_steps)
    latents = latents0
    trace = []
    feat = condition_encoder(condition_encoder_sd, low_light) if condition_encoder_sd is not None else None
    for i, t in enumerate(sched):
        tt = torch.full((low_light.shape[0],), t, dtype=torch.long, device=low_light.device)
        x = torch.cat([latents, low_light], dim=1) if feat is None else latents + feat
        eps = unet_oracle.unet_forward(sd, cfg, x, tt, strict_groupnorm)
        nz = noises[i] if i < len(sched) - 1 else None
        latents, _ = step(eps, t, latents, sched, abar, nz)
        trace.append((eps, latents))
    out = latents.clamp(-1, 1)
    return (out, trace) if return_all else out
