"""ORACLE — test infrastructure only; never imported by the product path.

CPU/GPU restatement of the reference's TRAINING step for the hot path, built on the pinned forward oracle
(unet_oracle.py) and torch autograd:
  * loss      LowLightDiffusion.forward / compute_loss   /root/reference/src/models/low_light_diffusion.py:115-175,250-277
              add_noise                                   /root/reference/src/models/lcm_scheduler.py:255-280
  * step      LowLightTrainer.train_epoch (non-AMP)       /root/reference/src/training/trainer.py:303-322
              EMAModel.update                             trainer.py:98-104
Parity status: PINNED — tests/golden/make_golden_train.py checks `loss_and_grads` against the unmodified reference's
loss.backward() (same ATen ops, gradients agree to 1e-4 relative) and stores the reference's own step results in
tests/golden/train_kat.npz.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import unet_oracle

LOSSES = {"mse": F.mse_loss, "huber": F.huber_loss, "l1": F.l1_loss}


def add_noise(x0: torch.Tensor, noise: torch.Tensor, t: torch.Tensor, abar: torch.Tensor) -> torch.Tensor:
    """lcm_scheduler.py:255-280: sqrt(abar_t) x0 + sqrt(1 - abar_t) noise, per-sample t."""
    a = abar.to(x0.device)[t].view(-1, 1, 1, 1)
    return a ** 0.5 * x0 + (1 - a) ** 0.5 * noise


def loss_fn(sd: Dict[str, torch.Tensor], cfg, abar, low, high, t, noise, loss_type: str = "mse", strict_groupnorm: bool = True,
            tap=None):
    noisy = add_noise(high, noise, t, abar)
    eps = unet_oracle.unet_forward(sd, cfg, torch.cat([noisy, low], dim=1), t, strict_groupnorm, tap=tap)
    return LOSSES[loss_type](eps, noise), eps


def loss_and_grads(sd, cfg, abar, low, high, t, noise, loss_type: str = "mse", strict_groupnorm: bool = True):
    """sd: tensors with requires_grad=True.  Returns (loss value, {name: gradient})."""
    loss, _ = loss_fn(sd, cfg, abar, low, high, t, noise, loss_type, strict_groupnorm)
    names = [k for k, v in sd.items() if v.requires_grad]
    grads = torch.autograd.grad(loss, [sd[k] for k in names])
    return loss.item(), dict(zip(names, grads))


def train_steps(sd0: Dict[str, torch.Tensor], cfg, abar, batches, lr=1e-4, weight_decay=0.01, betas=(0.9, 0.999), eps=1e-8,
                max_norm: Optional[float] = 1.0, ema_decay: Optional[float] = 0.9999, loss_type="mse", strict_groupnorm=True):
    """The reference loop on plain tensors: returns (losses, grad norms, final weights, EMA shadow)."""
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd0.items()}
    opt = torch.optim.AdamW(list(params.values()), lr=lr, weight_decay=weight_decay, betas=betas, eps=eps)
    shadow = {k: v.detach().clone() for k, v in params.items()} if ema_decay is not None else None
    losses, norms = [], []
    for low, high, t, noise in batches:
        opt.zero_grad()
        loss, _ = loss_fn(params, cfg, abar, low, high, t, noise, loss_type, strict_groupnorm)
        loss.backward()
        gn = torch.nn.utils.clip_grad_norm_(list(params.values()), max_norm) if max_norm else torch.zeros(())
        opt.step()
        if shadow is not None:
            with torch.no_grad():
                for k in shadow:
                    shadow[k].mul_(ema_decay).add_(params[k].detach(), alpha=1 - ema_decay)
        losses.append(loss.item())
        norms.append(float(gn))
    return losses, norms, {k: v.detach() for k, v in params.items()}, shadow


def distillation_loss(teacher_sd, student_sd, ema_sd, cfg, abar, low, high, noise, idx, num_ddim_timesteps: int = 50,
                      num_inference_steps: int = 4, num_train_timesteps: int = 1000, strict_groupnorm: bool = True):
    """LowLightLCMDistillation.consistency_distillation_loss (low_light_diffusion.py:325-408), concat conditioning, with the
    two random draws (noise, idx) injected.  student_sd tensors may require grad."""
    c = num_train_timesteps // num_ddim_timesteps
    k = num_ddim_timesteps // num_inference_steps
    t = idx * c + c - 1
    t_next = (idx + k) * c + c - 1
    x_t = add_noise(high, noise, t, abar)
    ab = abar.to(high.device)
    a_t, a_n = ab[t].view(-1, 1, 1, 1), ab[t_next].view(-1, 1, 1, 1)
    with torch.no_grad():
        teacher_eps = unet_oracle.unet_forward(teacher_sd, cfg, torch.cat([x_t, low], dim=1), t, strict_groupnorm)
        x0_pred = (x_t - (1 - a_t).sqrt() * teacher_eps) / a_t.sqrt()
        x_next = a_n.sqrt() * x0_pred + (1 - a_n).sqrt() * teacher_eps
        target = unet_oracle.unet_forward(ema_sd, cfg, torch.cat([x_next, low], dim=1), t_next, strict_groupnorm)
    student = unet_oracle.unet_forward(student_sd, cfg, torch.cat([x_t, low], dim=1), t, strict_groupnorm)
    student_x0 = (x_t - (1 - a_t).sqrt() * student) / a_t.sqrt()
    target_x0 = (x_next - (1 - a_n).sqrt() * target) / a_n.sqrt()
    return F.huber_loss(student_x0, target_x0)
