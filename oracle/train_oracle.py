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
