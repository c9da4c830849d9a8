"""``LowLightDiffusion`` with the reference's public surface, running on the B200 library.

Reference: src/models/low_light_diffusion.py:31-281.  ``enhance`` keeps the reference semantics —
schedule from ``LCMScheduler.set_timesteps``, initial latents ``torch.randn(..., generator=generator)``,
per-step noise from the *global* RNG (``randn_like``; the reference ignores ``generator`` there),
final ``clamp(-1, 1)`` — but the whole 4-8 step loop is ONE native call: concat, UNet forward and
scheduler step of every iteration are fused kernels on the caller's stream, with no host sync.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Union

import torch
import torch.nn as nn

from .modules import EfficientUNet, create_efficient_unet
from .scheduler import LCMScheduler


@dataclass
class LowLightDiffusionOutput:
    enhanced: torch.Tensor
    intermediate: Optional[list] = None


class LowLightDiffusion(nn.Module):
    def __init__(self, unet: Optional[EfficientUNet] = None, scheduler: Optional[LCMScheduler] = None,
                 unet_variant: str = "small", image_size: int = 256, num_inference_steps: int = 4,
                 condition_mode: str = "concat", groupnorm: str = "strict", precision: Optional[str] = None):
        super().__init__()
        if condition_mode not in ("concat", "add"):
            raise ValueError(f"Unknown condition mode: {condition_mode}")
        self.image_size = image_size
        self.num_inference_steps = num_inference_steps
        self.condition_mode = condition_mode
        in_channels = 6 if condition_mode == "concat" else 3          # reference :77
        self.unet = unet if unet is not None else create_efficient_unet(
            variant=unet_variant, image_size=image_size, groupnorm=groupnorm, in_channels=in_channels)
        if precision is not None:          # an explicitly passed unet keeps its own precision unless told otherwise
            self.unet.precision = precision
        self.scheduler = scheduler if scheduler is not None else LCMScheduler(
            num_train_timesteps=1000, beta_schedule="scaled_linear", prediction_type="epsilon",
            num_inference_steps=num_inference_steps, rescale_betas_zero_snr=True)
        if condition_mode == "add":   # created after the UNet and the scheduler, like the reference (:108-113): same random init
            self.condition_encoder = nn.Sequential(nn.Conv2d(3, 32, 3, padding=1), nn.SiLU(), nn.Conv2d(32, 3, 3, padding=1))

    # ---- inference -------------------------------------------------------------------------------
    @torch.no_grad()
    def enhance(self, low_light: torch.Tensor, num_inference_steps: Optional[int] = None,
                generator: Optional[torch.Generator] = None, return_intermediate: bool = False,
                latents: Optional[torch.Tensor] = None, noises: Optional[torch.Tensor] = None
                ) -> Union[torch.Tensor, LowLightDiffusionOutput]:
        """Reference signature plus two optional injection points (``latents`` [B,3,S,S], ``noises``
        [steps-1,B,3,S,S]) used for parity testing and for sharding one global noise draw over GPUs."""
        from .engine import get_engine
        if not low_light.is_cuda:
            raise RuntimeError("low_light must be a CUDA tensor: the B200 path has no CPU fallback")
        if self.scheduler.config.prediction_type != "epsilon":
            raise ValueError("the fused enhance loop implements epsilon prediction (the pipeline default)")
        device, b = low_light.device, low_light.shape[0]
        s = self.image_size
        if tuple(low_light.shape) != (b, 3, s, s):
            raise ValueError(f"low_light must be [B,3,{s},{s}] (image_size), got {tuple(low_light.shape)}")
        steps = num_inference_steps if num_inference_steps is not None else self.num_inference_steps
        self.scheduler.set_timesteps(steps, device=device)
        ts = list(self.scheduler._host_timesteps)
        if latents is None:
            latents = torch.randn(b, 3, s, s, device=device, generator=generator)
        else:
            latents = latents.to(device=device, dtype=torch.float32).clone()
        coefs = [self.scheduler.step_coefficients(t) for t in ts]
        # "last" follows the reference's rule prev_t == 0 (lcm_scheduler.py:228), not the loop index: with
        # num_train_timesteps // original_inference_steps == 1 the schedule ends in t = 0 and the step before it already
        # returns x0 without drawing noise.  Such a step runs as x_prev = 1 * x0 + 0 * (zero noise).
        early_last = [c[4] for c in coefs[:-1]]
        if noises is None and steps > 1:
            noises = torch.stack([torch.zeros_like(latents) if last else torch.randn_like(latents) for last in early_last])
        if any(early_last):
            coefs = [(c[0], c[1], 1.0, 0.0, True) if c[4] else c for c in coefs]
            noises = noises.clone()
            for i, last in enumerate(early_last):
                if last:
                    noises[i].zero_()
        eng = get_engine(self.unet, b, s, s, device)
        low = low_light.to(torch.float32).contiguous()
        add = self.condition_mode == "add"
        if add:     # model_input = latents + condition_encoder(low_light) in every step (:223-225); the features are constant
            from .engine import condition_encode
            low = condition_encode(low, self.condition_encoder)
        res = eng.enhance(low, latents.contiguous(), noises, ts, coefs, trace=return_intermediate, add_mode=add)
        self.scheduler._step_index = steps
        if return_intermediate:
            out, tr = res
            return LowLightDiffusionOutput(enhanced=out, intermediate=[tr[i] for i in range(steps)])
        return res

    @torch.no_grad()
    def enhance_uint8(self, images: torch.Tensor, num_inference_steps: Optional[int] = None,
                      generator: Optional[torch.Generator] = None, target_size: Optional[int] = None) -> torch.Tensor:
        """uint8 RGB ``[N,H,W,3]`` in, uint8 RGB ``[N,H,W,3]`` out, everything on the device: the reference's
        ``preprocess_image`` / ``postprocess_image`` (scripts/inference.py:99-134) around :meth:`enhance` —
        ``cv2.resize`` to ``target_size`` x ``target_size`` (if given), ``x / 127.5 - 1``, HWC -> NCHW, enhance,
        ``clip((y + 1) * 127.5)`` -> uint8, NCHW -> HWC, ``cv2.resize`` back to the original size.  All three steps are
        bit-identical to the reference's numpy / OpenCV arithmetic.  A quarter of the PCIe bytes of the fp32 surface."""
        from . import ops
        h, w = int(images.shape[1]), int(images.shape[2])
        x = images if target_size is None else ops.image_resize_u8(images, target_size, target_size)
        y = ops.image_postprocess_u8(self.enhance(ops.image_preprocess_u8(x), num_inference_steps=num_inference_steps,
                                                  generator=generator))
        return y if target_size is None else ops.image_resize_u8(y, h, w)

    def forward(self, low_light: torch.Tensor, normal_light: Optional[torch.Tensor] = None,
                timesteps: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None,
                return_dict: bool = True):
        """Training forward (reference :115-171): random timesteps, noise, ``add_noise``, concat conditioning, noise
        prediction.  With autograd enabled ``noise_pred`` carries gradients to the UNet parameters (native backward)."""
        if normal_light is None:
            return self.enhance(low_light)
        b, device = low_light.shape[0], low_light.device
        if timesteps is None:
            timesteps = torch.randint(0, self.scheduler.config.num_train_timesteps, (b,), device=device)
        if noise is None:
            noise = torch.randn_like(normal_light)
        noisy = self.scheduler.add_noise(normal_light, noise, timesteps)
        if self.condition_mode == "concat":
            noise_pred = self.unet(torch.cat([noisy, low_light], dim=1), timesteps)
        else:   # :158-160 — inference-style evaluation only: the native backward covers the concat pipeline
            if torch.is_grad_enabled() and self.training:
                raise NotImplementedError('training with condition_mode="add" is not implemented natively (the condition '
                                          'encoder and the UNet input would need gradients); use condition_mode="concat"')
            from .engine import condition_encode
            noise_pred = self.unet(noisy + condition_encode(low_light.to(torch.float32), self.condition_encoder), timesteps)
        if return_dict:
            return {"noise_pred": noise_pred, "noise": noise, "timesteps": timesteps}
        return noise_pred

    def compute_loss(self, low_light: torch.Tensor, normal_light: torch.Tensor, loss_type: str = "mse",
                     timesteps: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Training loss (reference :250-277) as ONE native forward + loss; ``loss.backward()`` runs the native backward
        pass and fills ``param.grad`` of every UNet parameter, so the reference's loop (trainer.py:283-322: GradScaler,
        ``clip_grad_norm_``, any torch optimizer) works unchanged.  ``timesteps`` / ``noise`` are optional injection points
        (the reference draws them inside ``forward``)."""
        from .training import LOSS_TYPES, native_loss
        if loss_type not in LOSS_TYPES:
            raise ValueError(f"Unknown loss type: {loss_type}")
        if not low_light.is_cuda:
            raise RuntimeError("low_light must be a CUDA tensor: the B200 path has no CPU fallback")
        b, device = low_light.shape[0], low_light.device
        if timesteps is None:
            timesteps = torch.randint(0, self.scheduler.config.num_train_timesteps, (b,), device=device)
        if noise is None:
            noise = torch.randn_like(normal_light)
        if self.condition_mode != "concat":
            import torch.nn.functional as F
            out = self.forward(low_light, normal_light, timesteps=timesteps, noise=noise)
            return {"mse": F.mse_loss, "huber": F.huber_loss, "l1": F.l1_loss}[loss_type](out["noise_pred"], out["noise"])
        noisy = self.scheduler.add_noise(normal_light.contiguous(), noise.contiguous(), timesteps)
        if not torch.is_grad_enabled():
            import torch.nn.functional as F
            eps = self.unet(torch.cat([noisy, low_light], dim=1), timesteps)
            return {"mse": F.mse_loss, "huber": F.huber_loss, "l1": F.l1_loss}[loss_type](eps, noise)
        loss, _ = native_loss(self.unet, noisy, low_light, timesteps, noise, loss_type)
        return loss

    def get_model_size(self) -> Dict[str, float]:
        return self.unet.get_memory_footprint()

    # ---- checkpoints in the reference trainer's format (trainer.py:415-456, scripts/inference.py:78-79) ------------------
    def load_reference_checkpoint(self, checkpoint, use_ema: bool = False, strict: bool = True):
        """Load what the reference writes: a path / dict with ``model_state_dict`` (``save_checkpoint``), a bare
        ``state_dict`` (``scripts/benchmark.py:56``), optionally replacing the weights by the ``ema_shadow`` the trainer
        stores next to them (keys are the parameter names, values the EMA weights — ``EMAModel.shadow``, trainer.py:86-96)."""
        ckpt = torch.load(checkpoint, map_location="cpu", weights_only=False) if isinstance(checkpoint, (str, bytes)) or hasattr(checkpoint, "__fspath__") else checkpoint
        sd = ckpt.get("model_state_dict", ckpt) if isinstance(ckpt, dict) else ckpt
        sd = dict(sd)
        if use_ema:
            shadow = ckpt.get("ema_shadow") if isinstance(ckpt, dict) else None
            if not shadow:
                raise ValueError("the checkpoint has no ema_shadow")
            for k, v in shadow.items():
                if k not in sd:
                    raise ValueError(f"ema_shadow entry '{k}' is not a parameter of the checkpoint")
                sd[k] = v
        result = self.load_state_dict(sd, strict=strict)
        self.unet.invalidate_engines()
        meta = {k: ckpt[k] for k in ("epoch", "global_step", "best_val_loss", "config") if isinstance(ckpt, dict) and k in ckpt}
        return result, meta

    def reference_checkpoint(self, ema_shadow: Optional[Dict[str, torch.Tensor]] = None, **meta) -> Dict:
        """The dict ``LowLightTrainer.save_checkpoint`` writes for this model (weights part): loadable by the reference's
        ``load_checkpoint`` / ``scripts/inference.py`` unchanged."""
        out = {"model_state_dict": {k: v.detach().cpu().clone() for k, v in self.state_dict().items()}}
        if ema_shadow is not None:
            out["ema_shadow"] = {k: v.detach().cpu().clone() for k, v in ema_shadow.items()}
        out.update(meta)
        return out


def normalize_image(image: torch.Tensor) -> torch.Tensor:
    """[0,1] -> [-1,1] (reference :412-414)."""
    return image * 2 - 1


def denormalize_image(image: torch.Tensor) -> torch.Tensor:
    """[-1,1] -> [0,1] (reference :417-419)."""
    return (image + 1) / 2
