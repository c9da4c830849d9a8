"""LCM scheduler with the reference's public surface, executing on the B200 library.

Reference: src/models/lcm_scheduler.py:34-305.  What is kept identical:
  * constructor arguments and ``.config.<name>`` access (:54-66),
  * the fp32 abar table incl. the zero-terminal-SNR rescale (:76-100, :116-129) —
    built once on the host with the same torch CPU ops so it is bit-identical,
  * the inference schedule (:150-161): e.g. 4 steps -> [739, 499, 259, 19],
  * ``step`` (:204-242): x0 = (x_t - sqrt(1-abar_t) eps) / sqrt(abar_t);
    x_prev = x0 on the last step, else sqrt(abar_prev) x0 + sqrt(1-abar_prev) noise.
    There is no consistency c_skip/c_out scaling and no x0 clamp in the reference.
What differs: the elementwise arithmetic of ``step``/``add_noise`` runs in one
fused CUDA kernel, and the host never synchronises with the device (the
reference does ``nonzero``/``.item()`` on device tensors every step, :169-174).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Tuple, Union

import torch


@dataclass
class LCMSchedulerOutput:
    prev_sample: torch.Tensor
    pred_original_sample: Optional[torch.Tensor] = None


class _Config(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


def lcm_timestep_list(num_inference_steps: int, num_train_timesteps: int = 1000,
                      original_inference_steps: int = 50) -> List[int]:
    """Integer restatement of the schedule rule (lcm_scheduler.py:150-161)."""
    c = num_train_timesteps // original_inference_steps
    origin = [(i + 1) * c - 1 for i in range(original_inference_steps)]
    skip = len(origin) // num_inference_steps
    picked = origin[::skip][:num_inference_steps]
    return picked[::-1]


class LCMScheduler:
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                 beta_schedule: str = "scaled_linear", prediction_type: str = "epsilon",
                 timestep_spacing: str = "leading", rescale_betas_zero_snr: bool = False,
                 num_inference_steps: int = 4, original_inference_steps: int = 50, lcm_origin_steps: int = 50):
        self.config = _Config(
            num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
            beta_schedule=beta_schedule, prediction_type=prediction_type, timestep_spacing=timestep_spacing,
            rescale_betas_zero_snr=rescale_betas_zero_snr, num_inference_steps=num_inference_steps,
            original_inference_steps=original_inference_steps, lcm_origin_steps=lcm_origin_steps)
        n = num_train_timesteps
        if beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, n)
        elif beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, n) ** 2
        elif beta_schedule == "squaredcos_cap_v2":
            x = torch.linspace(0, n, n + 1)
            ac = torch.cos(((x / n) + 0.008) / 1.008 * math.pi * 0.5) ** 2
            ac = ac / ac[0]
            betas = torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)
        else:
            raise ValueError(f"Unknown beta schedule: {beta_schedule}")
        if prediction_type not in ("epsilon", "v_prediction"):
            # the reference raises lazily inside step(); failing early is stricter, same type
            raise ValueError(f"Unknown prediction type: {prediction_type}")
        self.betas = betas
        self.alphas = 1.0 - betas
        abar = torch.cumprod(self.alphas, dim=0)
        if rescale_betas_zero_snr:
            root = abar.sqrt()
            r0, rT = root[0].clone(), root[-1].clone()
            root -= rT
            root *= r0 / (r0 - rT)
            abar = root ** 2
        self.alphas_cumprod = abar
        self.sigmas = ((1 - abar) / abar) ** 0.5
        self.final_alpha_cumprod = abar[0]
        self.num_inference_steps = None
        self.timesteps = None
        self._step_index = None
        self._host_timesteps: List[int] = []

    # ---- schedule -----------------------------------------------------------
    def set_timesteps(self, num_inference_steps: int = 4, device: Union[str, torch.device] = "cpu",
                      original_inference_steps: Optional[int] = None):
        self.num_inference_steps = num_inference_steps
        if original_inference_steps is None:
            original_inference_steps = self.config.original_inference_steps
        self._host_timesteps = lcm_timestep_list(num_inference_steps, self.config.num_train_timesteps,
                                                 original_inference_steps)
        self.timesteps = torch.tensor(self._host_timesteps, dtype=torch.long).to(device)
        self._step_index = 0
        self.sigmas = self.sigmas.to(device)

    def _get_prev_timestep(self, timestep: int) -> int:
        i = self._host_timesteps.index(int(timestep))
        return self._host_timesteps[i + 1] if i + 1 < len(self._host_timesteps) else 0

    def step_coefficients(self, timestep: int) -> Tuple[float, float, float, float, bool]:
        """fp32 scalars of one step, computed exactly like the reference's 0-d CPU tensor
        arithmetic (:208-217,239-242): (sqrt(1-abar_t), sqrt(abar_t), sqrt(abar_prev),
        sqrt(1-abar_prev), is_last)."""
        t = int(timestep)
        prev_t = self._get_prev_timestep(t)
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t > 0 else self.final_alpha_cumprod
        return (float((1 - a_t) ** 0.5), float(a_t ** 0.5), float(a_prev ** 0.5), float((1 - a_prev) ** 0.5),
                prev_t == 0)

    # ---- arithmetic (native) -----------------------------------------------
    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor,
             generator: Optional[torch.Generator] = None, return_dict: bool = True,
             noise: Optional[torch.Tensor] = None):
        """One denoising step.  ``noise`` (extension) injects the step noise; when omitted it
        is drawn with ``torch.randn_like(sample)`` from the *global* RNG exactly like the
        reference (:237 — ``generator`` is ignored there too)."""
        from .engine import lcm_step  # late import: needs the CUDA library
        if self._step_index is None:
            self._step_index = 0
        sb_t, sa_t, sa_p, sb_p, last = self.step_coefficients(int(timestep))
        if not last and noise is None:
            noise = torch.randn_like(sample)
        prev, x0 = lcm_step(model_output, sample, None if last else noise, self.config.prediction_type,
                            sb_t, sa_t, sa_p, sb_p)
        self._step_index += 1
        if return_dict:
            return LCMSchedulerOutput(prev_sample=prev, pred_original_sample=x0)
        return (prev, x0)

    def add_noise(self, original_samples: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor):
        from .engine import lcm_mix
        return lcm_mix(original_samples, noise, timesteps, self.alphas_cumprod, velocity=False)

    def get_velocity(self, sample: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor):
        from .engine import lcm_mix
        return lcm_mix(sample, noise, timesteps, self.alphas_cumprod, velocity=True)


def get_lcm_timesteps(num_inference_steps: int = 4, num_train_timesteps: int = 1000,
                      original_inference_steps: int = 50) -> List[int]:
    """Reference helper (lcm_scheduler.py:421-442): same rule as ``set_timesteps``."""
    return lcm_timestep_list(num_inference_steps, num_train_timesteps, original_inference_steps)
