"""``LowLightLCMDistillation`` with the reference's surface (src/models/low_light_diffusion.py:284-408).

teacher -> student consistency distillation: the teacher makes one DDIM jump of ``k`` solver steps from ``x_t``, the EMA
student predicts from the jumped sample, the student is trained so that its x0 prediction at ``(x_t, t)`` matches the EMA
student's at ``(x_{t_next}, t_next)`` (Huber).  Three UNet forwards (teacher, student, EMA student) and one backward
(student), all on the native plans; the element-wise steps between them are two small native kernels
(``lcm_ddim_step``, ``lcm_consistency_loss``).
"""
from __future__ import annotations

import copy
import ctypes as C
from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import native
from .engine import _stream_ptr
from .pipeline import LowLightDiffusion


class _ConsistencyLoss(torch.autograd.Function):
    """mean Huber(student_x0, target_x0) as a function of the student's eps; the kernel returns loss and d loss / d eps."""

    @staticmethod
    def forward(ctx, eps_student, x_t, t, x_next, eps_target, t_next, abar):
        lib = native.lib()
        b = x_t.shape[0]
        loss = torch.zeros(1, dtype=torch.float64, device=x_t.device)
        d_eps = torch.empty_like(x_t)
        with torch.cuda.device(x_t.device):
            native.check(lib.lcm_consistency_loss(
                C.c_void_p(x_t.data_ptr()), C.c_void_p(eps_student.data_ptr()), C.c_void_p(t.data_ptr()),
                C.c_void_p(x_next.data_ptr()), C.c_void_p(eps_target.data_ptr()), C.c_void_p(t_next.data_ptr()),
                C.c_void_p(abar.data_ptr()), C.c_void_p(loss.data_ptr()), C.c_void_p(d_eps.data_ptr()), b,
                x_t.numel() // b, _stream_ptr()))
        ctx.save_for_backward(d_eps)
        return loss[0].to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        (d_eps,) = ctx.saved_tensors
        return d_eps * g, None, None, None, None, None, None


class LowLightLCMDistillation(nn.Module):
    def __init__(self, teacher_model: LowLightDiffusion, student_model: LowLightDiffusion, num_ddim_timesteps: int = 50,
                 guidance_scale_range: Tuple[float, float] = (3.0, 15.0)):
        super().__init__()
        self.teacher = teacher_model
        self.teacher.eval()
        self.teacher.requires_grad_(False)
        self.student = student_model
        self.num_ddim_timesteps = num_ddim_timesteps
        self.guidance_scale_range = guidance_scale_range
        self.ema_student = copy.deepcopy(student_model)       # target model (:311-315)
        self.ema_student.eval()
        self.ema_student.requires_grad_(False)

    @torch.no_grad()
    def update_ema(self, decay: float = 0.95):
        """:317-323"""
        for e, s in zip(self.ema_student.parameters(), self.student.parameters()):
            e.data.mul_(decay).add_(s.data, alpha=1 - decay)
        self.ema_student.unet.mark_weights_changed()

    def consistency_distillation_loss(self, low_light: torch.Tensor, normal_light: torch.Tensor, num_inference_steps: int = 4,
                                      noise: Optional[torch.Tensor] = None, idx: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Reference :325-408 (``noise`` / ``idx`` are optional injection points; the reference draws them here)."""
        if self.teacher.condition_mode != "concat" or self.student.condition_mode != "concat":
            raise ValueError('only condition_mode="concat" is implemented natively')
        if not low_light.is_cuda:
            raise RuntimeError("low_light must be a CUDA tensor: the B200 path has no CPU fallback")
        b, device = low_light.shape[0], low_light.device
        if noise is None:
            noise = torch.randn_like(normal_light)
        c = self.teacher.scheduler.config.num_train_timesteps // self.num_ddim_timesteps
        k = self.num_ddim_timesteps // num_inference_steps
        if idx is None:
            idx = torch.randint(0, self.num_ddim_timesteps - k, (b,), device=device)
        t = (idx * c + c - 1).to(torch.long).contiguous()
        t_next = ((idx + k) * c + c - 1).to(torch.long).contiguous()
        sched = self.teacher.scheduler
        abar = sched.alphas_cumprod.to(device=device, dtype=torch.float32).contiguous()
        x_t = sched.add_noise(normal_light.contiguous(), noise.contiguous(), t)
        low = low_light.to(torch.float32).contiguous()
        lib = native.lib()
        with torch.no_grad():
            teacher_eps = self.teacher.unet(torch.cat([x_t, low], dim=1), t)
            x_next = torch.empty_like(x_t)
            with torch.cuda.device(device):
                native.check(lib.lcm_ddim_step(C.c_void_p(x_t.data_ptr()), C.c_void_p(teacher_eps.data_ptr()), C.c_void_p(t.data_ptr()),
                                               C.c_void_p(t_next.data_ptr()), C.c_void_p(abar.data_ptr()),
                                               C.c_void_p(x_next.data_ptr()), b, x_t.numel() // b, _stream_ptr()))
            target_eps = self.ema_student.unet(torch.cat([x_next, low], dim=1), t_next)
        student_eps = self.student.unet(torch.cat([x_t, low], dim=1), t)          # autograd -> native backward
        return _ConsistencyLoss.apply(student_eps.contiguous(), x_t, t, x_next, target_eps.contiguous(), t_next, abar)
