"""B200-native (sm_100a) LCM denoising hot path behind the reference's Python API.

    from cv_diffusion_model_b200 import LowLightDiffusion
    model = LowLightDiffusion(unet_variant="small", image_size=256).cuda().eval()
    enhanced = model.enhance(low_light)            # [B,3,256,256] in [-1,1]

See DESIGN.md for the path, include/lcm_unet.h for the C ABI.
"""
from .config import EfficientUNetConfig, variant_config
from .modules import EfficientUNet, create_efficient_unet
from .distillation import LowLightLCMDistillation
from .pipeline import LowLightDiffusion, LowLightDiffusionOutput, denormalize_image, normalize_image
from .scheduler import LCMScheduler, LCMSchedulerOutput, get_lcm_timesteps

__all__ = [
    "EfficientUNetConfig", "variant_config", "EfficientUNet", "create_efficient_unet", "LowLightDiffusion",
    "LowLightDiffusionOutput", "LowLightLCMDistillation", "LCMScheduler", "LCMSchedulerOutput", "get_lcm_timesteps", "normalize_image",
    "denormalize_image",
]
