"""Configuration of the time-conditioned EfficientUNet epsilon-predictor.

Mirrors the reference dataclass field for field so that user code which builds
an ``EfficientUNetConfig`` keeps working
(reference: src/models/efficient_unet.py:24-57 for the fields,
 :646-687 for the four variant presets).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, replace
from typing import Tuple


@dataclass
class EfficientUNetConfig:
    in_channels: int = 3
    out_channels: int = 3
    base_channels: int = 32
    channel_multipliers: Tuple[int, ...] = (1, 2, 4, 8)
    attention_resolutions: Tuple[int, ...] = (16, 8)
    num_attention_heads: int = 4
    use_linear_attention: bool = True
    num_res_blocks: int = 2
    expansion_ratio: int = 4
    use_se: bool = True
    se_ratio: float = 0.25
    time_embed_dim: int = 128
    dropout: float = 0.0
    quantization_friendly: bool = True
    image_size: int = 256

    # ---- derived quantities used by the native plan -------------------------
    @property
    def level_channels(self) -> Tuple[int, ...]:
        return tuple(self.base_channels * m for m in self.channel_multipliers)


# (base_channels, num_res_blocks, expansion_ratio, time_embed_dim, heads)
_VARIANTS = {
    "tiny": (16, 1, 2, 64, 2),
    "small": (32, 2, 4, 128, 4),
    "base": (48, 2, 4, 192, 6),
    "large": (64, 3, 4, 256, 8),
}

ATTN_DIM_HEAD = 32  # LinearAttention(dim_head=32), efficient_unet.py:254


def variant_config(variant: str = "small", image_size: int = 256, **overrides) -> EfficientUNetConfig:
    """Preset lookup; unknown names raise ValueError like the reference (:689-690)."""
    if variant not in _VARIANTS:
        raise ValueError(f"Unknown variant: {variant}. Choose from {list(_VARIANTS.keys())}")
    base, nres, exp, ted, heads = _VARIANTS[variant]
    cfg = EfficientUNetConfig(
        base_channels=base,
        channel_multipliers=(1, 2, 4, 8),
        num_res_blocks=nres,
        expansion_ratio=exp,
        time_embed_dim=ted,
        num_attention_heads=heads,
        image_size=image_size,
    )
    return replace(cfg, **overrides)


def group_count(channels: int, strict: bool = True) -> int:
    """Number of GroupNorm groups for a `channels`-wide tensor.

    The reference uses ``min(32, C)`` (efficient_unet.py:170-171,263,268,528), which
    is invalid when C is not a multiple of it (tiny: C=48, base: C=48/144 — the
    reference raises at construction).  ``strict=False`` selects the
    documented minimal deviation ``gcd(32, C)``, identical wherever the
    reference is valid.
    """
    g = min(32, channels)
    if channels % g == 0:
        return g
    if strict:
        raise ValueError(f"num_channels ({channels}) must be divisible by num_groups ({g})")
    return math.gcd(32, channels)
