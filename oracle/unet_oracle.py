"""ORACLE — test infrastructure only; never imported by the product path.

A CPU restatement, in plain functional torch fp32 ops, of the reference's
epsilon-predictor ``EfficientUNet.forward``
(/root/reference/src/models/efficient_unet.py:532-606) driven directly by a flat
``state_dict`` (keys without the ``unet.`` prefix).  It exists so that the CUDA
path can be checked on a GPU box where /root/reference does not exist.

Parity status: PINNED.  ``tests/golden/make_golden.py`` (run in the build
container) imports the unmodified reference and checks this file against it
bit-for-bit on CPU (same ATen ops in the same order), and the known-answer values
of SURVEY App. C are replayed from ``tests/golden/*.npz`` by
``tests/test_oracle.py``.  For the variants the reference cannot construct
(tiny, base: GroupNorm(min(32,C), C) with C=48/144 — SURVEY F1) the comparator
is the reference with ``num_groups=gcd(32,C)``; every such result is labelled
"patched".

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional

import torch
import torch.nn.functional as F


def _groups(c: int, strict: bool) -> int:
    g = min(32, c)  # efficient_unet.py:170-171,263,268,528
    if c % g and not strict:
        g = math.gcd(32, c)
    return g


def sinusoidal_embedding(t: torch.Tensor, dim: int, max_period: int = 10000) -> torch.Tensor:
    """efficient_unet.py:68-76 — cos half first, then sin half."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def time_embedding(sd: Dict[str, torch.Tensor], t: torch.Tensor, base_channels: int) -> torch.Tensor:
    """efficient_unet.py:412-417,550."""
    e = sinusoidal_embedding(t, base_channels)
    e = F.linear(e, sd["time_mlp.1.weight"], sd["time_mlp.1.bias"])
    e = F.silu(e)
    return F.linear(e, sd["time_mlp.3.weight"], sd["time_mlp.3.bias"])


def inverted_residual(sd, p: str, x: torch.Tensor, t_emb: torch.Tensor, strict: bool,
                      tap: Optional[Callable] = None) -> torch.Tensor:
    """efficient_unet.py:203-236."""
    ci = x.shape[1]
    ch = sd[p + "expand.weight"].shape[0]
    h = F.group_norm(x, _groups(ci, strict), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
    h = F.relu6(h)
    h = F.conv2d(h, sd[p + "expand.weight"])
    if tap: tap(p + "expand", h)
    h = F.group_norm(h, _groups(ch, strict), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
    ss = F.linear(F.silu(t_emb), sd[p + "time_mlp.1.weight"], sd[p + "time_mlp.1.bias"])[:, :, None, None]
    scale, shift = ss.chunk(2, dim=1)
    h = h * (1 + scale) + shift
    h = F.relu6(h)
    h = F.conv2d(h, sd[p + "depthwise.weight"], padding=1, groups=ch)
    if tap: tap(p + "depthwise", h)
    g = F.adaptive_avg_pool2d(h, 1)                                   # :97
    g = F.relu6(F.conv2d(g, sd[p + "se.fc1.weight"], sd[p + "se.fc1.bias"]))
    g = torch.sigmoid(F.conv2d(g, sd[p + "se.fc2.weight"], sd[p + "se.fc2.bias"]))
    if tap: tap(p + "se_gate", g)
    h = h * g
    h = F.conv2d(h, sd[p + "project.weight"])
    res = x
    if (p + "skip.weight") in sd:
        res = F.conv2d(x, sd[p + "skip.weight"])
    h = h + res
    if tap: tap(p + "out", h)
    return h


def linear_attention(sd, p: str, x: torch.Tensor, heads: int, strict: bool,
                     tap: Optional[Callable] = None) -> torch.Tensor:
    """efficient_unet.py:273-308 (dim_head fixed at 32, :254)."""
    b, c, hh, ww = x.shape
    h = F.group_norm(x, _groups(c, strict), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)
    qkv = F.conv2d(h, sd[p + "to_qkv.weight"])
    if tap: tap(p + "qkv", qkv)
    q, k, v = qkv.chunk(3, dim=1)
    d = q.shape[1] // heads

    def split(z):  # 'b (heads d) h w -> b heads (h w) d'
        return z.reshape(b, heads, d, hh * ww).transpose(2, 3)

    q, k, v = split(q), split(k), split(v)
    q = F.elu(q) + 1
    k = F.elu(k) + 1
    k_sum = k.sum(dim=-2, keepdim=True)
    kv = torch.einsum("bhnd,bhne->bhde", k, v)
    qk_sum = torch.einsum("bhnd,bhkd->bhnk", q, k_sum)
    out = torch.einsum("bhnd,bhde->bhne", q, kv) / (qk_sum + 1e-6)
    out = out.transpose(2, 3).reshape(b, heads * d, hh, ww)
    if tap: tap(p + "attn", out)
    out = F.conv2d(out, sd[p + "to_out.0.weight"])
    out = F.group_norm(out, _groups(c, strict), sd[p + "to_out.1.weight"], sd[p + "to_out.1.bias"], 1e-5)
    out = out + x
    if tap: tap(p + "out", out)
    return out


def standard_attention(sd, p: str, x: torch.Tensor, heads: int, strict: bool, tap: Optional[Callable] = None) -> torch.Tensor:
    """efficient_unet.py:335-357 (use_linear_attention=False): softmax(q k^T * dim_head^-0.5) v, plain to_out conv, + x."""
    b, c, hh, ww = x.shape
    h = F.group_norm(x, _groups(c, strict), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)
    qkv = F.conv2d(h, sd[p + "to_qkv.weight"])
    if tap: tap(p + "qkv", qkv)
    q, k, v = qkv.chunk(3, dim=1)
    d = q.shape[1] // heads

    def split(z):
        return z.reshape(b, heads, d, hh * ww).transpose(2, 3)

    q, k, v = split(q), split(k), split(v)
    attn = torch.einsum("bhid,bhjd->bhij", q, k) * (d ** -0.5)
    attn = F.softmax(attn, dim=-1)
    out = torch.einsum("bhij,bhjd->bhid", attn, v)
    out = out.transpose(2, 3).reshape(b, heads * d, hh, ww)
    if tap: tap(p + "attn", out)
    out = F.conv2d(out, sd[p + "to_out.weight"]) + x
    if tap: tap(p + "out", out)
    return out


def unet_forward(sd: Dict[str, torch.Tensor], cfg, x: torch.Tensor, timestep: torch.Tensor,
                 strict_groupnorm: bool = True, tap: Optional[Callable] = None) -> torch.Tensor:
    """efficient_unet.py:532-606.  `cfg` needs: base_channels, channel_multipliers,
    num_res_blocks, num_attention_heads, attention_resolutions, image_size."""
    st = strict_groupnorm
    widths = [cfg.base_channels * m for m in cfg.channel_multipliers]
    attention = linear_attention if getattr(cfg, "use_linear_attention", True) else standard_attention   # :448-454,473-474
    t_emb = time_embedding(sd, timestep, cfg.base_channels)
    if tap: tap("t_emb", t_emb)
    h = F.conv2d(x, sd["init_conv.weight"], sd["init_conv.bias"], padding=1)
    if tap: tap("init_conv", h)

    def run_level(prefix, n_blocks, res, h):
        idx = 0
        for _ in range(n_blocks):
            h = inverted_residual(sd, f"{prefix}.{idx}.", h, t_emb, st, tap)
            idx += 1
            if res in cfg.attention_resolutions:       # :447, absolute resolution from config.image_size
                h = attention(sd, f"{prefix}.{idx}.", h, cfg.num_attention_heads, st, tap)
                idx += 1
        return h

    skips, res = [], cfg.image_size
    for li in range(len(widths)):
        h = run_level(f"encoder_blocks.{li}", cfg.num_res_blocks, res, h)
        skips.append(h)                                                  # :567
        if li < len(widths) - 1:
            h = F.conv2d(h, sd[f"downsamplers.{li}.down.weight"], sd[f"downsamplers.{li}.down.bias"],
                         stride=2, padding=1)                            # :367
            if tap: tap(f"downsamplers.{li}", h)
            res //= 2
    h = inverted_residual(sd, "mid_block1.", h, t_emb, st, tap)
    h = attention(sd, "mid_attn.", h, cfg.num_attention_heads, st, tap)
    h = inverted_residual(sd, "mid_block2.", h, t_emb, st, tap)
    for li in range(len(widths)):
        if li > 0:
            h = F.interpolate(h, scale_factor=2, mode="bilinear", align_corners=False)   # :383
            h = F.conv2d(h, sd[f"upsamplers.{li - 1}.conv.weight"], sd[f"upsamplers.{li - 1}.conv.bias"], padding=1)
            if tap: tap(f"upsamplers.{li - 1}", h)
            res *= 2
        h = torch.cat([h, skips.pop()], dim=1)                           # :587-588, h first
        h = run_level(f"decoder_blocks.{li}", cfg.num_res_blocks + 1, res, h)
    h = F.group_norm(h, _groups(widths[0], st), sd["final_norm.weight"], sd["final_norm.bias"], 1e-5)
    h = F.silu(h)
    h = F.conv2d(h, sd["final_conv.weight"], sd["final_conv.bias"], padding=1)
    if tap: tap("eps", h)
    return h


def strip_unet_prefix(state_dict: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """`LowLightDiffusion.state_dict()` keys are `unet.*`."""
    return {k[5:] if k.startswith("unet.") else k: v for k, v in state_dict.items()}
