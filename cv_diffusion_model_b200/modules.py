"""Parameter containers with the reference's module tree / state_dict layout.

These classes hold weights only.  They keep the reference's public surface —
``EfficientUNet(config)(x, timestep)``, ``create_efficient_unet(variant, image_size, **kw)``,
the 321-key ``state_dict`` layout of SURVEY App. B — while every bit of arithmetic
runs in the sm_100a library behind the C ABI (``include/lcm_unet.h``).  There is
deliberately no torch implementation of the forward pass here: without the CUDA
library ``forward`` raises.

Sub-modules are created in the same order as the reference constructor
(src/models/efficient_unet.py:403-530) so that ``torch.manual_seed(s)`` followed
by construction yields bit-identical random-init weights — this is what lets
golden vectors made from the reference be replayed on a box that does not have it.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import List, Tuple

import torch
import torch.nn as nn

from .config import ATTN_DIM_HEAD, EfficientUNetConfig, group_count, variant_config


class _Named(nn.Module):
    """A bag of named children (lets us reproduce `Sequential`-style numeric keys)."""

    def __init__(self, **children):
        super().__init__()
        for k, v in children.items():
            self.add_module(k, v)

    @classmethod
    def indexed(cls, pairs):
        m = cls()
        for idx, child in pairs:
            m.add_module(str(idx), child)
        return m

    def forward(self, *a, **k):  # pragma: no cover - containers are never called
        raise RuntimeError("parameter container; the forward pass lives in the CUDA library")


def _conv(ci, co, k, bias, groups=1):
    return nn.Conv2d(ci, co, k, padding=k // 2, groups=groups, bias=bias)


def _gn(c, strict):
    return nn.GroupNorm(group_count(c, strict), c)


def _inverted_residual(ci: int, co: int, cfg: EfficientUNetConfig, strict: bool, se_ratio=None) -> _Named:
    """Weights of one MobileNetV3-style block (reference :147-201).  Creation order
    norm1, norm2, expand, depthwise, se.fc1, se.fc2, project, time_mlp.1, [skip]."""
    ch = int(ci * cfg.expansion_ratio)
    blk = _Named()
    blk.add_module("norm1", _gn(ci, strict))
    blk.add_module("norm2", _gn(ch, strict))
    blk.add_module("expand", _conv(ci, ch, 1, bias=False))
    blk.add_module("depthwise", _conv(ch, ch, 3, bias=False, groups=ch))
    if cfg.use_se:
        sq = max(1, int(ch * (cfg.se_ratio if se_ratio is None else se_ratio)))
        blk.add_module("se", _Named(fc1=_conv(ch, sq, 1, bias=True), fc2=_conv(sq, ch, 1, bias=True)))
    blk.add_module("project", _conv(ch, co, 1, bias=False))
    blk.add_module("time_mlp", _Named.indexed([(1, nn.Linear(cfg.time_embed_dim, 2 * ch))]))
    if ci != co:
        blk.add_module("skip", _conv(ci, co, 1, bias=False))
    blk.meta = ("block", ci, ch, co)
    return blk


def _linear_attention(c: int, heads: int, strict: bool) -> _Named:
    """Weights of the O(n) attention (reference :250-271): norm, to_qkv, to_out.{0,1}."""
    inner = heads * ATTN_DIM_HEAD
    att = _Named()
    att.add_module("norm", _gn(c, strict))
    att.add_module("to_qkv", _conv(c, 3 * inner, 1, bias=False))
    att.add_module("to_out", _Named.indexed([(0, _conv(inner, c, 1, bias=False)), (1, _gn(c, strict))]))
    att.meta = ("attn", c, heads)
    return att


def _standard_attention(c: int, heads: int, strict: bool) -> _Named:
    """Weights of the softmax attention (reference :322-333): norm, to_qkv, to_out (a plain conv)."""
    inner = heads * ATTN_DIM_HEAD
    att = _Named()
    att.add_module("norm", _gn(c, strict))
    att.add_module("to_qkv", _conv(c, 3 * inner, 1, bias=False))
    att.add_module("to_out", _conv(inner, c, 1, bias=False))
    att.meta = ("attn", c, heads)
    return att


class EfficientUNet(nn.Module):
    """Drop-in for the reference ``EfficientUNet`` (efficient_unet.py:387-606).

    ``forward(x[B,Cin,H,W], timestep[B]) -> eps[B,3,H,W]`` executes on the B200
    library.  ``precision`` selects the native arithmetic: ``"bf16"`` (tcgen05
    GEMMs, bf16 activations, fp32 accumulation/statistics) or ``"fp32"``
    (verification mode, fp32 activations and CUDA-core GEMMs).
    """

    def __init__(self, config: EfficientUNetConfig, groupnorm: str = "strict"):
        super().__init__()
        if groupnorm not in ("strict", "gcd"):
            raise ValueError(f"Unknown groupnorm mode: {groupnorm}")
        if not config.quantization_friendly or not config.use_se:
            raise ValueError("the B200 path implements the preset blocks only (ReLU6 + SE)")
        self.config = config
        self.groupnorm = groupnorm
        strict = groupnorm == "strict"
        cfg = config
        widths = list(cfg.level_channels)
        _attention = _linear_attention if cfg.use_linear_attention else _standard_attention   # reference :448-454

        self.time_mlp = _Named.indexed([
            (1, nn.Linear(cfg.base_channels, cfg.time_embed_dim)),
            (3, nn.Linear(cfg.time_embed_dim, cfg.time_embed_dim)),
        ])
        self.init_conv = _conv(cfg.in_channels, widths[0], 3, bias=True)

        def level(ci_first, co, n_blocks, res):
            mods, ci = [], ci_first
            for _ in range(n_blocks):
                mods.append(_inverted_residual(ci, co, cfg, strict))
                if res in cfg.attention_resolutions:
                    mods.append(_attention(co, cfg.num_attention_heads, strict))
                ci = co
            return nn.ModuleList(mods)

        self.encoder_blocks = nn.ModuleList()
        self.downsamplers = nn.ModuleList()
        res, ci = cfg.image_size, widths[0]
        for li, co in enumerate(widths):
            self.encoder_blocks.append(level(ci, co, cfg.num_res_blocks, res))
            ci = co
            if li < len(widths) - 1:
                self.downsamplers.append(_Named(down=nn.Conv2d(co, co, 3, stride=2, padding=1)))
                res //= 2

        mid = widths[-1]
        # the reference builds the two mid blocks with the constructor default se_ratio=0.25, not config.se_ratio
        # (efficient_unet.py:467-478)
        self.mid_block1 = _inverted_residual(mid, mid, cfg, strict, se_ratio=0.25)
        self.mid_attn = _attention(mid, cfg.num_attention_heads, strict)
        self.mid_block2 = _inverted_residual(mid, mid, cfg, strict, se_ratio=0.25)

        self.decoder_blocks = nn.ModuleList()
        self.upsamplers = nn.ModuleList()
        for li, co in enumerate(reversed(widths)):
            mods, first = [], True
            for _ in range(cfg.num_res_blocks + 1):
                mods.append(_inverted_residual(ci + co if first else co, co, cfg, strict))
                first = False
                if res in cfg.attention_resolutions:
                    mods.append(_attention(co, cfg.num_attention_heads, strict))
            self.decoder_blocks.append(nn.ModuleList(mods))
            ci = co
            if li < len(widths) - 1:
                self.upsamplers.append(_Named(conv=_conv(co, co, 3, bias=True)))
                res *= 2

        self.final_norm = _gn(widths[0], strict)
        self.final_conv = _conv(widths[0], cfg.out_channels, 3, bias=True)

        self.precision = "bf16"
        self._engines = OrderedDict()      # LRU of native plans, see engine.get_engine
        self._weights_epoch = 0

    # ---- reference surface --------------------------------------------------
    def forward(self, x: torch.Tensor, timestep: torch.Tensor, return_features: bool = False):
        if return_features:
            # analysis hook of the reference (:595-596,604-605): the output of every decoder level, read back from a
            # plan that keeps its intermediates (no buffer reuse; slower, inference only)
            from .engine import get_engine
            eng = get_engine(self, x.shape[0], x.shape[2], x.shape[3], x.device, taps=True)
            with torch.no_grad():
                out = eng.forward(x, timestep)
                last = {}
                for name in eng.taps():
                    if name.startswith("decoder_blocks.") and name.endswith(".out"):
                        last[int(name.split(".")[1])] = name          # the level's last block / attention output
                feats = [eng.read_tap(last[k]) for k in sorted(last)]
            return out, feats
        if self.training and torch.is_grad_enabled() and x.is_cuda and x.shape[1] > self.config.out_channels \
                and any(p.requires_grad for p in self.parameters()):
            from .training import native_unet_forward   # training: activations are kept, eps carries autograd
            return native_unet_forward(self, x, timestep)
        from .engine import unet_forward  # late import: needs the CUDA library
        return unet_forward(self, x, timestep)

    def __deepcopy__(self, memo):
        """Copies parameters and configuration; native plans (device handles, workspaces) are never shared or copied
        (``copy.deepcopy(student_model)`` in the distillation wrapper, low_light_diffusion.py:311-313)."""
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = OrderedDict() if k == "_engines" else copy.deepcopy(v, memo)
        return new

    def get_num_params(self) -> int:
        return sum(p.numel() for p in self.parameters() if p.requires_grad)

    def get_memory_footprint(self, input_size: Tuple[int, int] = (256, 256)) -> dict:
        n = self.get_num_params()
        mb = 1024 ** 2
        return {"num_params": n, "fp32_mb": n * 4 / mb, "fp16_mb": n * 2 / mb, "int8_mb": n / mb}

    # ---- native plumbing ----------------------------------------------------
    def mark_weights_changed(self) -> None:
        """Tell the native plans that parameters were modified in a way autograd's version counters do not see
        (``param.data.copy_``, e.g. the reference's EMA ``apply_shadow`` / ``restore``, trainer.py:106-118); the next
        call re-packs the weights.  Cheaper than :meth:`invalidate_engines` (plans and workspaces are kept)."""
        self._weights_epoch += 1

    def invalidate_engines(self) -> None:
        """Drop native plans (call after mutating weights in place, e.g. EMA swap)."""
        for e in self._engines.values():
            e.close()
        self._engines.clear()

    def load_state_dict(self, *a, **k):
        out = super().load_state_dict(*a, **k)
        self.invalidate_engines()
        return out


def create_efficient_unet(variant: str = "small", image_size: int = 256, groupnorm: str = "strict",
                          **kwargs) -> EfficientUNet:
    """Same call as the reference factory (efficient_unet.py:631-692)."""
    return EfficientUNet(variant_config(variant, image_size, **kwargs), groupnorm=groupnorm)


def forward_order(unet: EfficientUNet) -> List[Tuple[str, str]]:
    """(kind, dotted-name) of every parameterised stage in execution order; used by
    tests to cross-check the native plan's own enumeration."""
    out = [("time_mlp", "time_mlp"), ("conv", "init_conv")]

    def walk(prefix, mods):
        for i, m in enumerate(mods):
            out.append((m.meta[0], f"{prefix}.{i}"))

    for li, lvl in enumerate(unet.encoder_blocks):
        walk(f"encoder_blocks.{li}", lvl)
        if li < len(unet.downsamplers):
            out.append(("down", f"downsamplers.{li}"))
    out += [("block", "mid_block1"), ("attn", "mid_attn"), ("block", "mid_block2")]
    for li, lvl in enumerate(unet.decoder_blocks):
        if li > 0:
            out.append(("up", f"upsamplers.{li - 1}"))
        walk(f"decoder_blocks.{li}", lvl)
    out.append(("final", "final_conv"))
    return out
