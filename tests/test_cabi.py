"""CPU: the C-ABI library builds/loads, exports every symbol include/lcm_unet.h declares, and rejects
invalid configurations with the reference's error behaviour — no compute calls (there is no GPU here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from cv_diffusion_model_b200 import build, native
    build.build()
    return native.lib()


def test_header_symbols_exported_and_bound(lib):
    from cv_diffusion_model_b200 import native
    header = open(os.path.join(ROOT, "include", "lcm_unet.h")).read()
    declared = set(re.findall(r"\b(lcm_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in lcm_unet.h but not exported"
    assert declared == set(native.SIGNATURES), declared ^ set(native.SIGNATURES)
    assert lib.lcm_version() >= 1


def test_config_struct_layout():
    from cv_diffusion_model_b200 import native
    from cv_diffusion_model_b200.config import variant_config
    c = native.config_struct(variant_config("small", 256, in_channels=6), "strict")
    assert C.sizeof(c) == 4 * (4 + 8 + 1 + 8 + 3 + 1 + 3 + 1)
    assert c.standard_attention == 0 and native.config_struct(variant_config("small", 256, use_linear_attention=False), "strict").standard_attention == 1
    assert (c.base_channels, c.num_levels, list(c.channel_multipliers)[:4]) == (32, 4, [1, 2, 4, 8])
    assert list(c.attention_resolutions)[:2] == [16, 8] and c.time_embed_dim == 128 and c.image_size == 256


def test_invalid_configs_raise_value_error(lib):
    from cv_diffusion_model_b200 import native
    from cv_diffusion_model_b200.config import variant_config
    h = C.c_void_p()
    # tiny without the GroupNorm patch: the reference raises ValueError at construction
    cfg = native.config_struct(variant_config("tiny", 256, in_channels=6), "strict")
    with pytest.raises(ValueError, match="divisible"):
        native.check(lib.lcm_plan_create(C.byref(cfg), 1, 64, 64, native.PREC_BF16, 0, 0, C.byref(h)))
    cfg = native.config_struct(variant_config("small", 256, in_channels=6), "strict")
    with pytest.raises(ValueError):
        native.check(lib.lcm_plan_create(C.byref(cfg), 1, 60, 64, native.PREC_BF16, 0, 0, C.byref(h)))   # not /8
    with pytest.raises(ValueError, match="precision"):
        native.check(lib.lcm_plan_create(C.byref(cfg), 1, 64, 64, 7, 0, 0, C.byref(h)))
    with pytest.raises(ValueError):
        native.check(lib.lcm_scheduler_step(None, None, None, None, None, 0, 0, 0.0, 1.0, 1.0, 0.0, None))


def test_no_cpu_fallback(lib):
    import torch
    from cv_diffusion_model_b200 import LowLightDiffusion, native
    from cv_diffusion_model_b200.config import variant_config
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = native.config_struct(variant_config("small", 256, in_channels=6), "strict")
    h = C.c_void_p()
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        native.check(lib.lcm_plan_create(C.byref(cfg), 1, 64, 64, native.PREC_BF16, 0, 0, C.byref(h)))
    pipe = LowLightDiffusion(unet_variant="small", image_size=32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pipe.enhance(torch.zeros(1, 3, 32, 32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pipe.scheduler.set_timesteps(4)
        pipe.scheduler.step(torch.zeros(1, 3, 4, 4), 739, torch.zeros(1, 3, 4, 4))


def test_product_path_does_not_import_oracle():
    pkg = os.path.join(ROOT, "cv_diffusion_model_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src, f"{f} mentions the oracle"
