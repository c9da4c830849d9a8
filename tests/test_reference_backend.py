"""The reference-side binding (cv_diffusion_model_b200/reference_backend.py, shown in INTEGRATION.md) is real code:

* CPU / build container: with the UNMODIFIED reference importable (behind the diffusers stub of oracle/ref_shim) the
  binding's config mapping and the weight names it will feed to the plan agree with the reference's own objects.
* GPU box (no /root/reference there): a reference-shaped model (its own `unet` container with the reference's
  state_dict layout, a torch-only scheduler object exposing what `attach` touches) is attached and must reproduce the
  package's own `enhance` bit for bit under the reference's RNG protocol.
"""
import ctypes as C
import os
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def _import_reference():
    for p in (os.path.join(ROOT, "oracle", "ref_shim"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    from src.models.low_light_diffusion import LowLightDiffusion as RefPipeline
    return RefPipeline


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="the reference only exists in the build container")
def test_binding_matches_unmodified_reference_objects():
    from cv_diffusion_model_b200 import LowLightDiffusion, native, reference_backend as rb
    RefPipeline = _import_reference()
    torch.manual_seed(0)
    ref = RefPipeline(unet_variant="small", image_size=64, num_inference_steps=4)
    torch.manual_seed(0)
    mine = LowLightDiffusion(unet_variant="small", image_size=64, num_inference_steps=4)
    a, b = rb.config_struct(ref.unet.config), native.config_struct(mine.unet.config, "strict")
    for name, _ in rb.UNetConfigC._fields_:
        if name == "groupnorm_gcd":
            continue
        va, vb = getattr(a, name), getattr(b, name)
        assert (list(va) == list(vb)) if hasattr(va, "__len__") else (va == vb), name
    assert C.sizeof(a) == C.sizeof(b)
    # the weights `attach` uploads are exactly the reference's state_dict entries (same names, shapes and values)
    sd_ref, sd_my = ref.unet.state_dict(), mine.unet.state_dict()
    assert list(sd_ref) == list(sd_my)
    assert all(sd_ref[k].shape == sd_my[k].shape and torch.equal(sd_ref[k], sd_my[k]) for k in sd_ref)
    # coefficients from the reference's scheduler object == the package's scheduler
    ref.scheduler.set_timesteps(4)
    ts = [int(t) for t in ref.scheduler.timesteps]
    mine.scheduler.set_timesteps(4)
    want = [v for t in ts for v in mine.scheduler.step_coefficients(t)[:4]]
    assert rb.step_coefficients(ref.scheduler, ts) == want
    # without a GPU the binding fails loudly (no fallback)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            rb.attach(ref, 1)


class _TorchOnlyScheduler:
    """What `attach` touches of the reference's LCMScheduler (set_timesteps, timesteps, alphas_cumprod,
    final_alpha_cumprod), restated with the oracle's tables."""

    def __init__(self):
        from oracle import lcm_oracle
        self._o = lcm_oracle
        self.alphas_cumprod = lcm_oracle.alphas_cumprod()
        self.final_alpha_cumprod = self.alphas_cumprod[0]
        self.timesteps = None

    def set_timesteps(self, n, device="cpu"):
        self.timesteps = torch.tensor(self._o.timesteps(n), dtype=torch.long, device=device)


@pytest.mark.gpu
def test_attach_reproduces_package_enhance_bitwise():
    from cv_diffusion_model_b200 import LowLightDiffusion, reference_backend as rb
    from tests.util import seeded_unet
    size, b = 64, 2
    unet = seeded_unet("small", size, affine=True).cuda()
    model = types.SimpleNamespace(unet=unet, scheduler=_TorchOnlyScheduler(), image_size=size, num_inference_steps=4,
                                  condition_mode="concat")
    rb.attach(model, b)
    pipe = LowLightDiffusion(unet=seeded_unet("small", size, affine=True), image_size=size, num_inference_steps=4,
                             precision="bf16").cuda().eval()
    low = (torch.rand(b, 3, size, size, generator=torch.Generator().manual_seed(1234)) * 0.4 - 1).cuda()

    def run(fn):
        gen = torch.Generator(device="cuda").manual_seed(9)
        torch.manual_seed(5)
        return fn(low, generator=gen)

    got, want = run(model.enhance), run(pipe.enhance)
    assert torch.equal(got, want)
    x = torch.cat([torch.randn(b, 3, size, size, device="cuda"), low], dim=1)
    t = torch.tensor([739, 19], device="cuda")
    assert torch.equal(model.unet.forward(x, t), pipe.unet(x, t))
    model._b200.close()
