"""-m gpu: the CUDA path through the reference-facing API / C ABI against the oracle on identical
weights, inputs, timesteps and injected noise.

Stated tolerances (SURVEY §8d calibration: torch autocast-bf16 vs fp32 of the reference itself gives
eps rel-RMS 1.7 %, final pre-clamp PSNR 35 dB):
  fp32 mode : eps max-abs <= 1e-3 ; final pre-clamp latents max-abs <= 5e-3
  bf16 mode : eps rel-RMS <= 3 %  ; final pre-clamp latents PSNR(peak = 2) >= 33 dB
  timesteps / schedules: exact.
"""
import os

import numpy as np
import pytest
import torch

from oracle import lcm_oracle, unet_oracle
from tests.util import psnr, rel_rms, sd_digest, seeded_unet

pytestmark = pytest.mark.gpu

UNET_CASES = [  # tag, variant, cfg image_size, input size, batch, patched, affine
    ("small256_in64", "small", 256, 64, 2, False, False),
    ("small128_in64", "small", 128, 64, 2, False, False),
    ("small256_in32_affine", "small", 256, 32, 2, False, True),
    ("small64_in64_affine", "small", 64, 64, 1, False, True),
    ("large256_in32", "large", 256, 32, 1, False, False),
    ("tiny256_in64_patched", "tiny", 256, 64, 2, True, True),
    ("base256_in32_patched", "base", 256, 32, 1, True, True),
    ("small256_in64_stdattn", "small", 256, 64, 2, False, True),     # use_linear_attention=False (softmax attention)
    ("small64_in32_stdattn", "small", 64, 32, 2, False, True),
]
MODES = [("fp32", False), ("bf16", True), ("bf16", False)]   # (precision, simt_gemm)
if os.environ.get("LCM_SKIP_TC"):
    MODES = MODES[:2]


def _run_unet(m, x, t, precision, simt):
    from cv_diffusion_model_b200.engine import Engine
    eng = Engine(m, x.shape[0], x.shape[2], x.shape[3], precision=precision, simt_gemm=simt, device="cuda")
    y = eng.forward(x.cuda(), t.cuda()).cpu()
    eng.close()
    return y


@pytest.mark.parametrize("precision,simt", MODES)
@pytest.mark.parametrize("tag,variant,cfg_size,in_size,b,patched,affine", UNET_CASES)
def test_unet_forward_vs_golden(golden, weight_digests, tag, variant, cfg_size, in_size, b, patched, affine, precision, simt):
    m = seeded_unet(variant, cfg_size, patched, affine, use_linear_attention=not tag.endswith("stdattn"))
    assert sd_digest(m.state_dict()) == weight_digests[tag]
    torch.manual_seed(1)
    x = torch.randn(b, 6, in_size, in_size)
    t = torch.from_numpy(golden[f"unet_{tag}_t"])
    want = torch.from_numpy(golden[f"unet_{tag}_y"])        # produced by the unmodified reference
    y = _run_unet(m, x, t, precision, simt)
    if precision == "fp32":
        assert (y - want).abs().max().item() <= 1e-3
    else:
        assert rel_rms(y, want) <= 0.03
        assert (y - want).abs().max().item() <= 0.15 * want.abs().max().item()


@pytest.mark.parametrize("precision,simt", MODES)
def test_unet_forward_vs_oracle_batch_and_rect(precision, simt):
    """Fresh seeded case on the GPU box: non-square input, per-sample timesteps, batch 3."""
    m = seeded_unet("small", 256, affine=True, seed=3)
    g = torch.Generator().manual_seed(21)
    x = torch.randn(3, 6, 32, 64, generator=g)
    t = torch.tensor([999, 0, 401])
    with torch.no_grad():
        want = unet_oracle.unet_forward(m.state_dict(), m.config, x, t)
    y = _run_unet(m, x, t, precision, simt)
    if precision == "fp32":
        assert (y - want).abs().max().item() <= 1e-3
    else:
        assert rel_rms(y, want) <= 0.03


def test_module_call_surface():
    """`unet(x, t)` on CUDA tensors goes through the native plan (default precision bf16)."""
    m = seeded_unet("small", 256).cuda()
    g = torch.Generator().manual_seed(2)
    x = torch.randn(1, 6, 32, 32, generator=g)
    t = torch.tensor([259])
    y = m(x.cuda(), t.cuda())
    assert y.shape == (1, 3, 32, 32) and y.is_cuda and y.dtype == torch.float32
    with torch.no_grad():
        want = unet_oracle.unet_forward({k: v.cpu() for k, v in m.state_dict().items()}, m.config, x, t)
    assert rel_rms(y.cpu(), want) <= 0.03
    with pytest.raises(RuntimeError):
        m(x, t)   # CPU tensors: no fallback


def test_scheduler_step_bit_exact(golden):
    from cv_diffusion_model_b200 import LCMScheduler
    s = LCMScheduler(rescale_betas_zero_snr=True)
    s.set_timesteps(4, device="cuda")
    assert s.timesteps.tolist() == [739, 499, 259, 19] and s.timesteps.is_cuda
    smp, eps, nz = (torch.from_numpy(golden[k]).cuda() for k in ("step_sample", "step_eps", "step_noise"))
    out = s.step(eps, 739, smp, noise=nz)
    assert np.array_equal(out.prev_sample.cpu().numpy(), golden["step_prev_739"])
    assert np.array_equal(out.pred_original_sample.cpu().numpy(), golden["step_x0_739"])
    prev, x0 = s.step(eps, 19, smp, return_dict=False)
    assert np.array_equal(prev.cpu().numpy(), golden["step_prev_19"]) and torch.equal(prev, x0)
    # add_noise / get_velocity
    x0 = torch.randn(3, 3, 8, 8, device="cuda")
    n = torch.randn_like(x0)
    t = torch.tensor([0, 499, 998], device="cuda")
    ab = s.alphas_cumprod.cuda()[t].view(-1, 1, 1, 1)
    assert torch.allclose(s.add_noise(x0, n, t), ab ** 0.5 * x0 + (1 - ab) ** 0.5 * n, atol=1e-6)
    assert torch.allclose(s.get_velocity(x0, n, t), ab ** 0.5 * n - (1 - ab) ** 0.5 * x0, atol=1e-6)


def _golden_pipe(golden, weight_digests, tag, size, steps, precision):
    from cv_diffusion_model_b200 import LowLightDiffusion
    from tests.util import randomise_affine
    torch.manual_seed(0)
    pipe = LowLightDiffusion(unet_variant="small", image_size=size, num_inference_steps=steps, precision=precision)
    randomise_affine(pipe.unet)
    assert sd_digest(pipe.unet.state_dict()) == weight_digests[f"enhance_{tag}"]
    low = torch.from_numpy(golden[f"enh_{tag}_low"])
    lat0 = torch.from_numpy(golden[f"enh_{tag}_lat0"])
    noises = torch.from_numpy(golden[f"enh_{tag}_noises"])
    return pipe, low, lat0, noises


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("tag,size,b,steps", [("small64", 64, 2, 4), ("small32_8step", 32, 1, 8)])
def test_enhance_vs_reference_golden(golden, weight_digests, tag, size, b, steps, precision):
    """Free-running loop against the unmodified reference's output (teacher-forced randomness only).

    bf16 note: with random-init weights the sampler is chaotic (x0 = (x_t - ..)/sqrt(abar_t) amplifies eps
    errors 5-12x per step).  The reference's OWN torch-autocast-bf16 run scores 33.3 dB on the 64x64 4-step
    case and -0.1 dB on the 32x32 8-step case (calibrated in the build container), so the free-running bf16
    gate (33 dB) is applied to the 4-step case only; the ill-conditioned 32x32 8-step case is checked step by step
    below and a well-conditioned 8-step loop in test_enhance_8step_free_running_well_conditioned."""
    pipe, low, lat0, noises = _golden_pipe(golden, weight_digests, tag, size, steps, precision)
    if precision == "bf16" and steps == 8:
        pytest.skip("free-running bf16 parity is not meaningful at 8 steps (see docstring); teacher-forced test below")
    pipe = pipe.cuda().eval()
    res = pipe.enhance(low.cuda(), latents=lat0.cuda(), noises=noises.cuda(), return_intermediate=True)
    assert pipe.scheduler.timesteps.tolist() == golden[f"timesteps_{steps}"].tolist()
    pre = res.intermediate[-1].cpu()
    want_pre = torch.from_numpy(golden[f"enh_{tag}_preclamp"])
    want = torch.from_numpy(golden[f"enh_{tag}_out"])
    assert torch.equal(res.enhanced.cpu(), pre.clamp(-1, 1))
    if precision == "fp32":
        assert (pre - want_pre).abs().max().item() <= 5e-3
        assert (res.enhanced.cpu() - want).abs().max().item() <= 5e-3
    else:
        # measured 34.5 dB (bitwise reproducible); the reference under torch's own bf16 autocast: 33.3 dB
        # (profiles/r02_loop_parity_bf16_vs_autocast.txt)
        assert psnr(pre, want_pre, 2.0) >= 33.0
        assert psnr(res.enhanced.cpu(), want, 2.0) >= 33.0


@pytest.mark.parametrize("tag,size,b,steps", [("small64", 64, 2, 4), ("small32_8step", 32, 1, 8)])
def test_enhance_bf16_teacher_forced_steps(golden, weight_digests, tag, size, b, steps):
    """Every step of the loop in bf16, fed with the oracle's latents of that step.

    Gate per step: eps rel-RMS <= 3 % for the well-conditioned 64x64 case.  The 32x32 / batch-1 case is
    ill-conditioned (GroupNorm runs over 4x4 maps; bf16 rounding anywhere upstream is amplified chaotically:
    torch-autocast-bf16 of the *reference itself* is 45 % off at t=859 and 1-2 % at the other steps, and
    two equally valid bf16 roundings of the same op differ by 2-5 % in eps).  There the gate is
    max(6 %, 1.5 x the error torch's own bf16 autocast of the oracle makes on the same input): being in the
    same class as PyTorch's bf16 is the meaningful statement.  The fused scheduler step must reproduce the
    oracle's update from that eps to 1e-5."""
    from cv_diffusion_model_b200.engine import Engine
    pipe, low, lat0, noises = _golden_pipe(golden, weight_digests, tag, size, steps, "bf16")
    sd = {k: v.clone() for k, v in pipe.unet.state_dict().items()}
    _, trace = lcm_oracle.enhance(sd, pipe.unet.config, low, lat0, list(noises), steps, return_all=True)
    sched = lcm_oracle.timesteps(steps)
    abar = lcm_oracle.alphas_cumprod()
    eng = Engine(pipe.unet, b, size, size, precision="bf16", device="cuda")
    lat = lat0
    for i, t in enumerate(sched):
        x = torch.cat([lat, low], dim=1)
        tt = torch.full((b,), t, dtype=torch.long)
        eps = eng.forward(x.cuda(), tt.cuda()).cpu()
        with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
            eps_autocast = unet_oracle.unet_forward(sd, pipe.unet.config, x, tt).float()
        gate = 0.03 if size >= 64 else max(0.06, 1.5 * rel_rms(eps_autocast, trace[i][0]))
        err = rel_rms(eps, trace[i][0])
        assert err <= gate, (i, t, err, gate, rel_rms(eps_autocast, trace[i][0]))
        nz = noises[i] if i < steps - 1 else None
        want_next, _ = lcm_oracle.step(eps, t, lat, sched, abar, nz)
        pipe.scheduler.set_timesteps(steps, device="cuda")
        got = pipe.scheduler.step(eps.cuda(), t, lat.cuda(), noise=None if nz is None else nz.cuda()).prev_sample.cpu()
        assert (got - want_next).abs().max().item() <= 1e-5 * max(1.0, want_next.abs().max().item())
        lat = trace[i][1]
    eng.close()


def test_enhance_reference_rng_protocol():
    """Without injection the draws follow the reference: latents from `generator`, step noise from the
    global RNG.  Same seeds => same output; the injected path reproduces it exactly."""
    from cv_diffusion_model_b200 import LowLightDiffusion
    torch.manual_seed(0)
    pipe = LowLightDiffusion(unet_variant="small", image_size=32).cuda().eval()
    low = (torch.rand(2, 3, 32, 32, generator=torch.Generator().manual_seed(1234)) * 0.4 - 1).cuda()

    def run():
        gen = torch.Generator(device="cuda").manual_seed(9)
        torch.manual_seed(5)
        return pipe.enhance(low, generator=gen)

    a, b_ = run(), run()
    assert torch.equal(a, b_)
    lat0 = torch.randn(2, 3, 32, 32, device="cuda", generator=torch.Generator(device="cuda").manual_seed(9))
    torch.manual_seed(5)
    noises = torch.stack([torch.randn_like(lat0) for _ in range(3)])
    c = pipe.enhance(low, latents=lat0, noises=noises)
    assert torch.equal(a, c)
    assert a.min().item() >= -1 and a.max().item() <= 1


def test_production_size_reproducible_and_parity():
    """BASELINE config[1] shape (Small, 256x256): images of 512 row tiles, CTAs that cross image boundaries and run
    several tiles ahead of their consumers — what the small cases above cannot exercise (a weight-reload race on the
    K = 160 project GEMMs was only visible here).  Bitwise run-to-run equality at batch 8, and the bf16 tolerances of
    this file against the oracle at batch 2 (teacher-forced first step + free-running 4-step PSNR)."""
    from cv_diffusion_model_b200 import LowLightDiffusion
    torch.manual_seed(0)
    pipe = LowLightDiffusion(unet_variant="small", image_size=256, num_inference_steps=4, precision="bf16")
    sd = {k[5:]: v.clone() for k, v in pipe.state_dict().items()}
    cfg = pipe.unet.config
    pipe = pipe.cuda().eval()
    g = torch.Generator().manual_seed(1234)
    low8 = torch.rand(8, 3, 256, 256, generator=g) * 0.4 - 1
    lat8 = torch.randn(8, 3, 256, 256, generator=torch.Generator().manual_seed(9))
    torch.manual_seed(5)
    noi8 = torch.stack([torch.randn(8, 3, 256, 256) for _ in range(3)])
    a = pipe.enhance(low8.cuda(), latents=lat8.cuda(), noises=noi8.cuda())
    for _ in range(3):   # the third call onwards replays the captured CUDA graph of the loop (engine.py)
        b_ = pipe.enhance(low8.cuda(), latents=lat8.cuda(), noises=noi8.cuda())
        assert torch.equal(a, b_)
    # parity at batch 2 (the oracle runs on the host cores: ~10 s)
    low, lat0, noises = low8[:2].contiguous(), lat8[:2].contiguous(), noi8[:, :2].contiguous()
    res = pipe.enhance(low.cuda(), latents=lat0.cuda(), noises=noises.cuda(), return_intermediate=True)
    want, trace = lcm_oracle.enhance(sd, cfg, low, lat0, list(noises), 4, return_all=True)
    t0 = torch.full((2,), int(pipe.scheduler._host_timesteps[0]), dtype=torch.long)
    eps = pipe.unet(torch.cat([lat0, low], dim=1).cuda(), t0.cuda()).cpu()     # first step's noise prediction
    assert rel_rms(eps, trace[0][0]) <= 0.03
    assert psnr(res.intermediate[-1].cpu(), trace[-1][1], 2.0) >= 33.0


@pytest.mark.parametrize("variant,size", [("base", 256), ("large", 128)])
def test_wider_variants_forward_parity(variant, size):
    """Base (48-channel stem: K = 192 / 384 expands on the activation-stationary kernel, K = 96 on the resident-weight
    one) and Large (64-channel stem: K = 512 expands and everything above on the general kernel) at sizes where every
    level has whole 128-pixel tiles; one forward against the oracle, bf16 tolerances of this file, run twice bitwise."""
    m = seeded_unet(variant, 512, patched=(variant == "base"), affine=True, seed=11)   # 48 channels need gcd groups
    g = torch.Generator().manual_seed(77)
    x = torch.randn(1, 6, size, size, generator=g)
    t = torch.tensor([617])
    with torch.no_grad():
        want = unet_oracle.unet_forward(m.state_dict(), m.config, x, t, strict_groupnorm=(variant != "base"))
    y = _run_unet(m, x, t, "bf16", False)
    y2 = _run_unet(m, x, t, "bf16", False)
    assert torch.equal(y, y2)
    assert rel_rms(y, want) <= 0.03
    assert (y - want).abs().max().item() <= 0.15 * want.abs().max().item()


def test_enhance_uint8_surface():
    """uint8 HWC in / out around enhance: equals preprocess -> enhance -> postprocess of the oracle formats (bit-exact
    given the same enhance result), values stay bytes."""
    import numpy as np
    from cv_diffusion_model_b200 import LowLightDiffusion
    from oracle import image_io_oracle
    torch.manual_seed(0)
    pipe = LowLightDiffusion(unet_variant="small", image_size=64, num_inference_steps=4, precision="bf16").cuda().eval()
    rgb = torch.randint(0, 60, (2, 64, 64, 3), dtype=torch.uint8)            # a dark image
    torch.manual_seed(3)
    out = pipe.enhance_uint8(rgb.cuda())
    assert out.dtype == torch.uint8 and tuple(out.shape) == (2, 64, 64, 3)
    torch.manual_seed(3)                                                      # same RNG protocol -> same latents / noises
    x = torch.from_numpy(image_io_oracle.preprocess_u8(rgb.numpy())).cuda()
    y = pipe.enhance(x).cpu().numpy()
    assert np.array_equal(out.cpu().numpy(), image_io_oracle.postprocess_u8(y))
    # with the reference's resize to the model size and back (scripts/inference.py:109,130)
    big = torch.randint(0, 60, (1, 90, 120, 3), dtype=torch.uint8)
    torch.manual_seed(4)
    out2 = pipe.enhance_uint8(big.cuda(), target_size=64)
    assert out2.dtype == torch.uint8 and tuple(out2.shape) == (1, 90, 120, 3)
    torch.manual_seed(4)
    x2 = image_io_oracle.preprocess_u8(image_io_oracle.resize_bilinear_u8(big.numpy(), 64, 64))
    y2 = image_io_oracle.postprocess_u8(pipe.enhance(torch.from_numpy(x2).cuda()).cpu().numpy())
    assert np.array_equal(out2.cpu().numpy(), image_io_oracle.resize_bilinear_u8(y2, 90, 120))


# ---- parity at the shapes BASELINE.json names (configs 2, 3, 4) -------------------------------------------------
def _autocast_psnr(sd, cfg, low, lat0, noises, steps, ref_pre, strict=True):
    """PSNR the reference's own arithmetic reaches when PyTorch runs it under bf16 autocast (CPU) on the same inputs:
    the yardstick for what 'bf16' can mean on a chaotic sampler with random-init weights."""
    with torch.autocast("cpu", dtype=torch.bfloat16):
        _, tr = lcm_oracle.enhance(sd, cfg, low, lat0, list(noises), steps, strict_groupnorm=strict, return_all=True)
    return psnr(tr[-1][1].float(), ref_pre, 2.0)


def test_config2_small256_batch64_parity():
    """BASELINE config[1] at its full batch: 64 images on the GPU (every CTA schedule, tile count and image-boundary
    crossing of the benchmark), the oracle on images {0, 31, 63} (images are independent).  Gates: teacher-forced first
    step eps rel-RMS <= 3 %, free-running 4-step pre-clamp PSNR >= 33 dB (or the reference's own bf16 autocast, if lower)."""
    from cv_diffusion_model_b200 import LowLightDiffusion
    B, S, pick = 64, 256, [0, 31, 63]
    torch.manual_seed(0)
    pipe = LowLightDiffusion(unet_variant="small", image_size=S, num_inference_steps=4, precision="bf16")
    sd = {k[5:]: v.clone() for k, v in pipe.state_dict().items()}
    cfg = pipe.unet.config
    pipe = pipe.cuda().eval()
    low = torch.rand(B, 3, S, S, generator=torch.Generator().manual_seed(1234)) * 0.4 - 1
    lat0 = torch.randn(B, 3, S, S, generator=torch.Generator().manual_seed(9))
    noises = torch.randn(3, B, 3, S, S, generator=torch.Generator().manual_seed(5))
    res = pipe.enhance(low.cuda(), latents=lat0.cuda(), noises=noises.cuda(), return_intermediate=True)
    t0 = torch.full((B,), int(pipe.scheduler._host_timesteps[0]), dtype=torch.long)
    eps = pipe.unet(torch.cat([lat0, low], dim=1).cuda(), t0.cuda()).cpu()
    want, trace = lcm_oracle.enhance(sd, cfg, low[pick], lat0[pick], list(noises[:, pick]), 4, return_all=True)
    assert rel_rms(eps[pick], trace[0][0]) <= 0.03
    got = psnr(res.intermediate[-1].cpu()[pick], trace[-1][1], 2.0)
    ref_bf16 = _autocast_psnr(sd, cfg, low[pick], lat0[pick], noises[:, pick], 4, trace[-1][1])
    print(f"config2 B=64: eps rel-RMS {rel_rms(eps[pick], trace[0][0]):.4f}, 4-step PSNR {got:.1f} dB "
          f"(reference under torch bf16 autocast: {ref_bf16:.1f} dB)")
    assert got >= min(33.0, ref_bf16), (got, ref_bf16)


def test_config3_base512_8step_teacher_forced():
    """BASELINE config[2] shape: Base* (gcd-GroupNorm patch) at 512x512, 32 images per GPU, 8 LCM steps.  Every step is
    fed the oracle's latents of that step (image 0; the other 31 images carry different noise through the same launch),
    eps rel-RMS <= 3 % per step and the scheduler update reproduces the oracle's."""
    from cv_diffusion_model_b200.engine import Engine
    B, S, steps = 32, 512, 8
    m = seeded_unet("base", S, patched=True, affine=True, seed=11)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    low = torch.rand(B, 3, S, S, generator=torch.Generator().manual_seed(1234)) * 0.4 - 1
    lat0 = torch.randn(B, 3, S, S, generator=torch.Generator().manual_seed(9))
    noises = torch.randn(steps - 1, 1, 3, S, S, generator=torch.Generator().manual_seed(5))
    _, trace = lcm_oracle.enhance(sd, m.config, low[:1], lat0[:1], list(noises), steps, strict_groupnorm=False,
                                  return_all=True)
    sched = lcm_oracle.timesteps(steps)
    assert sched == [859, 739, 619, 499, 379, 259, 139, 19]
    eng = Engine(m, B, S, S, precision="bf16", device="cuda")
    lat = lat0.clone()
    errs = []
    for i, t in enumerate(sched):
        tt = torch.full((B,), t, dtype=torch.long)
        eps = eng.forward(torch.cat([lat, low], dim=1).cuda(), tt.cuda()).cpu()
        errs.append(rel_rms(eps[:1], trace[i][0]))
        lat = lat.clone()
        lat[:1] = trace[i][1]          # teacher forcing: next step starts from the oracle's latents
    eng.close()
    print("config3 base512 8-step teacher-forced eps rel-RMS per step:", [f"{e:.4f}" for e in errs])
    assert max(errs) <= 0.03, errs


def test_config4_large1024_forward():
    """BASELINE config[3] shape: Large at 1024x1024, 8 images per GPU: one forward (n = 16 384 linear attention, 30 blocks,
    K up to 4096) against the oracle on image 0."""
    from cv_diffusion_model_b200.engine import Engine
    B, S = 8, 1024
    m = seeded_unet("large", S, affine=True, seed=11)
    x = torch.randn(B, 6, S, S, generator=torch.Generator().manual_seed(77))
    t = torch.tensor([739, 19, 499, 259, 617, 0, 999, 380])
    with torch.no_grad():
        want = unet_oracle.unet_forward(m.state_dict(), m.config, x[:1], t[:1])
    eng = Engine(m, B, S, S, precision="bf16", device="cuda")
    y = eng.forward(x.cuda(), t.cuda()).cpu()
    eng.close()
    err = rel_rms(y[:1], want)
    print(f"config4 large1024 B=8: eps rel-RMS {err:.4f}, max-abs {(y[:1] - want).abs().max().item():.4f} of {want.abs().max().item():.3f}")
    assert err <= 0.03
    assert (y[:1] - want).abs().max().item() <= 0.15 * want.abs().max().item()


def test_enhance_8step_free_running_well_conditioned():
    """8-step loop (the step count of config 3) on a well-conditioned case (64x64, batch 2), free-running in bf16: the gate is
    the PSNR the reference's own bf16 autocast reaches on the same inputs minus 1 dB, capped at 33 dB."""
    from cv_diffusion_model_b200 import LowLightDiffusion
    from tests.util import randomise_affine
    S, B, steps = 64, 2, 8
    torch.manual_seed(0)
    pipe = LowLightDiffusion(unet_variant="small", image_size=S, num_inference_steps=steps, precision="bf16")
    randomise_affine(pipe.unet)
    sd = {k[5:]: v.clone() for k, v in pipe.state_dict().items()}
    cfg = pipe.unet.config
    pipe = pipe.cuda().eval()
    low = torch.rand(B, 3, S, S, generator=torch.Generator().manual_seed(1234)) * 0.4 - 1
    lat0 = torch.randn(B, 3, S, S, generator=torch.Generator().manual_seed(9))
    noises = torch.randn(steps - 1, B, 3, S, S, generator=torch.Generator().manual_seed(5))
    res = pipe.enhance(low.cuda(), latents=lat0.cuda(), noises=noises.cuda(), return_intermediate=True)
    _, trace = lcm_oracle.enhance(sd, cfg, low, lat0, list(noises), steps, return_all=True)
    got = psnr(res.intermediate[-1].cpu(), trace[-1][1], 2.0)
    ref_bf16 = _autocast_psnr(sd, cfg, low, lat0, noises, steps, trace[-1][1])
    print(f"8-step 64x64 free-running: PSNR {got:.1f} dB (reference under torch bf16 autocast: {ref_bf16:.1f} dB)")
    assert got >= min(33.0, ref_bf16 - 1.0), (got, ref_bf16)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_engines_on_two_devices_in_one_process():
    """Function attributes (dynamic shared memory opt-in) and occupancy caches are per device."""
    from cv_diffusion_model_b200.engine import Engine
    m = seeded_unet("small", 256, affine=True)
    x = torch.randn(2, 6, 128, 128, generator=torch.Generator().manual_seed(3))
    t = torch.tensor([739, 19])
    outs = []
    for d in (0, 1):
        before = torch.cuda.current_device()
        eng = Engine(m, 2, 128, 128, precision="bf16", device=f"cuda:{d}")
        assert torch.cuda.current_device() == before        # lcm_plan_create restores the caller's device
        outs.append(eng.forward(x.to(f"cuda:{d}"), t.to(f"cuda:{d}")).cpu())
        eng.close()
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_enhance_add_conditioning_vs_reference_golden(golden, weight_digests, precision):
    """condition_mode="add" (low_light_diffusion.py:108-113,223-225): 3-channel UNet on latents + condition_encoder(low_light),
    free-running 4-step loop against the unmodified reference's output."""
    from cv_diffusion_model_b200 import LowLightDiffusion
    from tests.util import randomise_affine
    torch.manual_seed(0)
    pipe = LowLightDiffusion(unet_variant="small", image_size=64, num_inference_steps=4, condition_mode="add", precision=precision)
    randomise_affine(pipe.unet)
    assert sd_digest(pipe.state_dict()) == weight_digests["enhance_add64"]
    assert pipe.unet.config.in_channels == 3 and "condition_encoder.2.bias" in pipe.state_dict()
    low, lat0 = torch.from_numpy(golden["enh_add64_low"]), torch.from_numpy(golden["enh_add64_lat0"])
    noises = torch.from_numpy(golden["enh_add64_noises"])
    pipe = pipe.cuda().eval()
    res = pipe.enhance(low.cuda(), latents=lat0.cuda(), noises=noises.cuda(), return_intermediate=True)
    pre, want_pre = res.intermediate[-1].cpu(), torch.from_numpy(golden["enh_add64_preclamp"])
    assert torch.equal(res.enhanced.cpu(), pre.clamp(-1, 1))
    if precision == "fp32":
        assert (pre - want_pre).abs().max().item() <= 5e-3
        assert (res.enhanced.cpu() - torch.from_numpy(golden["enh_add64_out"])).abs().max().item() <= 5e-3
    else:
        sd = {k[5:]: v.cpu() for k, v in pipe.state_dict().items() if k.startswith("unet.")}
        enc = {k[len("condition_encoder."):]: v.cpu() for k, v in pipe.state_dict().items() if k.startswith("condition_encoder.")}
        with torch.autocast("cpu", dtype=torch.bfloat16):
            _, tr = lcm_oracle.enhance(sd, pipe.unet.config, low, lat0, list(noises), 4, return_all=True, condition_encoder_sd=enc)
        ref_bf16 = psnr(tr[-1][1].float(), want_pre, 2.0)
        got = psnr(pre, want_pre, 2.0)
        print(f"add conditioning bf16: PSNR {got:.1f} dB (reference under torch bf16 autocast: {ref_bf16:.1f} dB)")
        assert got >= min(33.0, ref_bf16 - 1.0)
    # training-style forward without autograd works; with autograd it states what is missing
    pipe.train()
    with torch.no_grad():
        out = pipe(low.cuda(), lat0.cuda().clamp(-1, 1), timesteps=torch.tensor([10, 700]).cuda())
    assert out["noise_pred"].shape == (2, 3, 64, 64)
    with pytest.raises(NotImplementedError):
        pipe.compute_loss(low.cuda(), lat0.cuda().clamp(-1, 1))


def test_fp16_hidden_range_report_and_saturation():
    """The bf16 plan keeps the blocks' hidden tensors in fp16 (cvt.rn.satfinite).  `hidden_range_report` (fp32 plan + taps)
    must clear the default-init model with a wide margin, must flag a checkpoint whose expand weights are blown up past the
    fp16 range, and the bf16 path must stay finite on that checkpoint (saturation, never inf / nan)."""
    from cv_diffusion_model_b200.engine import Engine, hidden_range_report
    from tests.util import seeded_unet
    m = seeded_unet("small", 64, affine=True)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 6, 64, 64, generator=g).cuda()
    t = torch.tensor([739, 19], device="cuda")
    rep = hidden_range_report(m, x, t)
    assert rep["fits_fp16"] and rep["headroom"] > 100, rep["worst"]
    with torch.no_grad():
        m.encoder_blocks[0][0].expand.weight.mul_(3e5)
    rep2 = hidden_range_report(m, x, t)
    assert not rep2["fits_fp16"] and rep2["worst"][0].startswith("encoder_blocks.0.0"), rep2["worst"]
    eng = Engine(m, 2, 64, 64, precision="bf16", device="cuda")
    y = eng.forward(x, t)
    eng.close()
    assert torch.isfinite(y).all()


@pytest.mark.parametrize("precision,simt,tol", [("fp32", True, 1e-5), ("bf16", False, 2e-2)])
def test_linear_attention_taps_vs_oracle(precision, simt, tol):
    """The linear-attention kernels (attn_kv / attn_apply, efficient_unet.py:262-309) at every place the Small variant has
    them (two encoder, the mid and three decoder modules at config image_size 128): the tapped module outputs against the
    oracle's taps, per module.  Stated tolerance: rel-RMS 1e-5 in fp32 (measured 3e-7), 2 % in bf16 (measured 0.3-0.65 %: the
    accumulated error of all layers before)."""
    from cv_diffusion_model_b200.engine import Engine
    if precision == "bf16" and os.environ.get("LCM_SKIP_TC"):
        pytest.skip("LCM_SKIP_TC set")
    m = seeded_unet("small", 128, patched=False, affine=True)
    torch.manual_seed(11)
    b, size = 2, 64
    x = torch.randn(b, 6, size, size)
    t = torch.tensor([739, 19])
    want = {}
    with torch.no_grad():
        unet_oracle.unet_forward(m.state_dict(), m.config, x, t, strict_groupnorm=True,
                                 tap=lambda k, v: want.__setitem__(k.rstrip("."), v))
    eng = Engine(m, b, size, size, precision=precision, simt_gemm=simt, taps=True, device="cuda")
    eng.forward(x.cuda(), t.cuda())
    names = [n for n in eng.taps() if n.endswith(".attn") and n in want]
    assert len(names) >= 6, names            # 2 encoder + mid + 3 decoder attention modules
    worst = {}
    for n in names:
        got = eng.read_tap(n).cpu()
        assert got.shape == want[n].shape
        worst[n] = rel_rms(got, want[n])
    eng.close()
    print({k: f"{v:.2e}" for k, v in worst.items()})
    assert max(worst.values()) <= tol, worst
