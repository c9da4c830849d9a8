"""CPU: the oracle and the host-side mirrors against the golden vectors made from the
unmodified reference (tests/golden/make_golden.py) and SURVEY App. C known answers."""
import struct

import numpy as np
import pytest
import torch

from cv_diffusion_model_b200.scheduler import LCMScheduler, get_lcm_timesteps, lcm_timestep_list
from oracle import lcm_oracle, unet_oracle
from tests.util import randomise_affine, sd_digest, seeded_unet

UNET_CASES = [  # tag, variant, cfg image_size, input size, batch, patched, affine
    ("small256_in64", "small", 256, 64, 2, False, False),
    ("small128_in64", "small", 128, 64, 2, False, False),
    ("small256_in32_affine", "small", 256, 32, 2, False, True),
    ("small64_in64_affine", "small", 64, 64, 1, False, True),
    ("tiny256_in64_patched", "tiny", 256, 64, 2, True, True),
    ("base256_in32_patched", "base", 256, 32, 1, True, True),
]


def hexf(s):
    return struct.unpack("<f", struct.pack("<f", float.fromhex(s)))[0]


def test_schedules_exact(golden):
    assert lcm_oracle.timesteps(4) == [739, 499, 259, 19]
    assert lcm_oracle.timesteps(6) == [819, 659, 499, 339, 179, 19]
    assert lcm_oracle.timesteps(8) == [859, 739, 619, 499, 379, 259, 139, 19]
    s = LCMScheduler(rescale_betas_zero_snr=True)
    for n in (1, 2, 4, 6, 8, 10):
        want = golden[f"timesteps_{n}"].tolist()
        s.set_timesteps(n)
        assert s.timesteps.tolist() == want == lcm_oracle.timesteps(n) == lcm_timestep_list(n) == get_lcm_timesteps(n)
        assert s.timesteps.dtype == torch.long
        for i, t in enumerate(want):
            assert s._get_prev_timestep(t) == (want[i + 1] if i + 1 < len(want) else 0)


def test_abar_table_bit_exact(golden):
    abar = lcm_oracle.alphas_cumprod()
    assert np.array_equal(abar.numpy().view(np.uint32), golden["abar_bits"])
    assert np.array_equal(LCMScheduler(rescale_betas_zero_snr=True).alphas_cumprod.numpy().view(np.uint32),
                          golden["abar_bits"])
    assert np.array_equal(LCMScheduler().alphas_cumprod.numpy().view(np.uint32), golden["abar_norescale_bits"])
    # SURVEY App. C hex anchors
    for idx, hx in {0: "0x1.ff9096p-1", 19: "0x1.f646fep-1", 259: "0x1.460ca4p-1", 499: "0x1.f059eap-3",
                    739: "0x1.2f954ap-5", 859: "0x1.ca1224p-8", 998: "0x1.a69dfp-23"}.items():
        assert abar[idx].item() == hexf(hx), idx
    assert abar[999].item() == 0.0
    assert abs(lcm_oracle.alphas_cumprod(rescale_zero_snr=False)[739].item() - 0.06131751) < 1e-7


def test_unknown_names_raise():
    with pytest.raises(ValueError):
        LCMScheduler(beta_schedule="nope")
    with pytest.raises(ValueError):
        LCMScheduler(prediction_type="nope")
    from cv_diffusion_model_b200.modules import create_efficient_unet
    with pytest.raises(ValueError):
        create_efficient_unet("huge")
    with pytest.raises(ValueError):  # reference F1: tiny does not construct without the documented patch
        create_efficient_unet("tiny")


def test_step_kat(golden):
    abar = lcm_oracle.alphas_cumprod()
    smp, eps, nz = (torch.from_numpy(golden[k]) for k in ("step_sample", "step_eps", "step_noise"))
    prev, x0 = lcm_oracle.step(eps, 739, smp, [739, 499, 259, 19], abar, nz)
    assert torch.equal(prev, torch.from_numpy(golden["step_prev_739"]))
    assert torch.equal(x0, torch.from_numpy(golden["step_x0_739"]))
    np.testing.assert_allclose(prev[0, 0, 0, :3].numpy(), [0.50226676, -3.73667622, -1.36122000], rtol=0, atol=2e-6)
    np.testing.assert_allclose(x0[0, 0, 0, :3].numpy(), [1.92337728, -9.40831470, -2.14062238], rtol=0, atol=2e-6)
    last, x0l = lcm_oracle.step(eps, 19, smp, [739, 499, 259, 19], abar, None)
    assert last is x0l and torch.equal(last, torch.from_numpy(golden["step_prev_19"]))
    # host coefficients used by the CUDA step are the same fp32 scalars
    s = LCMScheduler(rescale_betas_zero_snr=True)
    s.set_timesteps(4)
    sb_t, sa_t, sa_p, sb_p, is_last = s.step_coefficients(739)
    assert not is_last and s.step_coefficients(19)[4]
    ref = (smp - sb_t * eps) / sa_t
    assert torch.allclose(ref, x0, rtol=0, atol=1e-6)


def test_sinusoidal_kat(golden):
    e = unet_oracle.sinusoidal_embedding(torch.tensor([739, 19, 0, 999]), 32)
    assert torch.equal(e, torch.from_numpy(golden["sin_emb_32"]))
    np.testing.assert_allclose(e[0, :3].numpy(), [-0.74801749, 0.63713479, 0.34883800], atol=1e-6)
    np.testing.assert_allclose(e[0, 16:19].numpy(), [-0.66367900, 0.77075237, 0.93718302], atol=1e-6)


@pytest.mark.parametrize("tag,variant,cfg_size,in_size,b,patched,affine", UNET_CASES)
def test_unet_oracle_matches_reference_golden(golden, weight_digests, tag, variant, cfg_size, in_size, b, patched, affine):
    m = seeded_unet(variant, cfg_size, patched, affine)
    sd = m.state_dict()
    assert sd_digest(sd) == weight_digests[tag], "random-init weights differ from the reference's"
    torch.manual_seed(1)
    x = torch.randn(b, 6, in_size, in_size)
    t = torch.from_numpy(golden[f"unet_{tag}_t"])
    with torch.no_grad():
        y = unet_oracle.unet_forward(sd, m.config, x, t, strict_groupnorm=not patched)
    want = torch.from_numpy(golden[f"unet_{tag}_y"])
    assert (y - want).abs().max().item() <= 1e-5
    if tag == "small256_in64":
        assert abs(y.double().sum().item() - 939.050226) < 2e-3
        assert abs(y.abs().mean().item() - 0.29126178) < 1e-6
    if tag == "small128_in64":
        assert abs(y.double().sum().item() + 3807.645018) < 5e-3


def test_state_dict_layout_small():
    m = seeded_unet("small", 256)
    sd = m.state_dict()
    assert len(sd) == 321 and sum(v.numel() for v in sd.values()) == 18008035
    for k in ("time_mlp.1.weight", "time_mlp.3.bias", "init_conv.weight", "encoder_blocks.1.0.skip.weight",
              "encoder_blocks.0.0.se.fc1.bias", "encoder_blocks.0.0.time_mlp.1.weight", "downsamplers.2.down.bias",
              "upsamplers.0.conv.weight", "mid_attn.to_qkv.weight", "mid_attn.to_out.0.weight",
              "mid_attn.to_out.1.bias", "decoder_blocks.3.2.project.weight", "final_norm.weight", "final_conv.bias"):
        assert k in sd, k
    assert tuple(sd["mid_attn.to_qkv.weight"].shape) == (384, 256, 1, 1)
    assert tuple(sd["decoder_blocks.0.0.expand.weight"].shape) == (2048, 512, 1, 1)


@pytest.mark.parametrize("tag,size,b,steps", [("small64", 64, 2, 4), ("small32_8step", 32, 1, 8)])
def test_enhance_oracle_matches_reference_golden(golden, weight_digests, tag, size, b, steps):
    m = seeded_unet("small", size, affine=True)
    sd = m.state_dict()
    assert sd_digest(sd) == weight_digests[f"enhance_{tag}"]
    low, lat0 = torch.from_numpy(golden[f"enh_{tag}_low"]), torch.from_numpy(golden[f"enh_{tag}_lat0"])
    noises = list(torch.from_numpy(golden[f"enh_{tag}_noises"]))
    out, trace = lcm_oracle.enhance(sd, m.config, low, lat0, noises, steps, return_all=True)
    assert (out - torch.from_numpy(golden[f"enh_{tag}_out"])).abs().max().item() <= 1e-5
    assert (trace[-1][1] - torch.from_numpy(golden[f"enh_{tag}_preclamp"])).abs().max().item() <= 1e-4


def test_image_io_oracle_matches_reference_golden():
    """oracle/image_io_oracle.py against the outputs of the unmodified reference's preprocess_image / postprocess_image
    (tests/golden/make_golden_image_io.py), bit-exact."""
    import os
    import numpy as np
    from oracle import image_io_oracle
    kat = np.load(os.path.join(os.path.dirname(__file__), "golden", "image_io_kat.npz"))
    pre = image_io_oracle.preprocess_u8(kat["rgb"])
    assert pre.dtype == np.float32 and np.array_equal(pre, kat["pre"])
    post = image_io_oracle.postprocess_u8(kat["y"])
    assert post.dtype == np.uint8 and np.array_equal(post, kat["post"])
    # cv2.resize (default INTER_LINEAR) as the reference calls it, outputs stored by the generator
    k = 0
    while f"rs{k}_src" in kat:
        src, dst = kat[f"rs{k}_src"], kat[f"rs{k}_dst"]
        assert np.array_equal(image_io_oracle.resize_bilinear_u8(src[None], dst.shape[0], dst.shape[1])[0], dst), k
        k += 1
    assert k >= 7
    # the byte grid: every value maps into [-1, 1] and comes back as itself or one below (truncation after two roundings)
    u = np.arange(256, dtype=np.uint8).reshape(1, 16, 16, 1).repeat(3, axis=3)
    x = image_io_oracle.preprocess_u8(u)
    assert x.min() == -1.0 and x.max() == 1.0
    back = image_io_oracle.postprocess_u8(x).astype(np.int32)
    assert ((u.astype(np.int32) - back) >= 0).all() and ((u.astype(np.int32) - back) <= 1).all()


def test_reference_checkpoint_format_roundtrip(tmp_path):
    """Checkpoints in the reference trainer's format (trainer.py:415-456): `model_state_dict` with the 321 `unet.*` keys,
    optional `ema_shadow` keyed by parameter name; a bare state_dict (scripts/benchmark.py:56) loads too."""
    from cv_diffusion_model_b200 import LowLightDiffusion
    torch.manual_seed(0)
    a = LowLightDiffusion(unet_variant="small", image_size=64)
    shadow = {k: v.detach() * 0.5 for k, v in a.named_parameters()}
    ckpt = a.reference_checkpoint(ema_shadow=shadow, epoch=3, global_step=120, best_val_loss=0.25)
    assert len(ckpt["model_state_dict"]) == 381 and all(k.startswith("unet.") for k in ckpt["model_state_dict"])
    path = tmp_path / "checkpoint_epoch_3.pt"
    torch.save(ckpt, path)
    torch.manual_seed(1)
    b = LowLightDiffusion(unet_variant="small", image_size=64)
    _, meta = b.load_reference_checkpoint(str(path))
    assert meta == {"epoch": 3, "global_step": 120, "best_val_loss": 0.25}
    assert all(torch.equal(p, q) for p, q in zip(a.parameters(), b.parameters()))
    b.load_reference_checkpoint(ckpt, use_ema=True)
    assert all(torch.equal(p * 0.5, q) for p, q in zip(a.parameters(), b.parameters()))
    b.load_reference_checkpoint(a.state_dict())            # bare state_dict
    assert all(torch.equal(p, q) for p, q in zip(a.parameters(), b.parameters()))
    with pytest.raises(ValueError):
        b.load_reference_checkpoint({"model_state_dict": a.state_dict()}, use_ema=True)
    with pytest.raises(RuntimeError):                        # torch's own strict-key error, like the reference
        b.load_reference_checkpoint({"model_state_dict": {k: v for k, v in list(a.state_dict().items())[:-1]}})
