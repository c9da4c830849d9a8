#!/usr/bin/env python
"""Golden fixture of ONE training step, produced by the UNMODIFIED reference (build container only).

Protocol (reference code paths: LowLightDiffusion.forward/compute_loss low_light_diffusion.py:115-175,250-277;
LowLightTrainer.train_epoch's non-AMP branch trainer.py:303-316; EMAModel.update :98-104):
  torch.manual_seed(0); model = LowLightDiffusion("small", image_size=64); affine parameters randomised like the
  inference goldens; two batches of 2 synthetic pairs (high = rand*2-1, low = ((high+1)/2)^3*2-1); per step the
  timesteps and the noise are drawn explicitly and passed in (`forward(..., timesteps=, noise=)`), then
  loss.backward(); clip_grad_norm_(1.0); AdamW(lr 1e-4, wd 0.01).step(); ema.update().
Stored: inputs, t, noise, the two losses, the pre-clip gradient norm, per-parameter gradient norms and checksums of step 1,
per-parameter norms of the weight change after the two steps and of the EMA shadow change.

    python tests/golden/make_golden_train.py
"""
import hashlib
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle", "ref_shim"), "/root/reference"]

from src.models.low_light_diffusion import LowLightDiffusion as RefPipeline  # noqa: E402

from oracle import train_oracle  # noqa: E402
from tests.golden.make_golden import randomise_affine, sd_digest  # noqa: E402

torch.set_num_threads(os.cpu_count())


def load_ema_class():
    """EMAModel from the reference trainer, loaded by file path (src/training/__init__ pulls in albumentations)."""
    import importlib.util
    import types
    for name in ("wandb", "tqdm"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = types.ModuleType(name)
                m.tqdm = lambda x, **k: x
                sys.modules[name] = m
    pkg = types.ModuleType("src.training")
    pkg.__path__ = []
    sys.modules.setdefault("src.training", pkg)
    ds = types.ModuleType("src.training.dataset")
    sys.modules.setdefault("src.training.dataset", ds)
    spec = importlib.util.spec_from_file_location("src.training.trainer", "/root/reference/src/training/trainer.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.EMAModel


def distill(out):
    """One consistency-distillation loss + backward of the unmodified reference wrapper (low_light_diffusion.py:284-408)."""
    from src.models.low_light_diffusion import LowLightLCMDistillation as RefDistill
    S, B = 64, 2
    torch.manual_seed(0)
    teacher = RefPipeline(unet_variant="small", image_size=S, num_inference_steps=4)
    randomise_affine(teacher.unet)
    torch.manual_seed(1)
    student = RefPipeline(unet_variant="small", image_size=S, num_inference_steps=4)
    randomise_affine(student.unet, seed=2)
    wrap = RefDistill(teacher, student)
    with torch.no_grad():      # make the EMA target differ from the student, like after some updates
        for e in wrap.ema_student.parameters():
            e.mul_(0.98)
    g = torch.Generator().manual_seed(555)
    high = torch.rand(B, 3, S, S, generator=g) * 2 - 1
    low = ((high + 1) / 2) ** 3 * 2 - 1
    torch.manual_seed(99)
    loss = wrap.consistency_distillation_loss(low, high, num_inference_steps=4)
    loss.backward()
    torch.manual_seed(99)      # the same two draws, made explicitly (randn_like, then randint: :341,348)
    noise = torch.randn(B, 3, S, S)
    idx = torch.randint(0, 50 - 50 // 4, (B,))
    ssd = {k: v.detach().clone().requires_grad_(True) for k, v in student.unet.state_dict().items()}
    o = train_oracle.distillation_loss(teacher.unet.state_dict(), ssd, wrap.ema_student.unet.state_dict(), student.unet.config,
                                       teacher.scheduler.alphas_cumprod, low, high, noise, idx)
    assert abs(o.item() - loss.item()) <= 1e-6 * abs(loss.item()), (o.item(), loss.item())
    og = torch.autograd.grad(o, list(ssd.values()))
    names = [n for n, _ in student.unet.named_parameters()]
    worst = max((student.unet.get_parameter(n).grad - gk).abs().max().item() / (gk.abs().max().item() + 1e-30) for n, gk in zip(names, og))
    print(f"distillation: loss {loss.item():.6f}, oracle vs reference gradients worst relative max-diff {worst:.2e}")
    assert worst <= 1e-4
    out.update(distill_low=low.numpy(), distill_high=high.numpy(), distill_noise=noise.numpy(), distill_idx=idx.numpy(),
               distill_loss=np.array(loss.item()),
               distill_grad_norms=np.array([student.unet.get_parameter(n).grad.norm().item() for n in names], dtype=np.float64))


def main():
    S, B, steps = 64, 2, 2
    torch.manual_seed(0)
    model = RefPipeline(unet_variant="small", image_size=S, num_inference_steps=4)
    randomise_affine(model.unet)
    model.train()
    sd0 = {k: v.clone() for k, v in model.unet.state_dict().items()}
    EMAModel = load_ema_class()
    ema = EMAModel(model, decay=0.9999)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0.01)
    g = torch.Generator().manual_seed(321)
    out = {}
    losses, gnorms = [], []
    for step in range(steps):
        high = torch.rand(B, 3, S, S, generator=g) * 2 - 1
        low = ((high + 1) / 2) ** 3 * 2 - 1
        t = torch.randint(0, 1000, (B,), generator=g)
        noise = torch.randn(B, 3, S, S, generator=g)
        opt.zero_grad()
        res = model(low, high, timesteps=t, noise=noise)
        loss = F.mse_loss(res["noise_pred"], res["noise"])          # compute_loss, loss_type="mse"
        loss.backward()
        if step == 0:
            # the oracle's autograd must reproduce the reference's gradients exactly (same ATen ops)
            sd = {k: v.detach().clone().requires_grad_(True) for k, v in sd0.items()}
            o_loss, o_grads = train_oracle.loss_and_grads(sd, model.unet.config, model.scheduler.alphas_cumprod, low, high, t, noise)
            assert abs(o_loss - loss.item()) <= 1e-6 * abs(loss.item()), (o_loss, loss.item())
            worst = max((p.grad - o_grads[n]).abs().max().item() / (p.grad.abs().max().item() + 1e-30)
                        for n, p in model.unet.named_parameters())
            print(f"oracle autograd vs reference gradients: worst relative max-diff {worst:.2e}")
            assert worst <= 1e-4
            names = [n for n, _ in model.unet.named_parameters()]
            out["grad_norms"] = np.array([model.unet.get_parameter(n).grad.norm().item() for n in names], dtype=np.float64)
            probe = torch.Generator().manual_seed(7)
            out["grad_probe"] = np.array([(model.unet.get_parameter(n).grad.flatten() *
                                           torch.randn(model.unet.get_parameter(n).numel(), generator=probe)).sum().item()
                                          for n in names], dtype=np.float64)
            out["names"] = np.array(names)
        gn = torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        ema.update(model)
        losses.append(loss.item())
        gnorms.append(gn.item())
        out[f"low_{step}"], out[f"high_{step}"] = low.numpy(), high.numpy()
        out[f"t_{step}"], out[f"noise_{step}"] = t.numpy(), noise.numpy()
        print(f"step {step}: loss {loss.item():.6f}  grad norm {gn.item():.6f}")
    names = [n for n, _ in model.unet.named_parameters()]
    sd1 = model.unet.state_dict()
    out["losses"] = np.array(losses, dtype=np.float64)
    out["grad_total_norms"] = np.array(gnorms, dtype=np.float64)
    out["delta_norms"] = np.array([(sd1[n] - sd0[n]).norm().item() for n in names], dtype=np.float64)
    out["ema_delta_norms"] = np.array([(ema.shadow["unet." + n] - sd0[n]).norm().item() for n in names], dtype=np.float64)
    out["weights_sha256"] = np.array(sd_digest(sd0))
    distill(out)
    np.savez_compressed(os.path.join(HERE, "train_kat.npz"), **out)
    print("wrote train_kat.npz", os.path.getsize(os.path.join(HERE, "train_kat.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
