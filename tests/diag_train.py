#!/usr/bin/env python
"""GPU diagnostic (not a test): per-tensor error of the native backward pass against autograd of the oracle.

    python tests/diag_train.py [fp32|bf16] [variant] [cfg_image_size] [input] [batch]

Prints d loss / d activation for every tapped forward tensor (in backward execution order) and d loss / d weight for
every parameter: the first row with a large error localises a broken backward op.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cv_diffusion_model_b200 import LCMScheduler  # noqa: E402
from cv_diffusion_model_b200.training import TrainEngine  # noqa: E402
from oracle import lcm_oracle, train_oracle  # noqa: E402
from tests.util import rel_rms, seeded_unet  # noqa: E402


def main():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    variant = sys.argv[2] if len(sys.argv) > 2 else "small"
    cfg_size = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    size = int(sys.argv[4]) if len(sys.argv) > 4 else 64
    b = int(sys.argv[5]) if len(sys.argv) > 5 else 2
    patched = variant in ("tiny", "base")
    m = seeded_unet(variant, cfg_size, patched=patched, affine=True)
    g = torch.Generator().manual_seed(100 + size)
    high = torch.rand(b, 3, size, size, generator=g) * 2 - 1
    low = ((high + 1) / 2) ** 3 * 2 - 1
    t = torch.randint(0, 1000, (b,), generator=g)
    noise = torch.randn(b, 3, size, size, generator=g)
    abar = lcm_oracle.alphas_cumprod()
    dev = "cuda"
    sd = {k: v.detach().to(dev).clone().requires_grad_(True) for k, v in m.state_dict().items()}
    taps = {}

    def tap(k, v):
        k = k.rstrip(".")
        if v.requires_grad:
            v.retain_grad()
            taps[k] = v

    loss, eps = train_oracle.loss_fn(sd, m.config, abar, low.to(dev), high.to(dev), t.to(dev), noise.to(dev), "mse", not patched, tap=tap)
    loss.backward()
    eng = TrainEngine(m, b, size, size, precision=prec, taps=True, device=dev)
    sched = LCMScheduler(rescale_betas_zero_snr=True)
    noisy = sched.add_noise(high.cuda(), noise.cuda(), t.cuda())
    e = eng.forward(noisy, low.cuda(), t.cuda())
    l = eng.loss(e, noise.cuda()).item()
    eng.backward(noisy, low.cuda(), t.cuda(), e, noise.cuda())
    torch.cuda.synchronize()
    print(f"# {prec} {variant} cfg={cfg_size} in={size} b={b}: loss {l:.6f} vs oracle {loss.item():.6f}; eps rel_rms {rel_rms(e.cpu(), eps.detach().cpu()):.3e}")
    print(f"{'d loss / d activation':48s} {'rel_rms':>10s} {'ref_rms':>10s}")
    fwd_shapes = {}
    for name in reversed(list(taps)):
        ref = taps[name].grad
        if ref is None or ref.dim() != 4:
            continue
        try:
            got = eng.read_grad_tap(name, ref.shape[1], ref.shape[2], ref.shape[3])
        except ValueError:
            continue
        print(f"{name:48s} {rel_rms(got.cpu(), ref.cpu()):10.3e} {ref.double().pow(2).mean().sqrt().item():10.3e}")
    print(f"\n{'d loss / d weight (backward order)':56s} {'rel':>10s} {'ref_norm':>10s}")
    grads = eng.grads()
    order = sorted(eng.infos, key=lambda r: r[3])
    for name, _, _, ready in order:
        ref = sd[name].grad
        got = grads[name]
        d = (got.double() - ref.double()).norm().item() / max(ref.double().norm().item(), 1e-300)
        print(f"{name:56s} {d:10.3e} {ref.double().norm().item():10.3e}   (op {ready})")


if __name__ == "__main__":
    main()
