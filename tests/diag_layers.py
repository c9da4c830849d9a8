#!/usr/bin/env python
"""GPU diagnostic (not a test): per-layer error of the native path against the oracle's taps.

    python tests/diag_layers.py [fp32|bf16] [simt|tc] [variant] [cfg_image_size] [input] [batch]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cv_diffusion_model_b200.engine import Engine  # noqa: E402
from oracle import unet_oracle  # noqa: E402
from tests.util import rel_rms, seeded_unet  # noqa: E402


def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    simt = (sys.argv[2] if len(sys.argv) > 2 else "simt") == "simt"
    variant = sys.argv[3] if len(sys.argv) > 3 else "small"
    cfg_size = int(sys.argv[4]) if len(sys.argv) > 4 else 64
    size = int(sys.argv[5]) if len(sys.argv) > 5 else 64
    b = int(sys.argv[6]) if len(sys.argv) > 6 else 2
    patched = variant in ("tiny", "base")
    m = seeded_unet(variant, cfg_size, patched=patched, affine=True)
    torch.manual_seed(1)
    x = torch.randn(b, 6, size, size)
    t = torch.tensor([739, 19, 499, 259][:b])
    taps = {}
    with torch.no_grad():
        want = unet_oracle.unet_forward(m.state_dict(), m.config, x, t, strict_groupnorm=not patched,
                                        tap=lambda k, v: taps.__setitem__(k.rstrip("."), v))
    eng = Engine(m, b, size, size, precision=prec, simt_gemm=simt, taps=True, device="cuda")
    y = eng.forward(x.cuda(), t.cuda()).cpu()
    print(f"# {prec} {'simt' if simt else 'tc'} {variant} cfg={cfg_size} in={size} b={b}")
    print(f"{'tap':48s} {'rel_rms':>10s} {'max_abs':>10s} {'ref_max':>10s}")
    for name in eng.taps():
        if name not in taps:
            continue
        got = eng.read_tap(name).cpu()
        ref = taps[name]
        print(f"{name:48s} {rel_rms(got, ref):10.3e} {(got - ref).abs().max().item():10.3e} {ref.abs().max().item():10.3e}")
    print(f"{'eps':48s} {rel_rms(y, want):10.3e} {(y - want).abs().max().item():10.3e} {want.abs().max().item():10.3e}")


if __name__ == "__main__":
    main()
