#!/usr/bin/env python
"""GPU diagnostic: per-op device times of one UNet forward (CUDA events around every launch) with the
algorithmic bytes / FLOPs of each op, grouped by kernel.

    python tests/prof_model.py [variant] [size] [batch] [precision] [simt|tc]
"""
import json
import os
import sys
from collections import defaultdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cv_diffusion_model_b200.engine import Engine  # noqa: E402
from tests.util import seeded_unet  # noqa: E402


def main():
    variant = sys.argv[1] if len(sys.argv) > 1 else "small"
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    b = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    prec = sys.argv[4] if len(sys.argv) > 4 else "bf16"
    simt = (sys.argv[5] if len(sys.argv) > 5 else "tc") == "simt"
    m = seeded_unet(variant, size, patched=variant in ("tiny", "base"))
    eng = Engine(m, b, size, size, precision=prec, simt_gemm=simt, device="cuda")
    x = torch.randn(b, 6, size, size, device="cuda")
    t = torch.full((b,), 499, device="cuda", dtype=torch.long)
    for _ in range(2):
        eng.forward(x, t)
    torch.cuda.synchronize()
    recs = eng.profile(x, t)
    recs = eng.profile(x, t)
    tot = sum(r["ms"] for r in recs)
    print(f"# {variant}@{size} B={b} {prec} {'simt' if simt else 'tc'}: {len(recs)} ops, {tot:.3f} ms per forward (event-timed per op), "
          f"workspace {eng.workspace.numel()/2**30:.2f} GiB, algorithmic {eng.algorithmic_bytes/1e9:.2f} GB {eng.algorithmic_flops/1e12:.2f} TFLOP")
    print(f"{'op':44s} {'kernel':16s} {'ms':>8s} {'GB/s':>8s} {'TFLOP/s':>8s}")
    for r in recs:
        if r["ms"] > 0.02:
            print(f"{r['name']:44s} {r['kernel']:16s} {r['ms']:8.3f} {r['bytes']/r['ms']/1e6:8.0f} {r['flops']/r['ms']/1e9:8.1f}")
    agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for r in recs:
        a = agg[r["kernel"]]
        a[0] += 1; a[1] += r["ms"]; a[2] += r["bytes"]; a[3] += r["flops"]
    print("\n# by kernel")
    for k, (n, ms, by, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:16s} n={n:3d} {ms:8.3f} ms {100*ms/tot:5.1f}%  {by/ms/1e6 if ms else 0:8.0f} GB/s {fl/ms/1e9 if ms else 0:8.1f} TFLOP/s")
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/prof_model_{variant}{size}_b{b}_{prec}.json", "w") as f:
        json.dump(recs, f)


if __name__ == "__main__":
    main()
