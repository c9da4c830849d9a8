#!/usr/bin/env python
"""GPU profile (not a test) of the native training step: step time, forward / backward / optimizer split, backward time by
kernel label (CUDA events around every backward op).

    python tests/prof_train.py [bf16|fp32] [batch] [size] [variant]
"""
import os
import sys
from collections import defaultdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cv_diffusion_model_b200 import LowLightDiffusion  # noqa: E402
from cv_diffusion_model_b200.training import NativeTrainer  # noqa: E402


def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    S = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    variant = sys.argv[4] if len(sys.argv) > 4 else "small"
    torch.manual_seed(0)
    pipe = LowLightDiffusion(unet_variant=variant, image_size=S, groupnorm="gcd" if variant in ("tiny", "base") else "strict",
                             precision=prec).cuda().train()
    tr = NativeTrainer(pipe, batch=B, precision=prec)
    eng = tr.engine
    g = torch.Generator().manual_seed(1)
    high = (torch.rand(B, 3, S, S, generator=g) * 2 - 1).cuda()
    low = ((high + 1) / 2) ** 3 * 2 - 1
    print(f"# {prec} {variant}@{S} B={B}: workspace {eng.workspace.numel() / 1e9:.1f} GB, {eng.num_backward_ops} backward ops")

    def timed(fn, n=3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    losses = [tr.train_step(low, high).item() for _ in range(3)]
    print("losses of the first steps:", [f"{l:.5f}" for l in losses])
    ms = timed(lambda: tr.train_step(low, high))
    print(f"train_step: {ms:.2f} ms -> {B / ms * 1e3:.1f} images/s")
    t = torch.randint(0, 1000, (B,), device="cuda")
    noise = torch.randn_like(high)
    noisy = pipe.scheduler.add_noise(high, noise, t)
    eps = eng.forward(noisy, low, t)
    fwd = timed(lambda: eng.forward(noisy, low, t))
    bwd = timed(lambda: eng.backward(noisy, low, t, eps, noise))
    up = timed(lambda: eng.upload_weights())
    print(f"forward {fwd:.2f} ms, backward {bwd:.2f} ms, weight re-pack {up:.2f} ms")
    ops = eng.backward_ops()
    by = defaultdict(lambda: [0.0, 0])
    rows = []
    eng.backward(noisy, low, t, eps, noise, op_begin=0, op_end=0)
    for i, (name, kern) in enumerate(ops):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.backward(noisy, low, t, eps, noise, op_begin=i, op_end=i + 1)
        e1.record()
        torch.cuda.synchronize()
        dt = e0.elapsed_time(e1)
        by[kern][0] += dt
        by[kern][1] += 1
        rows.append((dt, name, kern))
    tot = sum(v[0] for v in by.values())
    print(f"backward by kernel (sum {tot:.2f} ms):")
    for k, (msk, n) in sorted(by.items(), key=lambda kv: -kv[1][0]):
        print(f"  {k:22s} {msk:8.3f} ms  {n:4d} ops  {100 * msk / tot:5.1f} %")
    print("slowest ops:")
    for dt, name, kern in sorted(rows, reverse=True)[:25]:
        print(f"  {dt:8.3f} ms  {kern:20s} {name}")
    if os.environ.get("PROF_ALL"):
        print("all backward ops (execution order):")
        for dt, name, kern in rows:
            print(f"  {dt:8.3f} ms  {kern:20s} {name}")
        # per-op forward profile of the TRAINING plan (same C call as the inference profiler)
        import ctypes as C
        from cv_diffusion_model_b200 import native
        cap = eng.lib.lcm_plan_launches_per_forward(eng.handle)
        recs = (native.OpProfileC * cap)()
        tt = t.to(torch.long).contiguous()
        epsb = torch.empty_like(eps)
        for _ in range(2):
            n = native.check(eng.lib.lcm_plan_profile_forward(
                eng.handle, C.c_void_p(noisy.data_ptr()), noisy.shape[1], noisy.stride(0), C.c_void_p(low.data_ptr()), low.shape[1],
                low.stride(0), C.c_void_p(tt.data_ptr()), C.c_void_p(epsb.data_ptr()), C.c_void_p(eng.workspace.data_ptr()),
                C.c_void_p(torch.cuda.current_stream().cuda_stream), recs, cap))
        print("forward ops of the training plan:")
        fby = defaultdict(lambda: [0.0, 0])
        for i in range(min(n, cap)):
            r = recs[i]
            fby[r.kernel.decode()][0] += r.ms; fby[r.kernel.decode()][1] += 1
            print(f"  {r.ms:8.3f} ms  {r.kernel.decode():16s} {r.name.decode():44s} {r.bytes / max(r.ms, 1e-9) / 1e6:8.0f} GB/s")
        for k, (msk, nn) in sorted(fby.items(), key=lambda kv: -kv[1][0]):
            print(f"  fwd {k:18s} {msk:8.3f} ms {nn:4d} ops")


if __name__ == "__main__":
    main()
