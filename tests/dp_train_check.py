#!/usr/bin/env python
"""Multi-GPU check (not collected by pytest; run under torchrun on N GPUs of one box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dp_train_check.py

Data-parallel NativeTrainer (global batch sharded by image, NCCL bucketed all-reduce overlapped with backward) must
reproduce single-GPU training on the whole global batch: same losses (mean of the rank losses), same gradient norm, same
weights after two steps.  fp32 plan; tolerance 2 % of each tensor's update norm (ReLU6 kinks, see tests/test_gpu_train.py).
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cv_diffusion_model_b200 import LowLightDiffusion  # noqa: E402
from cv_diffusion_model_b200.training import NativeTrainer  # noqa: E402
from tests.util import randomise_affine  # noqa: E402


def model(prec):
    torch.manual_seed(0)
    pipe = LowLightDiffusion(unet_variant="small", image_size=64, num_inference_steps=4, precision=prec)
    randomise_affine(pipe.unet)
    return pipe.cuda().train()


def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    per, S, steps = 2, 64, 2
    GB = per * world
    g = torch.Generator().manual_seed(77)
    batches = []
    for _ in range(steps):
        high = torch.rand(GB, 3, S, S, generator=g) * 2 - 1
        low = ((high + 1) / 2) ** 3 * 2 - 1
        t = torch.randint(0, 1000, (GB,), generator=g)
        noise = torch.randn(GB, 3, S, S, generator=g)
        batches.append((low, high, t, noise))
    pipe = model(prec)
    w0 = {n: p.detach().clone() for n, p in pipe.unet.named_parameters()}
    tr = NativeTrainer(pipe, batch=per, precision=prec, n_buckets=4)
    sl = slice(rank * per, (rank + 1) * per)
    losses, norms = [], []
    for low, high, t, noise in batches:
        loss = tr.train_step(low[sl].cuda(), high[sl].cuda(), timesteps=t[sl].cuda(), noise=noise[sl].cuda())
        dist.all_reduce(loss, op=dist.ReduceOp.SUM)
        losses.append(loss.item() / world)
        norms.append(tr.grad_norm().item())
    ok = True
    if rank == 0:
        ref = model(prec)
        tr1 = NativeTrainer(ref, batch=GB, precision=prec, process_group=None)
        tr1.world, tr1.buckets, tr1._comm = 1, tr1.buckets[:1], None      # plain single-GPU step on the global batch
        l1, n1 = [], []
        for low, high, t, noise in batches:
            l1.append(tr1.train_step(low.cuda(), high.cuda(), timesteps=t.cuda(), noise=noise.cuda()).item())
            n1.append(tr1.grad_norm().item())
        worst = 0.0
        for (n, p), (_, q) in zip(pipe.unet.named_parameters(), ref.unet.named_parameters()):
            d_dp, d_1 = (p.detach() - w0[n]), (q.detach() - w0[n])
            worst = max(worst, (d_dp - d_1).norm().item() / max(d_1.norm().item(), 1e-12))
        tol_l, tol_w = (1e-5, 0.02) if prec == "fp32" else (2e-3, 0.2)
        print(f"DP {world} GPUs vs 1 GPU [{prec}]: losses {losses} vs {l1}; grad norms {norms} vs {n1}; worst update deviation {worst:.3e}")
        ok = all(abs(a - b) <= tol_l * abs(b) for a, b in zip(losses, l1)) and all(abs(a - b) <= 1e-2 * b + (0 if prec == "fp32" else 0.05 * b) for a, b in zip(norms, n1)) \
            and worst <= tol_w
        print("DP_TRAIN_OK" if ok else "DP_TRAIN_MISMATCH")
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
